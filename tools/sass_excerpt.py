#!/usr/bin/env python
"""Writes profiles/sass_noise_grid2_octave_loop_rNN.txt: the SASS of ONE octave loop of the headline kernel (cuobjdump -sass of the built library) with its
opcode histogram - the evidence for the FFMA2 / FMUL2 / LDS counts DESIGN.md quotes.   python tools/sass_excerpt.py profiles/sass_noise_grid2_octave_loop_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "3dworld_b200", "lib3dworld_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.split("\n")
keep, on = [], False
for ln in sass:
    if "Function : " in ln:
        on = "noise_grid2_kernelILb1ELb1ELi0E" in ln
    if on:
        keep.append(ln)
ins = [(int(m.group(1), 16), m.group(2)) for ln in keep for m in [re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)] if m]
loops = []
for a, t in ins:
    m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        body = [(x, u) for x, u in ins if int(m.group(1), 16) <= x <= a]
        if 100 < len(body) < 400 and sum(1 for _, u in body if "FFMA2" in u or "FMUL2" in u or "FADD2" in u) > 40:
            loops.append(body)
body = loops[0]
hist = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for _, t in body)
out = ["SASS excerpt (cuobjdump -sass 3dworld_b200/lib3dworld_b200.so, sm_100a) of noise_grid2_kernel<simplex, warp, shape 0>: ONE octave loop of gen_noise2",
       "(= one fBm octave for a pair of cells; the kernel contains %d such loops, one per gen_noise2 call of the domain warp). Loop 0x%04x .. 0x%04x, %d instructions:" % (len(loops), body[0][0], body[-1][0], len(body)),
       "  " + ", ".join("%s %d" % kv for kv in hist.most_common()),
       "FMA-pipe issue cycles per loop iteration and warp: 2 x (FFMA2 + FMUL2 + FADD2) + FMUL + FADD + FFMA = %d; packed instructions are the reference's unfused multiplies / adds for two cells"
       % (2 * (hist["FFMA2"] + hist["FMUL2"] + hist["FADD2"]) + hist["FMUL"] + hist["FADD"] + hist["FFMA"]),
       "(a packed add of two products is fma(x, ONE, y) with an opaque ONE, every other packed add a plain FADD2, see csrc/tw_noise2.cuh); LDS = hash / gradient table look-ups; FRND = floor().", ""]
out += ["        /*%04x*/  %s ;" % (x, t) for x, t in body]
open(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "/dev/stdout", "w").write("\n".join(out) + "\n")
