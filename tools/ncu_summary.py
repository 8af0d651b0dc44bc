#!/usr/bin/env python
"""Turns an ncu report (.ncu-rep, `ncu --set full --clock-control none`) into the small JSON summaries committed under profiles/.
Runs where ncu is installed (no GPU needed):   python tools/ncu_summary.py gpurun_out/prof.ncu-rep --kernel noise_grid2 -o profiles/ncu_x_r02.json
The numbers bench.py quotes as roofline.traffic / executed operations come from the file this writes (profiles/roofline_r02.json, --roofline)."""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "sm__sass_thread_inst_executed_op_fp32_pred_on.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
    body = rows[2:] if units else rows[1:]
    return hdr, units, body


def opcode_mix(rep, kernel):
    """Per-opcode executed counts from the SASS view of the source page (needs `--import-source on` only for the CUDA-C view, not for this):
    warp instructions and predicated-on thread instructions per mnemonic, and the fp32-pipe lane operations (packed FFMA2/FMUL2/FADD2 count twice)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    mix, cur, hdr, ok = {}, None, None, False
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur, hdr = r[1], None
            ok = (not kernel) or re.search(kernel, cur) is not None
            continue
        if not ok or not r:
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if hdr is None:
            continue
        src = r[hdr.index("Source")].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0] if src else "?"
        wi = float(r[hdr.index("Instructions Executed")] or 0)
        ti = float(r[hdr.index("Predicated-On Thread Instructions Executed")] or 0)
        m = mix.setdefault(op, [0.0, 0.0, 0])
        m[0] += wi; m[1] += ti; m[2] += 1
    fp32 = {"FFMA": 1, "FMUL": 1, "FADD": 1, "FFMA2": 2, "FMUL2": 2, "FADD2": 2, "FMNMX": 1, "FSET": 1, "FSETP": 1, "FSEL": 1, "FRND": 1, "FCHK": 1, "MUFU": 1, "F2I": 1, "I2F": 1, "F2F": 1, "I2FP": 1, "F2FP": 1, "FMNMX3": 1}
    lane_ops = sum(v[1] * fp32[k] for k, v in mix.items() if k in fp32)
    arith = sum(v[1] * fp32[k] for k, v in mix.items() if k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
    top = sorted(mix.items(), key=lambda kv: -kv[1][0])
    return {"warp_instructions_by_opcode": {k: int(v[0]) for k, v in top[:24]}, "static_sites_by_opcode": {k: v[2] for k, v in top[:24]},
            "warp_instructions_total": int(sum(v[0] for v in mix.values())), "thread_instructions_total": int(sum(v[1] for v in mix.values())),
            "fp32_lane_ops_executed": int(lane_ops), "fp32_mul_add_fma_lane_ops_executed": int(arith)}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default="", help="regex on the kernel name (first match wins unless --all)")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("-o", "--out", default="")
    ap.add_argument("--note", default="")
    ap.add_argument("--opcodes", action="store_true", help="add the executed per-opcode mix from the SASS view (slow for big kernels)")
    ap.add_argument("--roofline", help="also write the bench.py roofline input (profiles/roofline_rNN.json) from the first matched kernel")
    ap.add_argument("--units", type=float, default=0.0, help="work units of one launch (cells, voxels, droplet moves): adds per-unit figures")
    args = ap.parse_args()
    hdr, units, body = raw_rows(args.rep)
    name_col = hdr.index("Kernel Name")
    res = []
    for r in body:
        if args.kernel and not re.search(args.kernel, r[name_col]):
            continue
        d = {"kernel": r[name_col], "report": args.rep.split("/")[-1]}
        if args.note:
            d["note"] = args.note
        stalls = {}
        for i, h in enumerate(hdr):
            if h in KEEP:
                d[h] = num(r[i])
                if units and units[i]:
                    d.setdefault("_units", {})[h] = units[i]
            m = STALL.fullmatch(h)
            if m:
                stalls[m.group(1) or m.group(2)] = num(r[i])
        if stalls:
            d["stall_cycles_per_issue"] = {k: v for k, v in sorted(stalls.items(), key=lambda kv: -kv[1] if isinstance(kv[1], float) else 0)[:8]}
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if isinstance(rd, float) and isinstance(wr, float):
            def to_bytes(v, key):   # ncu scales every metric on its own (the read side can be in Kbyte while the write side is in Mbyte)
                u = (d.get("_units", {}).get(key, "byte") or "byte").lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
            d["dram_read_bytes"], d["dram_write_bytes"] = to_bytes(rd, "dram__bytes_read.sum"), to_bytes(wr, "dram__bytes_write.sum")
            d["dram_bytes_per_launch"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        if args.opcodes:
            d["opcode_mix"] = opcode_mix(args.rep, args.kernel)
        if args.units:
            d["units_per_launch"] = args.units
            if "dram_bytes_per_launch" in d:
                d["dram_bytes_per_unit"] = d["dram_bytes_per_launch"] / args.units
            if isinstance(d.get("smsp__inst_executed.sum"), float):
                d["warp_instructions_per_unit"] = d["smsp__inst_executed.sum"] / args.units
            if args.opcodes:
                d["fp32_lane_ops_per_unit"] = d["opcode_mix"]["fp32_lane_ops_executed"] / args.units
                d["fp32_mul_add_fma_lane_ops_per_unit"] = d["opcode_mix"]["fp32_mul_add_fma_lane_ops_executed"] / args.units
        res.append(d)
        if not args.all:
            break
    if not res:
        sys.exit("no kernel matched %r" % args.kernel)
    if args.roofline:   # the per-launch numbers bench.py quotes (roofline.traffic, executed fp32 operations per cell)
        d = res[0]
        keys = ("kernel", "report", "gpu__time_duration.sum", "dram_read_bytes", "dram_write_bytes", "dram_bytes_per_launch", "dram_bytes_per_unit", "units_per_launch",
                "warp_instructions_per_unit", "fp32_mul_add_fma_lane_ops_per_unit", "fp32_lane_ops_per_unit")
        entry = {k: d[k] for k in keys if k in d}
        if args.units:
            entry["algorithmic_bytes_per_launch"] = int(4 * args.units)
        json.dump({"_comment": "per-launch numbers of the dominant kernel of the headline step, written from the ncu --set full capture by tools/ncu_summary.py --roofline "
                               "(same capture as the ncu_*.json summary next to it); bench.py quotes dram_bytes_per_launch as roofline.traffic and "
                               "fp32_mul_add_fma_lane_ops_per_unit as the executed fp32 operations per cell",
                   "noise_grid2_kernel": entry}, open(args.roofline, "w"), indent=1)
    text = json.dumps(res if args.all else res[0], indent=1)
    if args.out:
        open(args.out, "w").write(text + "\n")
    try:
        print(text)
    except BrokenPipeError:
        pass


if __name__ == "__main__":
    main()
