"""CPU: static evidence on the built library's SASS (cuobjdump) for the instruction-level properties DESIGN.md section 4 claims for the headline kernel's octave loop:
packed fp32x2 arithmetic, plain FADD2 for the adds without a product operand, no integer adds or float->int conversions for the table addressing, 12 table
look-ups per cell pair. A guard against silently losing them (a flag change, a compiler update), not a performance test. Skipped when cuobjdump is not installed."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_headline_octave_loop_instruction_mix(tw, tmp_path):
    out = str(tmp_path / "loop.txt")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "sass_excerpt.py"), out])
    text = open(out).read()
    hist = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"([A-Z0-9_]+) (\d+)(?:,|\n)", text.split("\n")[2] + "\n"))
    assert hist.get("FMUL2", 0) >= 20 and hist.get("FFMA2", 0) >= 10 and hist.get("FADD2", 0) >= 15, hist      # packed arithmetic; plain adds where allowed
    assert hist.get("FFMA2", 0) <= 20, hist                                                                    # (the opaque-ONE form only for the nine sums of products + genuine fmas)
    assert hist.get("LDS", 0) == 12, hist                                                                      # 6 hash + 6 gradient look-ups per cell pair
    assert hist.get("IADD3", 0) <= 2 and "F2I" not in hist and "I2F" not in hist, hist                       # denormal addressing: the FMA result is the address
    assert hist.get("FRND", 0) == 8, hist                                                                      # the floors stay on the XU pipe
    assert sum(hist.values()) <= 130, hist
