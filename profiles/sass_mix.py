#!/usr/bin/env python
"""Instruction mix of the chunk loop of ddc_main_kernel from the SASS of the shipped library (no GPU needed):
   python profiles/sass_mix.py [NF FMT FAST]     (default 5 2 0 = D >= 5, cf32, exact)
Splits the function at branch targets and prints the opcode histogram of every straight-line block of >= 400
instructions - those are the 32-sample chunk bodies (ordinary and oscillator-wrap variants, two copies from `unroll 2`)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nf, fmt, fast = (sys.argv[1:4] + ["5", "2", "0"][len(sys.argv) - 1:])[:3] if len(sys.argv) > 1 else ("5", "2", "0")
sym = "_ZN7aeroddc15ddc_main_kernelILi%sELi%sELb%sEEEvNS_10MainParamsE" % (nf, fmt, fast)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, os.path.join(ROOT, "aero-cli_b200", "libaeroddc.so")], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
targets = set()
for a, op, rest in ins:
    if op.startswith("BRA"):
        t = re.search(r"0x([0-9a-f]+)", rest)
        if t:
            targets.add(int(t.group(1), 16))
blocks, cur, start = [], [], ins[0][0]
for a, op, rest in ins:
    if a in targets and cur:
        blocks.append((start, cur)); cur = []; start = a
    cur.append(op.split(".")[0])
    if op.startswith(("BRA", "EXIT", "RET")):
        blocks.append((start, cur)); cur = []; start = a + 16
if cur:
    blocks.append((start, cur))
print("%s: %d SASS instructions" % (sym, len(ins)))
for s, b in blocks:
    if len(b) >= 400:
        c = collections.Counter(b)
        fp2, fp1 = c["FMUL2"] + c["FFMA2"] + c["FADD2"], c["FADD"] + c["FMUL"] + c["FFMA"]
        print("block @0x%x: %d instructions; packed FP32 %d, scalar FP32 %d -> %.1f FMA-pipe cycles per VFO-sample (2 cycles per warp instruction, 32 samples per chunk)"
              % (s, len(b), fp2, fp1, 2.0 * (fp2 + fp1) / 32))
        print("   " + ", ".join("%s %d" % kv for kv in c.most_common(14)))
