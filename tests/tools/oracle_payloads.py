#!/usr/bin/env python
"""TEST INFRASTRUCTURE: run the CPU chain (the compiled reference if present, else the oracle port) over an
IQ file with the VFO tree of a settings file and write per-topic payload dumps in the same layout as
`aero-publish-b200 --dump DIR`, so that CPU and GPU payloads can be compared byte for byte
(`cmp cpu/VFO01.i16 gpu/VFO01.i16`) and replayed into an unchanged aero-decode (tools/replay_payloads.py).

    tests/tools/oracle_payloads.py settings.ini capture.cu8 cu8 cpu_dump/ [dcc]
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_bind import Oracle, dc_correct, unpack  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "aero-cli_b200", "aero-publish-b200")
FMT = {"cu8": (0, np.uint8), "cs16": (1, np.int16), "cf32": (2, np.float32)}


def main():
    ini, iq, fmt, out = sys.argv[1:5]
    dcc = len(sys.argv) > 5 and sys.argv[5] in ("dcc", "1")
    dc_state = np.zeros(2, np.float32)
    os.makedirs(out, exist_ok=True)
    plan = json.loads(subprocess.run([BIN, "--plan", ini], check=True, capture_output=True, text=True).stdout)
    Fs, B = plan["sample_rate"], plan["block"]
    code, dt = FMT[fmt]
    mains, leaves = [], []
    for v in plan["vfos"]:
        if v["kind"] == "main":
            mains.append(Oracle(Fs, B, v["decim"], 0, v["mixer"], 0.01, 0, 0, 1, 1))
        else:
            leaves.append((v, Oracle(v["fs"], v["block"], v["decim"], v["late"], v["mixer"], v["gain"], v["filter_bw"])))
    files = {}
    for v, o in leaves:
        t = v["topic"][:5]
        files[t] = open(os.path.join(out, t + ".i16"), "wb")
        open(os.path.join(out, t + ".meta"), "w").write("%d %d\n" % (o.out_rate, o.out_bytes))
    with open(iq, "rb") as f:
        nblk = 0
        while True:
            raw = np.frombuffer(f.read(2 * B * np.dtype(dt).itemsize), dt)
            if raw.size < 2 * B:
                break
            x = raw if code == 2 else unpack(code, raw)
            if dcc:      # Publisher::demodData's DC removal ahead of every VFO (publisher.cpp:292-296)
                x = dc_correct(np.array(x, np.float32), dc_state)
            mid = []
            for m in mains:
                m.process(x)
                mid.append(m.stage(m.D))
            for v, o in leaves:
                files[v["topic"][:5]].write(o.process(mid[v["parent"]] if v["parent"] >= 0 else x))
            nblk += 1
    for fh in files.values():
        fh.close()
    print("wrote %d blocks for %d topics to %s" % (nblk, len(files), out))


if __name__ == "__main__":
    main()
