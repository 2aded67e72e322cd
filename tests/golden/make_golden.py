"""Generate tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref/libref_vfo.so,
built from /root/reference/publish by oracle/Makefile). Needs the reference tree, so it runs in the
build container only; the JSON it writes is what travels.

For every case: FNV-1a-64 and SHA-256 of the concatenated ZMQ frame-3 payloads, SHA-256 per block,
the output rate, the first 16 payload bytes of every block, and the first 8 stage-D complex samples
of the last block (as float32 bit patterns).

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_bind import RefVfo, fnv1a64, ref_lib, synth_anchor, synth_raw, unpack  # noqa: E402
from golden.cases import CASES, NESTED, case_dict, nested_dict  # noqa: E402


def block_input(spec, n0, n):
    if spec == "anchor":
        return synth_anchor(n0, n)
    _, fmt, seed, amp = spec
    return unpack(fmt, synth_raw(fmt, n0, n, seed, amp))


def main():
    if ref_lib() is None:
        raise SystemExit("oracle/_ref/libref_vfo.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for c in CASES:
        d = case_dict(c)
        r = RefVfo(d["Fs"], d["B"], d["D"], d["L"], d["mixer"], d["gain"], d["filter_bw"], d["demod_usb"], d["cstyle"], d["scalecomp"])
        allb, per_block, heads, rate = b"", [], [], 0
        for b in range(d["blocks"]):
            x = block_input(d["input"], b * d["B"], d["B"])
            msgs = r.process(x)
            rate, payload = msgs[r.topic]
            allb += payload
            per_block.append(hashlib.sha256(payload).hexdigest())
            heads.append(payload[:16].hex())
        st = r.stage(d["D"])[:16]
        out[d["name"]] = {
            "payload_bytes": len(allb),
            "fnv1a64": "%016x" % fnv1a64(allb),
            "sha256": hashlib.sha256(allb).hexdigest(),
            "block_sha256": per_block,
            "block_head_hex": heads,
            "rate": rate,
            "stage_d_last_head_u32": [int(v) for v in st.view(np.uint32)],
        }
        print(d["name"], len(allb), out[d["name"]]["fnv1a64"], rate)
        r.close()
    for c in NESTED:
        d = nested_dict(c)
        fm, Dm = d["main"]
        main = RefVfo(d["Fs"], d["B"], Dm, 0, fm, 0.01, 0, 0, 1, 1, topic="MAIN0")   # publisher.cpp:136-147
        fs_sub, blk_sub = d["Fs"] >> Dm, d["B"] >> Dm
        subs = []
        for i, (f, D, L, g, bw) in enumerate(d["subs"]):
            s_ = RefVfo(fs_sub, blk_sub, D, L, f, g, bw, 1, 1, 1, topic="S%04d" % i)     # publisher.cpp:196-217
            main.add_sub(s_)
            subs.append(s_)
        per = {s_.topic: {"block_sha256": [], "rate": 0, "bytes": 0} for s_ in subs}
        for b in range(d["blocks"]):
            msgs = main.process(block_input(d["input"], b * d["B"], d["B"]))
            assert "MAIN0" not in msgs          # a main VFO with sub-VFOs publishes nothing itself (vfo.cpp:167-172)
            for t, (rate, payload) in msgs.items():
                per[t]["block_sha256"].append(hashlib.sha256(payload).hexdigest())
                per[t]["rate"] = rate
                per[t]["bytes"] += len(payload)
        out[d["name"]] = per
        print(d["name"], {t: (v["rate"], v["bytes"]) for t, v in per.items()})
        main.close()
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
