"""Shared helpers for the parity tests: inputs of the golden cases and runners for each side."""
import hashlib
import json
import os

import numpy as np

from golden.cases import CASES, NESTED, case_dict, nested_dict
from oracle_bind import FMT_CF32, Oracle, fnv1a64, synth_anchor, synth_raw, unpack

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
ALL_CASES = [case_dict(c) for c in CASES]
NESTED_CASES = [nested_dict(c) for c in NESTED]


def case_fmt(d):
    return FMT_CF32 if d["input"] == "anchor" else d["input"][1]


def raw_block(d, b):
    """Block b of case d in its RAW format (what the GPU bank ingests)."""
    if d["input"] == "anchor":
        return synth_anchor(b * d["B"], d["B"])
    _, fmt, seed, amp = d["input"]
    return synth_raw(fmt, b * d["B"], d["B"], seed, amp)


def float_block(d, b):
    """Block b as interleaved float32 (what the CPU chain ingests)."""
    raw = raw_block(d, b)
    return raw if raw.dtype == np.float32 else unpack(case_fmt(d), raw)


def run_oracle(d):
    o = Oracle(d["Fs"], d["B"], d["D"], d["L"], d["mixer"], d["gain"], d["filter_bw"], d["demod_usb"], d["cstyle"], d["scalecomp"])
    blocks = [o.process(float_block(d, b)) for b in range(d["blocks"])]
    stage = o.stage(d["D"])
    rate = o.out_rate
    o.close()
    return blocks, stage, rate


def check_against_golden(name, blocks, stage, rate):
    g = GOLDEN[name]
    allb = b"".join(blocks)
    assert len(allb) == g["payload_bytes"]
    assert rate == g["rate"]
    for i, blk in enumerate(blocks):
        assert blk[:16].hex() == g["block_head_hex"][i], "block %d head differs" % i
        assert hashlib.sha256(blk).hexdigest() == g["block_sha256"][i], "block %d differs from the reference" % i
    assert hashlib.sha256(allb).hexdigest() == g["sha256"]
    assert "%016x" % fnv1a64(allb) == g["fnv1a64"]
    if stage is not None:
        got = np.asarray(stage[:16], np.float32)
        want = np.array(g["stage_d_last_head_u32"], np.uint32).view(np.float32)
        assert np.array_equal(got, want), "stage-D head differs from the reference"  # == treats -0 and +0 alike


def parity_metrics(got_i16, want_i16):
    """SURVEY.md section 8d parity metric: (max |err| / 32768, error SNR in dB)."""
    g = got_i16.astype(np.float64)
    w = want_i16.astype(np.float64)
    err = g - w
    maxerr = float(np.abs(err).max()) / 32768.0 if err.size else 0.0
    pe = float((err * err).sum())
    ps = float((w * w).sum())
    snr = float("inf") if pe == 0 else 10.0 * np.log10(ps / pe)
    return maxerr, snr


def run_nested_oracle(d):
    """Main VFO -> sub-VFOs with the oracle: the main's stage-D stream is the subs' cf32 input block
    (vfo.cpp:167-172). Returns {topic: [payload per block]}, {topic: rate}."""
    fm, Dm = d["main"]
    main = Oracle(d["Fs"], d["B"], Dm, 0, fm, 0.01, 0, 0, 1, 1)
    fs_sub, blk_sub = d["Fs"] >> Dm, d["B"] >> Dm
    subs = [Oracle(fs_sub, blk_sub, D, L, f, g, bw) for (f, D, L, g, bw) in d["subs"]]
    out = {"S%04d" % i: [] for i in range(len(subs))}
    for b in range(d["blocks"]):
        main.process(float_block(d, b))
        mid = main.stage(Dm)
        for i, s_ in enumerate(subs):
            out["S%04d" % i].append(s_.process(mid))
    rates = {"S%04d" % i: s_.out_rate for i, s_ in enumerate(subs)}
    return out, rates


def check_nested_golden(name, out, rates):
    g = GOLDEN[name]
    assert sorted(out) == sorted(g)
    for t, blocks in out.items():
        assert rates[t] == g[t]["rate"]
        assert sum(len(b) for b in blocks) == g[t]["bytes"]
        for i, blk in enumerate(blocks):
            assert hashlib.sha256(blk).hexdigest() == g[t]["block_sha256"][i], "%s %s block %d differs from the reference" % (name, t, i)
