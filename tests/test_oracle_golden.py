"""CPU suite, part 1: the oracle (oracle/ddc_oracle.c) is pinned against
  * the four SURVEY.md section-8c anchors (FNV-1a-64 of the reference's ZMQ payloads), and
  * tests/golden/golden.json, produced by running the unmodified reference (tests/golden/make_golden.py),
and, where the compiled reference is present (this container), against the reference directly on
further seeded configurations and per component (filter designs, NCO table)."""
import numpy as np
import pytest

from case_util import ALL_CASES, NESTED_CASES, check_against_golden, check_nested_golden, float_block, run_nested_oracle, run_oracle
from oracle_bind import Oracle, RefVfo, oracle_lib, ref_lib

SURVEY_ANCHORS = {
    "anchor_1536k_d5": ("4435d1a5843eed39", [-644, -675, -183, 821, 178, 897]),
    "anchor_288k_d1_l6": ("78e7d3dd6abf468a", [5883, 5489, -405, 1918, -4929, -6487]),
    "anchor_288k_d0_l6_bw": ("4d86251b7e306418", [1251, 1831, 2214, 2378, 2305, 2004]),
    "anchor_1920k_d3_l5": ("c315edd09d50c16e", [-125, 1225, -108, -832, -853, -1204]),
}


@pytest.mark.parametrize("d", ALL_CASES, ids=[c["name"] for c in ALL_CASES])
def test_oracle_matches_reference_golden(d):
    blocks, stage, rate = run_oracle(d)
    check_against_golden(d["name"], blocks, stage, rate)
    if d["name"] in SURVEY_ANCHORS:
        from oracle_bind import fnv1a64

        fnv, last6 = SURVEY_ANCHORS[d["name"]]
        assert "%016x" % fnv1a64(b"".join(blocks)) == fnv
        assert list(np.frombuffer(blocks[-1], np.int16)[:6]) == last6


@pytest.mark.parametrize("d", NESTED_CASES, ids=[c["name"] for c in NESTED_CASES])
def test_oracle_nested_matches_reference_golden(d):
    out, rates = run_nested_oracle(d)
    check_nested_golden(d["name"], out, rates)


def test_oracle_rejects_contract_violations():
    with pytest.raises(ValueError):
        Oracle(288000, 57601, 1, 0, 0.0, 0.5)  # B not divisible by 2^D
    with pytest.raises(ValueError):
        Oracle(288000, 57600, 9, 0, 0.0, 0.5)  # more than 8 stages (vfo.h:63)
    with pytest.raises(ValueError):
        Oracle(288000, 57600, 4, 7, 0.0, 0.5)  # 3600 stage-D samples not divisible by 7


needs_ref = pytest.mark.skipif(ref_lib() is None, reason="compiled reference (oracle/_ref) not present on this box")


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_reference_random_configs(seed):
    """Seeded random VFO configurations, stage-D floats and payload bytes, block by block."""
    rng = np.random.default_rng(100 + seed)
    fs, blk = [(288000, 57600), (1536000, 384000), (1920000, 480000), (288000, 72000)][seed % 4]
    D = int(rng.integers(0, 6))
    L = [0, 5, 6][seed % 3]
    if L and (blk >> D) % L:
        L = 0
    bw = [0, 3000, 1500][int(rng.integers(0, 3))] if (fs >> D) // max(L, 1) >= 12000 else 0
    f = float(rng.integers(-fs // 2 + 1000, fs // 2 - 1000))
    gain = float(rng.uniform(0.05, 0.6))
    o = Oracle(fs, blk, D, L, f, gain, bw)
    r = RefVfo(fs, blk, D, L, f, gain, bw)
    for b in range(6):   # crosses five block boundaries and, at B = Fs/5 or Fs/4, an NCO table wrap
        x = (rng.standard_normal(2 * blk) * 0.2).astype(np.float32)
        po = o.process(x)
        pr = r.process(x)[r.topic][1]
        assert np.array_equal(o.stage(D), r.stage(D)), "stage-D differs in block %d" % b
        assert po == pr, "payload differs in block %d" % b


@needs_ref
def test_components_vs_reference():
    O, R = oracle_lib(), ref_lib()
    for args in [(2, 240000, 24000, 12000.0), (2, 48000, 12000, 3000.0), (2, 48000, 3000, 750.0), (2, 288000, 24000, 9600.0),
                 (2, 24000, 6000, 1500.0), (2, 12000, 1500, 375.0)]:
        a = np.zeros(8192, np.float32)
        b = np.zeros(8192, np.float32)
        na = O.ddc_lowpass_taps(*args, a.ctypes.data, 8192)
        nb = R.ref_lowpass(*args, b.ctypes.data, 8192)
        assert na == nb and np.array_equal(a.view(np.uint32), b.view(np.uint32)), args
    for fs in (12000, 2400, 15000, 9600, 90):
        a = np.zeros(125, np.float32)
        b = np.zeros(125, np.float32)
        O.ddc_hilbert_taps(125, fs, a.ctypes.data)
        R.ref_hilbert(125, fs, b.ctypes.data)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), fs
    for fs, f in [(288000.0, -34567.0), (1536000.0, 123456.0), (288000.0, 0.0)]:
        L = int(fs)
        q = np.zeros(2 * L, np.float32)
        O.ddc_nco_table(fs, f, q.ctypes.data)
        first = L - 100
        r = np.zeros(2 * 300, np.float32)
        R.ref_nco(fs, f, first, 300, r.ctypes.data)   # values vfo::process would use for samples first..first+299
        want = np.concatenate([q[2 * first:], q[: 2 * 200]])
        assert np.array_equal(r.view(np.uint32), want.view(np.uint32))
        r0 = np.zeros(4, np.float32)
        R.ref_nco(fs, f, 0, 2, r0.ctypes.data)
        assert np.array_equal(r0[:2], q[-2:]) and np.array_equal(r0[2:], q[2:4])   # sample 0 uses q[L-1]


@needs_ref
def test_oracle_int16_overflow_matches_reference_build():
    """Out-of-range float->short is undefined in C++; the oracle pins what the reference's x86 build does."""
    fs, blk = 288000, 57600
    from oracle_bind import synth_anchor
    for (f, D, L, gain) in [(12345.0, 2, 0, 40.0), (-3000.0, 1, 6, 1e6)]:
        o = Oracle(fs, blk, D, L, f, gain, 0)
        r = RefVfo(fs, blk, D, L, f, gain, 0)
        for b in range(3):
            x = synth_anchor(b * blk, blk)
            assert o.process(x) == r.process(x)[r.topic][1]
