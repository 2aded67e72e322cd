"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/aeroddc.h declares,
refuses to compute without a CUDA device (no CPU fallback), and its host-side designers agree
bit for bit with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import aeroddc
from oracle_bind import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "aeroddc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aeroddc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = aeroddc.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libaeroddc.so does not export %s" % s
    assert sorted(aeroddc.ABI_SYMBOLS) == syms, "binding list out of date with include/aeroddc.h"
    assert lib.aeroddc_abi_version() == 1


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    with pytest.raises(aeroddc.AeroDdcError) as e:
        aeroddc.Bank(288000, 57600, aeroddc.CF32, 0)
    assert "no CUDA device" in str(e.value) or "fallback" in str(e.value)


def test_argument_errors_are_reported_before_any_device_use():
    lib = aeroddc.lib()
    h = ctypes.c_void_p()
    assert lib.aeroddc_bank_create(ctypes.byref(h), 0, 100, 2, 0) == -1          # bad rate
    assert lib.aeroddc_bank_create(ctypes.byref(h), 288000, 288016, 2, 0) == -1  # block longer than Fs (dsp.cpp:43)
    assert lib.aeroddc_bank_create(ctypes.byref(h), 288000, 57600, 7, 0) == -1   # unknown format
    assert lib.aeroddc_bank_create(ctypes.byref(h), 288000, 57601, 2, 0) == -1   # not a multiple of 16
    assert b"multiple" in lib.aeroddc_last_error()
    assert lib.aeroddc_bank_process(None, None, 0) == -1
    assert lib.aeroddc_bank_wait(None) == -1


@pytest.mark.parametrize("args", [(2, 240000, 24000, 12000.0), (2, 288000, 24000, 9600.0), (2, 48000, 12000, 3000.0),
                                  (2, 48000, 6000, 1500.0), (2, 48000, 3000, 750.0), (2, 48000, 1500, 375.0), (2, 24000, 3000, 750.0)])
def test_lowpass_design_matches_oracle(args):
    got = aeroddc.design_lowpass(*args)
    buf = np.zeros(8192, np.float32)
    n = oracle_lib().ddc_lowpass_taps(*args, buf.ctypes.data, 8192)
    assert n == got.size
    assert np.array_equal(got.view(np.uint32), buf[:n].view(np.uint32))
    assert abs(float(got.astype(np.float64).sum()) - 2.0) < 1e-5          # DC gain 2 (vfo.cpp:71-79)
    assert got.size % 2 == 1


def test_lowpass_design_rejects_like_the_reference():
    lib = aeroddc.lib()
    assert lib.aeroddc_design_lowpass(2.0, 48000.0, 30000.0, 100.0, None, 0) == -4   # cutoff above fs/2 (firfilter.cpp:100-112)
    assert lib.aeroddc_design_lowpass(2.0, 48000.0, 3000.0, 0.0, None, 0) == -4      # zero transition width


@pytest.mark.parametrize("fs", [12000, 2400, 9600, 15000, 90])
def test_hilbert_design_matches_oracle(fs):
    got = aeroddc.design_hilbert(125, fs)
    want = np.zeros(125, np.float32)
    oracle_lib().ddc_hilbert_taps(125, fs, want.ctypes.data)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert got[62] == 0.0 and abs(float((got.astype(np.float64) ** 2).sum()) - 1.0) < 1e-6


def test_rotation_matches_oracle():
    O = oracle_lib()
    O.ddc_nco_rotation.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    for fs, f in [(61440000.0, 1234567.0), (288000.0, -34567.0), (1536000.0, 0.0), (2400000.0, 1199999.0)]:
        c, s = ctypes.c_float(), ctypes.c_float()
        O.ddc_nco_rotation(fs, f, ctypes.byref(c), ctypes.byref(s))
        assert aeroddc.design_rotation(fs, f) == (c.value, s.value)


def test_segment_plan_invariants():
    """Host arithmetic that cuts a block into segments and chained parts (aeroddc_plan_segments)."""
    import math

    rng = np.random.default_rng(0)
    cases = [(15360000, 8, 1024), (15360000, 8, 128), (384000, 5, 30), (57600, 1, 1), (60000, 0, 3), (480000, 3, 2), (5120, 8, 2)]
    for _ in range(300):
        D = int(rng.integers(0, 9))
        step = math.lcm(32, 1 << D)
        blk = step * int(rng.integers(max(1, (20 << D) // step + 1), 4000))
        cases.append((blk, D, int(rng.integers(1, 3000))))
    for blk, D, nv in cases:
        for waves, parts in ((1.0, 0), (0.25, 1), (3.0, 5), (1.0, 60)):
            p = aeroddc.plan_segments(blk, D, nv, 148, waves, parts)
            DA = min(D, 5)                          # stages of the main kernel; stages 5..D-1 run in the deep kernel
            al = math.lcm(256, 1 << DA)
            assert p["warmup"] % math.lcm(32, 1 << DA) == 0 and p["warmup"] >= (10 * ((1 << DA) - 1) if DA else 0)
            assert p["warmup"] <= 320
            assert p["boundary_warmup"] == (11 << D if D else 0)
            assert p["segment_len"] % al == 0 and p["part_len"] % al == 0
            assert p["n_segments"] * p["segment_len"] >= blk > (p["n_segments"] - 1) * p["segment_len"]
            assert p["parts"] >= 1 and p["parts"] * p["part_len"] >= p["segment_len"]
            assert p["segment_len"] >= 2 * p["warmup"]
            assert p["vfo_groups"] == -(-nv // 32)   # one warp of 32 VFOs per CTA
            per_group = p["ctas"] // p["vfo_groups"] - 1           # parts of all segments; the last segment may need fewer
            assert p["ctas"] % p["vfo_groups"] == 0 and p["n_segments"] <= per_group <= p["parts"] * p["n_segments"]
            assert per_group * p["part_len"] >= blk
    with pytest.raises(aeroddc.AeroDdcError):
        aeroddc.plan_segments(57601, 1, 1)


@pytest.mark.parametrize("n_tiles,n_mid,n_sm", [(16, 480000, 148), (2, 480000, 148), (1, 12288, 148), (3, 15360, 148), (1, 1024, 148), (5, 99968, 132), (16, 480000, 7)])
def test_tensor_mode_stretch_plan_covers_every_output_once(n_tiles, n_mid, n_sm):
    """aeroddc_plan_tensor_stretches (host side of AERODDC_MODE_TENSOR): every (64-VFO tile, stage-5 output) belongs to exactly one
    stretch, a CTA's stretches are consecutive in tile-major order, stretches start at 0 or at a multiple of 256 (the fused
    stages' run-in reaches 96 outputs back, never before the block), and the CTAs' loads differ by less than two tiles."""
    import aeroddc

    P, st = aeroddc.plan_tensor_stretches(n_tiles, n_mid, n_sm)
    assert 1 <= P <= n_sm
    cover = {t: [] for t in range(n_tiles)}
    load = [0] * P
    last = -1
    for c, t, lo, hi in st:
        assert 0 <= t < n_tiles and 0 <= lo < hi <= n_mid
        assert lo == 0 or (lo % 256 == 0 and lo >= 96)
        assert hi == n_mid or hi % 256 == 0
        assert t * n_mid + lo > last                      # tile-major, increasing
        last = t * n_mid + lo
        cover[t].append((lo, hi))
        load[c] += hi - lo
    for t in range(n_tiles):
        pos = 0
        for lo, hi in sorted(cover[t]):
            assert lo == pos
            pos = hi
        assert pos == n_mid
    assert sum(load) == n_tiles * n_mid
    busy = [x for x in load if x]
    assert max(busy) - min(busy) <= 512 or len(busy) == 1
