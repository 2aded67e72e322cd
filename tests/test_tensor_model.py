"""CPU suite: the algebra of AERODDC_MODE_TENSOR (aero-cli_b200/csrc/tc_kernels.cuh) against the oracle.

The tensor kernel computes stage-5 sample m of a VFO as  q(n_s) * sum_t (g[t] u^t) x[n_s + t]  over a 320-sample padded
window: g = the five half-band stages collapsed into one 311-tap FIR, u = rot / |rot|, q(n_s) = the reference
oscillator at the window start (exact checkpoint every 256 steps times a unit rotation), operands split into bf16
hi + mid with three products kept. Here the same formula runs in numpy on the samples of one block and is compared
with the stage-5 stream of the oracle's bit-exact chain (oscillator.cpp:19-24, vfo.cpp:155-161,
halfbanddecimator.cpp:35-60): it must sit far inside BASELINE.json's tolerance (max |err| <= 1e-4 of full scale, error
SNR >= 80 dB). The GPU tests (tests/test_tensor_mode.py) check the kernel; this one pins the formulation without a GPU."""
import importlib.util
import os

import numpy as np
import pytest

from oracle_bind import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("tc_model", os.path.join(ROOT, "scratch", "tc_model.py"))
tc = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tc)


@pytest.mark.parametrize("freq", [123456.0, -654321.0])
def test_factored_windowed_fir_with_bf16_split_matches_the_oracle_stage5(freq):
    fs, n = 1536000, 32 * 400
    rng = np.random.default_rng(int(abs(freq)))
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.1).astype(np.complex64)
    iq = np.empty(2 * n, np.float32)
    iq[0::2], iq[1::2] = x.real, x.imag
    o = Oracle(fs, n, 5, 0, freq, 0.5)
    o.process(iq)
    st = o.stage(5)
    o.close()
    ref = st[0::2].astype(np.float64) + 1j * st[1::2].astype(np.float64)      # 400 stage-5 samples, zero history before the block

    ang = 2 * np.pi * freq / fs
    c, d = np.float32(np.cos(ang)), np.float32(np.sin(ang))
    S = tc.nco_table(c, d, n + 1)                                              # S(k): the recurrence after k steps from (1, 0)
    theta = np.arctan2(float(d), float(c))
    g = tc.composite()
    assert len(g) == 311 and abs(g.sum() - 1.0) < 1e-6                        # five unity-gain half-band stages
    tp = np.arange(320)
    gp = np.zeros(320)
    gp[2:313] = g
    G = gp * np.exp(1j * theta * tp)
    Gr_h, Gr_m = tc.split2(G.real)
    Gi_h, Gi_m = tc.split2(G.imag)
    xd = x.astype(np.complex128)
    xr_h, xr_m = tc.split2(xd.real)
    xi_h, xi_m = tc.split2(xd.imag)

    def prod(ar, ai, br, bi):
        return np.sum(ar * br - ai * bi) + 1j * np.sum(ar * bi + ai * br)

    ms = np.arange(64, n // 32)                                                # past the block head, which stays on the FP32 kernel
    z64 = np.empty(len(ms), np.complex128)
    zbf = np.empty(len(ms), np.complex128)
    for k, m in enumerate(ms):
        ns = 32 * m - 312
        i1 = ns + 1                                                            # sample ns is mixed with S(ns + 1)
        ck = i1 >> 8
        fac = S[256 * ck] * np.exp(1j * theta * (i1 - 256 * ck))              # exact checkpoint x unit rotation
        sl = slice(ns, ns + 320)
        z64[k] = fac * np.sum(G * xd[sl])
        zbf[k] = fac * (prod(Gr_h, Gi_h, xr_h[sl], xi_h[sl]) + prod(Gr_h, Gi_h, xr_m[sl], xi_m[sl]) + prod(Gr_m, Gi_m, xr_h[sl], xi_h[sl]))
    r = ref[ms]

    def snr(a):
        return 10 * np.log10(np.sum(np.abs(r) ** 2) / np.sum(np.abs(a - r) ** 2))

    assert np.sqrt(np.mean(np.abs(r) ** 2)) > 0.01                            # the comparison is not vacuous
    assert snr(z64) >= 115.0, snr(z64)                                        # the factoring itself: float32 noise of the exact chain
    assert snr(zbf) >= 100.0, snr(zbf)                                        # with the bf16 hi + mid operand split
    assert np.max(np.abs(zbf - r)) <= 1e-5                                    # << 1e-4 of full scale
