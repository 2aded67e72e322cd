"""CPU model of the tensor-core formulation (scratch; numpy only).

Five half-band stages + the NCO mix collapse into one 311-tap complex FIR per VFO whose taps carry the VFO's
rotation: z[m] = osc(n_s) * sum_t (g[t] u^t) x[n_s + t], n_s = 32 m - 310 (window start), osc = the reference's
drifting float32 recurrence taken at the window start (from the exact checkpoints), u = rot / |rot|.
This script checks the algebra (index offsets, signs) and the error level of the bf16 hi+mid split against a
float64 run of the per-sample chain with the float32 oscillator emulated step by step.
"""
import numpy as np

P0, P2, P4, P5 = np.float32(0.0060431029837374152), np.float32(-0.049372515458761493), np.float32(0.29332944952052842), np.float32(0.5)
h = np.array([P0, 0, P2, 0, P4, P5, P4, 0, P2, 0, P0], dtype=np.float64)


def composite(stages=5):
    g = h.copy()
    for s in range(1, stages):
        up = np.zeros(10 * (1 << s) + 1)
        up[:: 1 << s] = h
        g = np.convolve(g, up)
    return g


def nco_table(c, d, n):
    """float32 recurrence of oscillator.cpp:19-24; returns S(k), k = 0..n (S(0) = (1, 0))."""
    f = np.float32
    a, b = f(1), f(0)
    out = np.empty(n + 1, dtype=np.complex128)
    out[0] = 1
    for k in range(1, n + 1):
        nr = f(f(a * c) - f(b * d))
        ni = f(f(a * d) + f(b * c))
        nm = f(f(1.95) - f(f(nr * nr) + f(ni * ni)))
        a, b = f(nr * nm), f(ni * nm)
        out[k] = complex(a, b)
    return out


def bf16_round(x):
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def split2(x):
    hi = bf16_round(x)
    mid = bf16_round((np.asarray(x, np.float32) - hi).astype(np.float32))
    return hi.astype(np.float64), mid.astype(np.float64)


def main():
    rng = np.random.default_rng(1)
    fs, freq = 61440000.0, 12345678.0
    ang = 2 * np.pi * freq / fs
    c, d = np.float32(np.cos(ang)), np.float32(np.sin(ang))
    N = 32 * 400
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)) * 0.1
    x = x.astype(np.complex64).astype(np.complex128)
    S = nco_table(c, d, N + 1)
    osc = S[1:N + 1]                       # sample n (table index n) is mixed with S(n + 1)
    mixed = osc * x
    # the cascade, float64, zero history, LTI
    y = mixed
    for s in range(5):
        full = np.convolve(y, h)           # full[k] = sum_t h[t] y[k - t]; y_s[j] = sum_t h[t] y[2j - 10 + t] = full[2j] (h symmetric)
        y = full[0:len(y):2]
    z_ref = y                              # z_ref[m], m = 0..N/32-1
    g = composite()
    assert len(g) == 311
    theta = np.arctan2(float(d), float(c))
    m0 = 64
    ms = np.arange(m0, N // 32)
    z = np.empty(len(ms), dtype=np.complex128)
    zb = np.empty(len(ms), dtype=np.complex128)
    # padded window of 320 samples starting at 32m - 312 (8-sample aligned): tap index = t' - 2
    tp = np.arange(320)
    gp = np.zeros(320)
    gp[2:313] = g
    G = gp * np.exp(1j * theta * tp)
    Gr_h, Gr_m = split2(G.real)
    Gi_h, Gi_m = split2(G.imag)
    xr_h, xr_m = split2(x.real)
    xi_h, xi_m = split2(x.imag)
    for k, m in enumerate(ms):
        ns = 32 * m - 312
        idx1 = ns + 1                      # osc(ns) = S(ns + 1)
        cc = idx1 >> 8
        r = idx1 - 256 * cc
        fac = S[256 * cc] * np.exp(1j * theta * r)
        w = x[ns:ns + 320]
        z[k] = fac * np.sum(G * w)
        sl = slice(ns, ns + 320)
        # three bf16 products: hi*hi + mid*hi + hi*mid, fp32-ish accumulate (float64 here)
        def prod(ar, ai, br, bi):
            return np.sum(ar * br - ai * bi) + 1j * np.sum(ar * bi + ai * br)
        acc = prod(Gr_h, Gi_h, xr_h[sl], xi_h[sl]) + prod(Gr_h, Gi_h, xr_m[sl], xi_m[sl]) + prod(Gr_m, Gi_m, xr_h[sl], xi_h[sl])
        zb[k] = fac * acc
    ref = z_ref[ms]
    def snr(a, b):
        return 10 * np.log10(np.sum(np.abs(b) ** 2) / np.sum(np.abs(a - b) ** 2))
    print("rms(ref) %.4g" % np.sqrt(np.mean(np.abs(ref) ** 2)))
    print("factored formula, float64 : max|err| %.3g  SNR %.1f dB" % (np.max(np.abs(z - ref)), snr(z, ref)))
    print("bf16 hi+mid, 3 products   : max|err| %.3g  SNR %.1f dB" % (np.max(np.abs(zb - ref)), snr(zb, ref)))


if __name__ == "__main__":
    main()
