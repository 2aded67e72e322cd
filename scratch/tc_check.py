"""Tensor mode against the exact mode on the GPU (scratch): stage-D error, payload error, kernel time.
usage: python scratch/tc_check.py [n_vfos] [fs] [block] [blocks] [late] [tensor-only]"""
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, 'aero-cli_b200')
import numpy as np, aeroddc

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fs = int(sys.argv[2]) if len(sys.argv) > 2 else 61440000
blk = int(sys.argv[3]) if len(sys.argv) > 3 else fs // 4
nblocks = int(sys.argv[4]) if len(sys.argv) > 4 else 3
D, G = 8, 0.05
L = int(sys.argv[5]) if len(sys.argv) > 5 else 5
only_tensor = len(sys.argv) > 6
rng = np.random.default_rng(5)
freqs = rng.integers(int(-0.45 * fs), int(0.45 * fs), nv).astype(np.float64)

def block(k):
    r = np.random.default_rng(100 + k)
    x = (r.standard_normal(2 * blk) * 0.07).astype(np.float32)
    n = np.arange(blk, dtype=np.float64) + k * blk
    for v in range(0, nv, max(1, nv // 16)):           # a carrier 700 Hz off some VFO centres: in-band tones of normal level
        ph = 2 * np.pi * (((-freqs[v] + 700.0) / fs * n) % 1.0)
        x[0::2] += (0.02 * np.cos(ph)).astype(np.float32)
        x[1::2] += (0.02 * np.sin(ph)).astype(np.float32)
    return x

def run(mode):
    b = aeroddc.Bank(fs, blk, aeroddc.CF32, 0)
    for v in range(nv):
        b.add_vfo(float(freqs[v]), D, L, 0, G, 1, 1, 1, "T%04d" % v)
    b.set_mode(mode)
    b.finalize()
    outs, stages, ms = [], [], []
    for k in range(nblocks):
        b.process(block(k))
        ms.append(b.last_main_ms())
        outs.append([np.frombuffer(b.output(v)[0], np.int16).copy() for v in range(nv)])
        stages.append([b.stage_d(v, blk >> D).copy() for v in range(0, nv, max(1, nv // 32))])
    b.close()
    return outs, stages, ms

if only_tensor:
    got = run(aeroddc.MODE_TENSOR)
    print("tensor only: main ms per block", ["%.3f" % m for m in got[2]])
    sys.exit(0)
ref = run(aeroddc.MODE_EXACT)
for name, mode in (("fast", aeroddc.MODE_FAST), ("tensor", aeroddc.MODE_TENSOR)):
    t0 = time.time()
    got = run(mode)
    for k in range(nblocks):
        e = np.concatenate([g.astype(np.float64) - r.astype(np.float64) for g, r in zip(got[0][k], ref[0][k])])
        s = np.concatenate([r.astype(np.float64) for r in ref[0][k]])
        snr = 10 * np.log10((s ** 2).sum() / max((e ** 2).sum(), 1e-30))
        se = np.concatenate([g - r for g, r in zip(got[1][k], ref[1][k])]).astype(np.float64)
        ss = np.concatenate(ref[1][k]).astype(np.float64)
        ssnr = 10 * np.log10((ss ** 2).sum() / max((se ** 2).sum(), 1e-30))
        worst_v = int(np.argmax([np.abs(g.astype(np.int32) - r.astype(np.int32)).max() for g, r in zip(got[0][k], ref[0][k])]))
        print("%-6s block %d: main %.3f ms (exact %.3f) | int16 max|err| %d LSB (vfo %d) SNR %.1f dB rms %.0f | stage-D max|err| %.3g SNR %.1f dB"
              % (name, k, got[2][k], ref[2][k], np.abs(e).max(), worst_v, snr, np.sqrt((s ** 2).mean()), np.abs(se).max(), ssnr))
        if name == "tensor" and np.abs(se).max() > 1e-3:
            g0, r0 = got[1][k][0], ref[1][k][0]
            d = np.abs(g0 - r0)
            bad = np.nonzero(d > 1e-3)[0]
            print("   vfo 0: %d of %d stage-D floats off; first at %s; got %s want %s" % (len(bad), len(d), bad[:8], g0[bad[:4]], r0[bad[:4]]))
