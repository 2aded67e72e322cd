"""Config B (ini-style bank: 1 main D=0 + 30 subs) alone, a few blocks: for the ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aero-cli_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import aeroddc, bench
import oracle_bind as ob
dev = torch.device("cuda", 0)
fs, blk = 1536000, 384000
rng = np.random.default_rng(54); ds = [7] * 20 + [6] * 8 + [5] * 2
fr = rng.integers(int(-0.45 * fs), int(0.45 * fs), 30).astype(np.float64); gains = rng.integers(5, 11, 30) / 100.0
b = aeroddc.Bank(fs, blk, aeroddc.CU8, 0); m = b.add_vfo(0.0, 0, 0, 0, 0.01, 0, 1, 1, "MAIN0")
for i in range(30): b.add_vfo(float(fr[i]), ds[i], 0, 0, float(gains[i]), 1, 1, 1, "B%04d" % i, parent=m)
b.finalize()
raws = [ob.synth_raw(ob.FMT_CU8, k * blk, blk, seed=12, amp=0.5) for k in range(2)]
ts = [torch.from_numpy(a).to(dev) for a in raws]; torch.cuda.synchronize()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ms, mm, launches = bench._timed_device_loop(b, [t.data_ptr() for t in ts], steps, 5)
print("config B: %.1f us per step, %.2f Gsps, main kernels %.1f us, %d launches" % (1e3 * ms / steps, 30 * blk * steps / (ms * 1e-3) / 1e9, 1e3 * mm, launches))
b.close()
