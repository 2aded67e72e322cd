"""DC correction on the benchmark configuration: ms per block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aero-cli_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import aeroddc, bench
dev = torch.device("cuda", 0)
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
freqs = bench.vfo_freqs(bench.N_VFOS)
b = aeroddc.Bank(bench.FS, bench.BLOCK, aeroddc.CF32, 0)
for v in range(nv): b.add_vfo(float(freqs[v]), 8, 5, 0, 0.05, 1, 1, 1, "V%04d" % v)
b.set_dc_correction(True); b.finalize()
ts = [torch.from_numpy(bench.synth_block(1)).to(dev), torch.from_numpy(bench.synth_block(2)).to(dev)]
torch.cuda.synchronize()
cs = bench.ClockSampler(0); cs.start()
ms, mm, l = bench._timed_device_loop(b, [t.data_ptr() for t in ts], 6, 3)
print(cs.stop()); print(sorted(set(x.split(",")[1].strip() for x in cs.lines))[:10])
print("dcc on, %d VFOs: %.2f ms per block (main kernels %.2f ms), %.1f Gsps, %.2f x real time" % (nv, ms / 6, mm, nv * bench.BLOCK * 6 / (ms * 1e-3) / 1e9, 250.0 / (ms / 6)))
b.close()
