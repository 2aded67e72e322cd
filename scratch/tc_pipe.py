"""Pipelined step time of the tensor mode (two blocks in flight, device-resident inputs) - scratch.
usage: [AERODDC_POST_CTAS=deep,tail] python scratch/tc_pipe.py [n_vfos] [mode]"""
import sys
sys.path.insert(0, 'aero-cli_b200')
import numpy as np, torch, aeroddc
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = {"tensor": aeroddc.MODE_TENSOR, "fast": aeroddc.MODE_FAST, "exact": aeroddc.MODE_EXACT}[sys.argv[2] if len(sys.argv) > 2 else "tensor"]
fs, blk = 61440000, 15360000
rng = np.random.default_rng(5)
freqs = rng.integers(int(-0.45 * fs), int(0.45 * fs), nv).astype(np.float64)
b = aeroddc.Bank(fs, blk, aeroddc.CF32, 0)
for v in range(nv):
    b.add_vfo(float(freqs[v]), 8, 5, 0, 0.05, 1, 1, 1, "T%04d" % v)
b.set_mode(mode)
b.finalize()
x = [torch.randn(2 * blk, device="cuda") * 0.1 for _ in range(2)]
torch.cuda.synchronize()
def loop(n):
    infl = 0
    for k in range(n):
        if infl == 2:
            b.wait(); infl -= 1
        b.submit_device(x[k & 1].data_ptr(), None); infl += 1
    while infl:
        b.wait(); infl -= 1
loop(4)
b.stopwatch_start(False)
loop(20)
ms = b.stopwatch_stop()
print("step %.3f ms  main %.3f ms  -> %.0f Gsps" % (ms / 20, b.last_main_ms(), nv * blk * 20 / (ms * 1e-3) / 1e9))
b.close()
