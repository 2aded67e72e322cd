"""Time of the DC-correction path at 61.44 MS/s (scratch; the bench line `dc_correction` is the reported number)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "aero-cli_b200"))
import numpy as np, torch, aeroddc as a
FS, BLK, NV = 61440000, 15360000, int(sys.argv[1]) if len(sys.argv) > 1 else 128
bank = a.Bank(FS, BLK, a.CF32, 0)
rng = np.random.default_rng(1)
for v in range(NV):
    bank.add_vfo(float(rng.integers(-27000000, 27000000)), 8, 5, 0, 0.05, 1, 1, 1, "V%04d" % v)
bank.set_dc_correction(True)
bank.finalize()
x = torch.randn(2 * BLK, device="cuda") * 0.1 + 0.05
bufs = [x, x.clone()]
torch.cuda.synchronize()
def run(n):
    infl = 0
    for k in range(n):
        if infl == 2:
            bank.wait(); infl -= 1
        bank.submit_device(bufs[k & 1].data_ptr(), None); infl += 1
    while infl:
        bank.wait(); infl -= 1
run(2)
torch.cuda.synchronize()
bank.stopwatch_start(False)
run(6)
ms = bank.stopwatch_stop() / 6
print("dcc on: %.2f ms per block, %.1f Gsps (%d VFOs), %.2fx real time" % (ms, NV * BLK / ms / 1e6, NV, 250.0 / ms))
bank.close()
