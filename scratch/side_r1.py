"""Round-1 library (scratch/r1_src, built from commit 097774d) on bench.py's side configurations A/B/C, GPU part only:
the 'before' numbers for VERDICT item 7 (config B >= 5x its round-1 Gsps). Usage: python scratch/side_r1.py [r1|now]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aero-cli_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import aeroddc
which = sys.argv[1] if len(sys.argv) > 1 else "r1"
if which == "r1":
    aeroddc.LIB_PATH = os.path.join(ROOT, "scratch", "r1_src", "libaeroddc_r1.so")
    aeroddc.ABI_SYMBOLS[:] = []
    # the r1 library lacks the newer entry points: bind only what exists
    class Lax(ctypes.CDLL):
        def __getattr__(self, name):
            try:
                return super().__getattr__(name)
            except AttributeError:
                class Dummy:  # accepts argtypes/restype assignment
                    pass
                d = Dummy(); object.__setattr__(self, name, d); return d
    _orig = ctypes.CDLL
    ctypes.CDLL = Lax
    aeroddc.lib()
    ctypes.CDLL = _orig
import bench
import oracle_bind as ob
dev = torch.device("cuda", 0)
out = {}
def run(name, bank, raws, steps, nch, blk):
    ts = [torch.from_numpy(a).to(dev) for a in raws]; torch.cuda.synchronize()
    ms, mm, launches = bench._timed_device_loop(bank, [t.data_ptr() for t in ts], steps, 5)
    out[name] = {"gsps": nch * blk * steps / (ms * 1e-3) / 1e9, "ms_per_step": ms / steps, "launches": launches}
    bank.close()
fs, blk = 2400000, 480000
b = aeroddc.Bank(fs, blk, aeroddc.CU8, 0); b.add_vfo(123456.0, 5, 0, 0, 0.05, 1, 1, 1, "A0000"); b.finalize()
run("A", b, [ob.synth_raw(ob.FMT_CU8, k * blk, blk, seed=11, amp=0.5) for k in range(2)], 200, 1, blk)
fs, blk = 1536000, 384000
rng = np.random.default_rng(54); ds = [7] * 20 + [6] * 8 + [5] * 2
fr = rng.integers(int(-0.45 * fs), int(0.45 * fs), 30).astype(np.float64); gains = rng.integers(5, 11, 30) / 100.0
b = aeroddc.Bank(fs, blk, aeroddc.CU8, 0); m = b.add_vfo(0.0, 0, 0, 0, 0.01, 0, 1, 1, "MAIN0")
for i in range(30): b.add_vfo(float(fr[i]), ds[i], 0, 0, float(gains[i]), 1, 1, 1, "B%04d" % i, parent=m)
b.finalize()
run("B", b, [ob.synth_raw(ob.FMT_CU8, k * blk, blk, seed=12, amp=0.5) for k in range(2)], 200, 30, blk)
freqs = bench.vfo_freqs(bench.N_VFOS)[:256]
b = aeroddc.Bank(bench.FS, bench.BLOCK, aeroddc.CF32, 0)
for v in range(256): b.add_vfo(float(freqs[v]), 8, 5, 0, 0.05, 1, 1, 1, "C%04d" % v)
b.finalize()
run("C", b, [bench.synth_block(1), bench.synth_block(2)], 20, 256, bench.BLOCK)
print(json.dumps({"lib": which, **out}))
