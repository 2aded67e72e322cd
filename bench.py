#!/usr/bin/env python
"""bench.py - throughput of the multi-VFO DDC bank (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--vfos V]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on; it fits one GPU): V = 1024 VFOs
at integer-Hz offsets within +-0.45 Fs over ONE synthetic 61.44 MS/s cf32 stream, D = 8 half-band
stages + late /5 FIR -> 48 kHz USB-demodulated int16 per VFO. A step is one block of
B = Fs/4 = 15 360 000 complex samples (the reference's block contract) through the whole chain for
all V VFOs. With N GPUs the VFOs are sharded v mod N (strong scaling: total work fixed), the raw
block is broadcast from rank 0 with NCCL, and every rank returns its own payloads to its host.

metric / value : aggregate VFO-channel input samples/s = V * B * K / device time, Gsps, raw blocks
                 resident in HBM (rank 0's HBM for N > 1; the NCCL broadcast is inside the region),
                 two alternating 123 MB input buffers (larger than the 126 MB L2 together).
e2e            : the same through the host-facing C-ABI calls (aeroddc_bank_submit/wait): pinned
                 host blocks, H2D copy of every block and D2H of every payload inside the region.
roofline       : the dominant kernel (ddc_main_kernel: unpack + NCO mix + half-band cascade) against
                 the FP32 FFMA issue peak measured in the same run (aeroddc_measure_fp32_peak).
cpu_baseline   : the reference's own vfo::process chain (oracle/_ref, or the oracle port) on the
                 host cores, one VFO per thread, bounded sample.
--impl reference prints the CPU figure as the headline line instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "aero-cli_b200"))

FS = 61440000
BLOCK = FS // 4
DECIM, LATE, GAIN = 8, 5, 0.05
N_VFOS = 1024
# algorithmic flops per VFO-input-sample (mul and add each 1 flop):
# mix 6 + half-band cascade 20*(1 - 2^-D); tail (late FIR 4*49/5 + Hilbert/delay/convert 253/5)/2^D
FLOPS_MAIN = 6.0 + 20.0 * (1.0 - 2.0 ** -DECIM)
FLOPS_TAIL = ((4 * 49) / 5.0 + 253 / 5.0) / 2.0 ** DECIM
FLOPS_TOTAL = FLOPS_MAIN + FLOPS_TAIL


def vfo_freqs(n):
    rng = np.random.default_rng(20261018)
    return rng.integers(int(-0.45 * FS), int(0.45 * FS), n).astype(np.float64)


def synth_block(seed):
    """One block of synthetic cf32: noise plus a few carriers, RMS about 0.1."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(2 * BLOCK) * 0.07).astype(np.float32)
    n = np.arange(BLOCK, dtype=np.float64)
    for f, a in ((1234567.0, 0.05), (-20000123.0, 0.04), (9876543.0, 0.03)):
        ph = 2 * np.pi * ((f / FS * n) % 1.0)
        x[0::2] += (a * np.cos(ph)).astype(np.float32)
        x[1::2] += (a * np.sin(ph)).astype(np.float32)
    return x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load: above 60% of the maximum observed power
        thr = 0.6 * max(pw)
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


class CpuChain:
    """Reference CPU chain, one `vfo` object per host thread (BASELINE.md plan (ii)).
    Objects (and their 491 MB oscillator tables) are built once, outside any timed region."""

    def __init__(self, n_threads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_bind as ob

        self.kind = "reference" if ob.ref_lib() is not None else "port"
        self.n = n_threads
        cls = ob.RefVfo if self.kind == "reference" else ob.Oracle
        self.objs = [cls(FS, BLOCK, DECIM, LATE, float(f), GAIN) for f in vfo_freqs(n_threads)]

    def single(self, x):
        t0 = time.perf_counter()
        self.objs[0].process_repeat(x, 1)
        return BLOCK / (time.perf_counter() - t0) / 1e9

    def step(self, x, blocks_per_vfo=1):
        """All threads process `blocks_per_vfo` blocks; returns wall seconds."""
        ths = [threading.Thread(target=o.process_repeat, args=(x, blocks_per_vfo)) for o in self.objs]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    def sample_text(self, blocks):
        return ("%d VFOs (one per host thread) x %d blocks of %d samples, Fs 61.44 MS/s, D=8, late /5; table build untimed"
                % (self.n, blocks, BLOCK))

    def close(self):
        for o in self.objs:
            o.close()


def cpu_baseline(x):
    cores = os.cpu_count() or 1
    chain = CpuChain(cores)
    single = chain.single(x)
    chain.step(x, 1)            # warm
    blocks = 3
    dt = chain.step(x, blocks)
    chain.close()
    return {"value": cores * blocks * BLOCK / dt / 1e9, "unit": "Gsps", "cores": cores, "kind": chain.kind,
            "single_thread_gsps": single, "sample": chain.sample_text(blocks)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    x = synth_block(1)
    chain = CpuChain(cores)
    single = chain.single(x)
    for _ in range(args.warmup):
        chain.step(x, 1)
    times = [chain.step(x, 1) for _ in range(args.steps)]
    chain.close()
    value = cores * BLOCK * len(times) / sum(times) / 1e9
    line = {
        "impl": "reference", "metric": "aggregate VFO-channel input samples/s", "value": value, "unit": "Gsps",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference vfo::process chain on the host cores, Fs 61.44 MS/s cf32, D=8, late /5 -> 48 kHz (the GPU arm's per-VFO chain); each step = %d VFOs x 1 block of %d samples" % (cores, BLOCK),
                   "n_vfos_per_step": cores, "block_len": BLOCK, "sample_rate": FS},
        "cpu_baseline": {"value": value, "unit": "Gsps", "cores": cores, "kind": chain.kind, "sample": chain.sample_text(1),
                         "single_thread_gsps": single},
        "e2e": {"value": value, "unit": "Gsps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vfos", type=int, default=N_VFOS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast", action="store_true", help="skip the tolerance-mode side measurement")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1, device-resident value: 'peer' = every GPU's kernel pulls the raw tiles straight out of rank 0's HBM "
                         "over NVLink (CUDA IPC mapping, fused with the compute); 'nccl' = ncclBroadcast into a local buffer first")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import aeroddc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the DDC bank has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    # ---- bank: this rank's VFO shard (v mod world) ----
    freqs = vfo_freqs(args.vfos)
    from aeroddc.shard import shard_vfos

    mine = shard_vfos(args.vfos, world, rank)
    bank = aeroddc.Bank(FS, BLOCK, aeroddc.CF32, local_rank)
    for v in mine:
        bank.add_vfo(float(freqs[v]), DECIM, LATE, 0, GAIN, 1, 1, 1, "V%04d" % v)
    t0 = time.perf_counter()
    bank.finalize()
    t_finalize = time.perf_counter() - t0

    # ---- inputs: two distinct blocks; pinned on the host (bank ring), and in rank 0's HBM ----
    host = [bank.host_slot(0), bank.host_slot(1)]
    if rank == 0:
        host[0][:] = synth_block(1)
        host[1][:] = synth_block(2)
    dbuf = [torch.empty(2 * BLOCK, dtype=torch.float32, device=dev) for _ in range(2)]   # grown to 4 for N > 1 below
    src = [torch.empty(2 * BLOCK, dtype=torch.float32, device=dev) for _ in range(2)] if (world > 1 and rank == 0) else None
    if rank == 0:
        for i in range(2):
            (src if world > 1 else dbuf)[i].copy_(torch.from_numpy(host[i]))
    torch.cuda.synchronize()

    # N > 1, --exchange peer: rank 0's two source blocks live in plain cudaMalloc memory exported over CUDA IPC;
    # every other rank maps them and hands the PEER address to its bank, whose TMA tile loads then read rank 0's
    # HBM across NVLink while computing (no broadcast step, no staging buffer).
    peer_ptr = None
    if world > 1 and args.exchange == "peer":
        handles = [None, None]
        own = []
        if rank == 0:
            for i in range(2):
                ptr = aeroddc.dev_alloc(local_rank, host[i].nbytes)
                aeroddc.dev_upload(local_rank, ptr, host[i])
                own.append(ptr)
                handles[i] = aeroddc.ipc_export(local_rank, ptr)
        dist.broadcast_object_list(handles, src=0)
        peer_ptr = own if rank == 0 else [aeroddc.ipc_import(local_rank, h) for h in handles]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the raw block travels rank 0 -> all ranks by NCCL broadcast TWO blocks ahead of its use (four device
    # buffers), so the collective's kernel runs while the previous block's CTAs drain instead of between two blocks.
    nbuf = 4 if world > 1 else 2
    while len(dbuf) < nbuf:
        dbuf.append(torch.empty(2 * BLOCK, dtype=torch.float32, device=dev))
    events = [None] * nbuf

    def prefetch(k, from_host):
        """Start moving block k into this rank's HBM buffer k % nbuf (async)."""
        if world == 1:
            return
        b = k % nbuf
        if rank == 0:
            if from_host:
                dbuf[b].copy_(torch.from_numpy(host[k & 1]), non_blocking=True)    # H2D from the pinned ring
            else:
                dbuf[b].copy_(src[k & 1], non_blocking=True)                        # stays in HBM
        dist.broadcast(dbuf[b], src=0)
        e = torch.cuda.Event()
        e.record()
        events[b] = e

    def run_blocks(n, from_host, on_wait=None, allow_peer=True):
        """n blocks, two in flight. from_host: the host-facing path (pinned host blocks, H2D inside)."""
        peer = peer_ptr is not None and not from_host and allow_peer
        for k in range(min(2, n)):
            if not peer:
                prefetch(k, from_host)
        inflight = 0
        for k in range(n):
            if inflight == 2:
                bank.wait(); inflight -= 1
                if on_wait:
                    on_wait()
            if world == 1:
                if from_host:
                    bank.submit(host[k & 1])
                else:
                    bank.submit_device(dbuf[k & 1].data_ptr(), None)
            elif peer:
                bank.submit_device(peer_ptr[k & 1], None)      # the kernel reads rank 0's HBM directly
            else:
                b = k % nbuf
                bank.submit_device(dbuf[b].data_ptr(), events[b].cuda_event)
                if k + 2 < n:
                    prefetch(k + 2, from_host)       # buffer (k+2) % 4 was last read by block k-2, which has completed
            inflight += 1
        while inflight:
            bank.wait(); inflight -= 1
            if on_wait:
                on_wait()

    def run_device(n):
        run_blocks(n, False)

    def run_e2e(n):
        run_blocks(n, True)

    peak_tflops, probe_clock = aeroddc.measure_fp32_peak(local_rank)

    # ---- timed region 1: inputs resident in HBM ----
    run_device(args.warmup)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_ms = []
    bank.stopwatch_start(False)
    t0 = time.perf_counter()
    run_blocks(args.steps, False, on_wait=lambda: main_ms.append(bank.last_main_ms()))
    dev_ms = bank.stopwatch_stop()
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    kern_ms, launches_per_step = bank.last_timing()

    # ---- timed region 2: end to end through the host-facing API ----
    run_e2e(2)
    sync_all()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    # ---- tolerance mode (AERODDC_MODE_FAST), reported beside the byte-identical headline ----
    fast_ms = None
    if not args.no_fast:
        fbank = aeroddc.Bank(FS, BLOCK, aeroddc.CF32, local_rank)
        for v in mine:
            fbank.add_vfo(float(freqs[v]), DECIM, LATE, 0, GAIN, 1, 1, 1, "V%04d" % v)
        fbank.set_mode(aeroddc.MODE_FAST)
        fbank.finalize()
        main_bank, bank = bank, fbank
        run_blocks(args.warmup, False, allow_peer=False)
        sync_all()
        fast_main = []
        bank.stopwatch_start(False)
        # the tolerance mode consumes raw samples ~1.6x faster; at 8 GPUs seven peers pulling from rank 0 would saturate its
        # NVLink egress, so this side measurement distributes the block with the NCCL broadcast instead
        run_blocks(args.steps, False, on_wait=lambda: fast_main.append(bank.last_main_ms()), allow_peer=False)
        fast_ms = bank.stopwatch_stop()
        sync_all()
        fbank.close()
        bank = main_bank

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, wall_ms, fast_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, wall_ms, fast_ms_r = [float(v) for v in t.tolist()]
        fast_ms = fast_ms_r if fast_ms is not None else None
        n_mine = torch.tensor([len(mine)], dtype=torch.int64, device=dev)
        dist.all_reduce(n_mine)
        assert int(n_mine.item()) == args.vfos

    if rank == 0:
        total = float(args.vfos) * BLOCK * args.steps
        value = total / (dev_ms * 1e-3) / 1e9
        e2e = total / (e2e_ms * 1e-3) / 1e9
        mm = float(np.mean(main_ms))
        per_launch_samples = float(len(mine)) * BLOCK
        achieved = per_launch_samples * FLOPS_MAIN / (mm * 1e-3) / 1e12
        traffic = None
        tf = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("ddc_main_kernel", {}).get(str(len(mine)))
            except Exception:
                traffic = None
        hbm_peak = None
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            hbm_peak = json.load(open(mp)).get("hbm_gbs")
        alg_bytes = BLOCK * 8.0 + len(mine) * (BLOCK >> DECIM) * 8.0   # raw block read once + stage-D stream written
        line = {
            "metric": "aggregate VFO-channel input samples/s", "value": value, "unit": "Gsps",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "%d VFOs x 61.44 MS/s cf32 (BASELINE configs[3]; %d VFOs per GPU, v mod N), D=8 half-band + late /5 FIR -> 48 kHz USB int16; step = one block of %d samples"
                            % (args.vfos, len(mine), BLOCK),
                "n_vfos": args.vfos, "block_len": BLOCK, "sample_rate": FS,
                "l2": "two alternating 123 MB raw blocks (246 MB > 126 MB L2); no explicit flush",
                "parallelism": ("vfo-shard x%d, raw block read from rank 0's HBM over NVLink inside the kernel (peer memory); e2e leg: NCCL broadcast" % world if args.exchange == "peer" else "vfo-shard x%d, NCCL broadcast of the raw block" % world) if world > 1 else "single GPU",
                "realtime_x": value * 1e9 / (args.vfos * FS),
                "finalize_s": t_finalize, "device_mb": bank.device_bytes() / 1e6,
                "flop_per_vfo_sample": FLOPS_TOTAL,
            },
            "e2e": {"value": e2e, "unit": "Gsps", "h2d_bytes_per_step": BLOCK * 8,
                    "d2h_bytes_per_step": int(args.vfos * (BLOCK >> DECIM) // LATE * 2),
                    "note": "aeroddc_bank_submit/wait with pinned host blocks, two blocks in flight; wall clock between device syncs"},
            "gpu_launches": int(launches_per_step * args.steps * (2 if fast_ms is None else 3)),
            "roofline": {
                "bound": "fp32", "kernel": "ddc_main_kernel", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": achieved / peak_tflops, "traffic": traffic,
                "peak_source": "FFMA2 issue probe measured in this run (aeroddc_measure_fp32_peak, FMA = 2 flop); MEASURED_PEAKS.json carries no FP32 figure",
                "launch_ms": mm, "share_of_step": mm / (dev_ms / args.steps),
                "algorithmic_flop_per_vfo_sample": FLOPS_MAIN,
                "issue_bound_note": "the reference's arithmetic is un-fused (1 flop per lane-op) plus a 14 lane-op exact NCO step, so 100%% FP32-pipe use = %.1f%% of the FMA peak" % (100 * FLOPS_MAIN / (2 * (FLOPS_MAIN + 14.0))),
                "hbm": {"achieved_gbs": alg_bytes / (mm * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": (alg_bytes / (mm * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None},
            },
            "clocks": clocks,
            "fast_mode": None if fast_ms is None else {
                "value": total / (fast_ms * 1e-3) / 1e9, "unit": "Gsps", "ms_per_step": fast_ms / args.steps,
                "roofline_frac": float(len(mine)) * BLOCK * FLOPS_MAIN / (float(np.mean(fast_main)) * 1e-3) / 1e12 / peak_tflops,
                "note": "AERODDC_MODE_FAST: fused multiply-adds + rotation-only oscillator between exact checkpoints; NOT bit-identical, "
                        "within max|err| <= 1e-4 FS / SNR >= 80 dB (tests/test_gpu_parity.py::test_fast_mode_within_stated_tolerance); "
                        "the headline value above is the byte-identical mode"},
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(np.array(host[0]))
        print(json.dumps(line))
    bank.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
