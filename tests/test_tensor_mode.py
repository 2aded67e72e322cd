"""GPU suite (-m gpu): AERODDC_MODE_TENSOR - the NCO mix and the first five half-band stages as one complex GEMM on the
tcgen05 tensor cores, the remaining stages fused into its epilogue (aero-cli_b200/csrc/tc_kernels.cuh) - against the
oracle (the reference's vfo.cpp chain restated; pinned on the compiled reference in test_oracle_golden.py).

It is a tolerance mode. Bound, from BASELINE.json's north_star and asserted here UNCONDITIONALLY on signals of normal
level: max |err| <= 1e-4 of full scale (3.27 int16 LSB) and error SNR >= 80 dB on the int16 payloads; the float
stage-D stream is held to >= 100 dB."""
import numpy as np
import pytest

from oracle_bind import FMT_CF32, FMT_CU8, Oracle
from case_util import parity_metrics

pytestmark = pytest.mark.gpu


def _aeroddc():
    import aeroddc

    return aeroddc


def _signal(fs, blk, k, tones, seed):
    """cf32 block k: wideband noise plus one tone 700 Hz above the centre of each listed VFO (mixer = centre - f, so the tone
    sits at -mixer + 700 Hz): every checked channel carries an in-band signal of normal level."""
    r = np.random.default_rng(seed + k)
    x = (r.standard_normal(2 * blk) * 0.05).astype(np.float32)
    n = np.arange(blk, dtype=np.float64) + float(k) * blk
    for f, amp in tones:
        ph = 2 * np.pi * (((-f + 700.0) / fs * n) % 1.0)
        x[0::2] += (amp * np.cos(ph)).astype(np.float32)
        x[1::2] += (amp * np.sin(ph)).astype(np.float32)
    return x


def _bank(a, fs, blk, fmt, vfos, mode, dcc=False):
    b = a.Bank(fs, blk, fmt, 0)
    for i, (f, D, L, g) in enumerate(vfos):
        b.add_vfo(f, D, L, 0, g, 1, 1, 1, "N%04d" % i)
    b.set_mode(mode)
    if dcc:
        b.set_dc_correction(True)
    b.finalize()
    return b


def _check(bank, oracles, picks, vfos, blk, x, k, floor_rms=1000.0):
    for i, o in zip(picks, oracles):
        want = np.frombuffer(o.process(x), np.int16)
        got = np.frombuffer(bank.output(i)[0], np.int16)
        assert got.size == want.size
        maxerr, snr = parity_metrics(got, want)
        rms = float(np.sqrt((want.astype(np.float64) ** 2).mean()))
        assert rms >= floor_rms, (i, k, rms)                 # the test signal is of normal level by construction
        assert maxerr <= 1e-4, (i, k, maxerr)
        assert snr >= 80.0, (i, k, snr, rms)
        D = vfos[i][1]
        sg = bank.stage_d(i, blk >> D).astype(np.float64)
        so = o.stage(D).astype(np.float64)
        snr_f = 10 * np.log10((so * so).sum() / max(((sg - so) ** 2).sum(), 1e-300))
        assert snr_f >= 100.0, (i, k, snr_f)


def test_tensor_mode_wideband_mixed_stage_counts_across_the_table_restart():
    """61.44 MS/s cf32, blocks of Fs/4 (the benchmark's geometry), 70 VFOs = one full 64-VFO tile and a ragged one, stage counts
    8 / 7 / 6 / 5 mixed inside the tiles (the fused epilogue runs 3 / 2 / 1 / 0 stages per lane), five blocks: four block
    boundaries (shifted half-band history of all eight stages) and the oscillator-table restart at sample Fs (block 4)."""
    a = _aeroddc()
    fs, blk = 61440000, 15360000
    rng = np.random.default_rng(61)
    freqs = rng.integers(int(-0.45 * fs), int(0.45 * fs), 70).astype(np.float64)
    Ds = [8, 8, 7, 6, 5]
    vfos = [(float(freqs[i]), Ds[i % 5], 5, 0.5) for i in range(70)]
    picks = [0, 2, 3, 4, 63, 64, 69]
    bank = _bank(a, fs, blk, FMT_CF32, vfos, a.MODE_TENSOR)
    oracles = [Oracle(fs, blk, vfos[i][1], 5, vfos[i][0], 0.5, 0) for i in picks]
    tones = [(vfos[i][0], 0.2) for i in picks]
    for k in range(5):
        x = _signal(fs, blk, k, tones, 610)
        bank.process(x)
        _check(bank, oracles, picks, vfos, blk, x, k)
    bank.close()


@pytest.mark.parametrize("gather", ["default", "0"])
def test_tensor_mode_table_restart_inside_a_block_and_sliced_input(gather, monkeypatch):
    """1.536 MS/s with blocks of 393216 samples: the oscillator table (1 536 000 entries) restarts in the middle of block 3,
    beyond the block head - the FP32 kernel takes over a second zone there. The raw block is handed over in three device
    slices (on a node: one per GPU), as bench.py's peer exchange does: by default the bank's copy engines gather them into
    its local input buffer first; with AERODDC_GATHER=0 the kernels' own loads read every slice in place."""
    a = _aeroddc()
    if gather == "0":
        monkeypatch.setenv("AERODDC_GATHER", "0")
    else:
        monkeypatch.delenv("AERODDC_GATHER", raising=False)
    fs, blk = 1536000, 393216
    rng = np.random.default_rng(15)
    freqs = rng.integers(int(-0.45 * fs), int(0.45 * fs), 40).astype(np.float64)
    Ds = [8, 7, 6, 8, 5]
    vfos = [(float(freqs[i]), Ds[i % 5], 0, 0.5) for i in range(40)]
    picks = [0, 1, 2, 4, 38, 39]
    bank = _bank(a, fs, blk, FMT_CF32, vfos, a.MODE_TENSOR)
    oracles = [Oracle(fs, blk, vfos[i][1], 0, vfos[i][0], 0.5, 0) for i in picks]
    tones = [(vfos[i][0], 0.15) for i in picks]
    slice_len = 131072
    ptrs = [a.dev_alloc(0, slice_len * 8) for _ in range(3)]
    for k in range(6):
        x = _signal(fs, blk, k, tones, 150)
        for s, p in enumerate(ptrs):
            a.dev_upload(0, p, x[2 * s * slice_len:2 * (s + 1) * slice_len])
        bank.submit_device_sliced(ptrs, slice_len)
        bank.wait()
        _check(bank, oracles, picks, vfos, blk, x, k)
    for p in ptrs:
        a.dev_free(0, p)
    bank.close()


def test_tensor_mode_differs_from_exact_only_within_tolerance_and_pipelines():
    """Two blocks in flight (submit / submit / wait / wait, the way bench.py and the Publisher drive the bank): the stage-D rows
    are single-buffered between the tensor kernel and the tail kernel, so block k+1's kernel must wait for block k's tail.
    The payloads equal those of one-at-a-time processing bit for bit, and differ from the exact mode's by at most 1e-4 FS."""
    a = _aeroddc()
    fs, blk = 2400000, 491520
    rng = np.random.default_rng(24)
    freqs = rng.integers(int(-0.45 * fs), int(0.45 * fs), 130).astype(np.float64)
    vfos = [(float(freqs[i]), 8 if i % 2 else 6, 0, 0.5) for i in range(130)]
    tones = [(vfos[i][0], 0.03) for i in range(0, 130, 13)]
    blocks = [_signal(fs, blk, k, tones, 240) for k in range(4)]
    serial = _bank(a, fs, blk, FMT_CF32, vfos, a.MODE_TENSOR)
    piped = _bank(a, fs, blk, FMT_CF32, vfos, a.MODE_TENSOR)
    exact = _bank(a, fs, blk, FMT_CF32, vfos, a.MODE_EXACT)
    want, ref = [], []
    for x in blocks:
        serial.process(x)
        want.append([serial.output(i)[0] for i in range(130)])
        exact.process(x)
        ref.append([exact.output(i)[0] for i in range(130)])
    got = []
    piped.submit(blocks[0])
    for k in range(1, 4):
        piped.submit(blocks[k])
        piped.wait()
        got.append([piped.output(i)[0] for i in range(130)])
    piped.wait()
    got.append([piped.output(i)[0] for i in range(130)])
    assert got == want
    differs = 0
    for k in range(4):
        for i in range(130):
            g, r = np.frombuffer(want[k][i], np.int16), np.frombuffer(ref[k][i], np.int16)
            maxerr, _ = parity_metrics(g, r)
            assert maxerr <= 1e-4, (k, i, maxerr)
            differs += int(np.any(g != r))
    assert differs > 0                                   # the tensor path really ran
    for b in (serial, piped, exact):
        b.close()


def test_tensor_mode_with_dc_correction_on_cu8_input():
    """With DC correction on, every input format reaches the VFO kernels as the corrected cf32 block, so the tensor mode applies
    to a cu8 stream as well; against the oracle fed with the same corrected samples."""
    from oracle_bind import dc_correct, synth_raw, unpack
    a = _aeroddc()
    fs, blk = 2400000, 491520
    vfos = [(123456.0, 8, 0, 0.5), (-654321.0, 7, 0, 0.5), (800000.0, 6, 0, 0.5)]
    bank = _bank(a, fs, blk, FMT_CU8, vfos, a.MODE_TENSOR, dcc=True)
    oracles = [Oracle(fs, blk, D, 0, f, g, 0) for (f, D, L, g) in vfos]
    state = np.zeros(2, np.float32)
    for k in range(3):
        raw = synth_raw(FMT_CU8, k * blk, blk, seed=31, amp=0.3)
        # an in-band tone per VFO on top of the synthetic capture
        xf = unpack(FMT_CU8, raw).astype(np.float32)
        n = np.arange(blk, dtype=np.float64) + float(k) * blk
        for (f, D, L, g) in vfos:
            ph = 2 * np.pi * (((-f + 700.0) / fs * n) % 1.0)
            xf[0::2] += (0.15 * np.cos(ph)).astype(np.float32)
            xf[1::2] += (0.15 * np.sin(ph)).astype(np.float32)
        raw = np.clip(np.round(xf * 128.0 + 127.4), 0, 255).astype(np.uint8)
        bank.process(raw)
        x = dc_correct(np.array(unpack(FMT_CU8, raw), np.float32), state)
        _check(bank, oracles, [0, 1, 2], vfos, blk, x, k)
    bank.close()


def test_tensor_mode_must_be_chosen_before_finalize_and_falls_back_where_it_does_not_apply():
    a = _aeroddc()
    late = a.Bank(2400000, 491520, a.CF32, 0)
    late.add_vfo(1000.0, 8, 0, 0, 0.5, 1, 1, 1, "LATE0")
    late.finalize()
    with pytest.raises(a.AeroDdcError):
        late.set_mode(a.MODE_TENSOR)
    late.close()
    # cu8 without DC correction, and a VFO with fewer than six stages: the bank runs as in AERODDC_MODE_FAST
    fs, blk = 288000, 57600
    vfos = [(20000.0, 4, 0, 0.5), (-31000.0, 2, 0, 0.5)]
    bank = _bank(a, fs, blk, FMT_CU8, vfos, a.MODE_TENSOR)
    from oracle_bind import synth_raw, unpack
    oracles = [Oracle(fs, blk, D, 0, f, g, 0) for (f, D, L, g) in vfos]
    for k in range(3):
        raw = synth_raw(FMT_CU8, k * blk, blk, seed=7, amp=0.8)
        bank.process(raw)
        x = unpack(FMT_CU8, raw)
        for i, o in enumerate(oracles):
            maxerr, _ = parity_metrics(np.frombuffer(bank.output(i)[0], np.int16), np.frombuffer(o.process(x), np.int16))
            assert maxerr <= 1e-4
    bank.close()


def test_tensor_mode_beside_a_main_vfo_with_sub_vfos():
    """One bank holding a main VFO that feeds sub-VFOs (publisher.cpp:118-219; such a bank runs its kernels in order on one
    stream) AND flat VFOs with seven / eight stages, which take the tensor path: the sub-VFOs (not eligible: they mix the
    main VFO's output) stay within tolerance on the FP32 kernels, the flat ones on the tensor kernel."""
    a = _aeroddc()
    fs, blk = 1536000, 393216
    bank = a.Bank(fs, blk, FMT_CF32, 0)
    fm = -200000.0
    main = bank.add_vfo(fm, 3, 0, 0, 0.01, 0, 1, 1, "MAIN0")
    subs = [(15000.0, 2, 0, 0.5), (-22000.0, 1, 0, 0.5)]
    flats = [(300000.0, 8, 0, 0.5), (-450000.0, 7, 0, 0.5), (123000.0, 6, 0, 0.5)]
    sub_ids = [bank.add_vfo(f, D, L, 0, g, 1, 1, 1, "SUB%02d" % i, parent=main) for i, (f, D, L, g) in enumerate(subs)]
    flat_ids = [bank.add_vfo(f, D, L, 0, g, 1, 1, 1, "FLT%02d" % i) for i, (f, D, L, g) in enumerate(flats)]
    bank.set_mode(a.MODE_TENSOR)
    bank.finalize()
    o_main = Oracle(fs, blk, 3, 0, fm, 0.01, 0, 0, 1, 1)
    o_subs = [Oracle(fs >> 3, blk >> 3, D, L, f, g, 0) for (f, D, L, g) in subs]
    o_flats = [Oracle(fs, blk, D, L, f, g, 0) for (f, D, L, g) in flats]
    # tones: inside every flat channel, and inside every sub channel (relative to the main VFO's centre)
    tones = [(f, 0.15) for (f, D, L, g) in flats] + [(fm + f, 0.15) for (f, D, L, g) in subs]
    for k in range(5):
        x = _signal(fs, blk, k, tones, 330)
        bank.process(x)
        o_main.process(x)
        mid = o_main.stage(3)
        for i, o in zip(sub_ids, o_subs):
            maxerr, snr = parity_metrics(np.frombuffer(bank.output(i)[0], np.int16), np.frombuffer(o.process(mid), np.int16))
            assert maxerr <= 1e-4 and snr >= 80.0, (i, k, maxerr, snr)
        for i, o, (f, D, L, g) in zip(flat_ids, o_flats, flats):
            want = np.frombuffer(o.process(x), np.int16)
            maxerr, snr = parity_metrics(np.frombuffer(bank.output(i)[0], np.int16), want)
            assert maxerr <= 1e-4 and snr >= 80.0, (i, k, maxerr, snr)
            sg = bank.stage_d(i, blk >> D).astype(np.float64)
            so = o.stage(D).astype(np.float64)
            assert 10 * np.log10((so * so).sum() / max(((sg - so) ** 2).sum(), 1e-300)) >= 100.0, (i, k)
    bank.close()


def test_fleet_in_tensor_mode_equals_the_single_bank():
    """The native multi-GPU layer in tensor mode (two GPUs: the raw block lies in one slice per GPU and every bank's copy
    engines gather both over NVLink into its local input buffer before its kernels start): every VFO's payload equals the single-bank tensor run bit for bit (same kernel,
    same arithmetic, only the block's location differs). Degenerates to one GPU where there is only one."""
    import torch

    a = _aeroddc()
    ndev = min(2, torch.cuda.device_count())
    fs, blk = 2400000, 491520
    rng = np.random.default_rng(77)
    freqs = rng.integers(int(-0.45 * fs), int(0.45 * fs), 70).astype(np.float64)
    vfos = [(float(freqs[i]), [8, 7, 6][i % 3], 0, 0.5) for i in range(70)]
    fleet = a.Fleet(fs, blk, a.CF32, tuple(range(ndev)))
    bank = a.Bank(fs, blk, a.CF32, 0)
    for i, (f, D, L, g) in enumerate(vfos):
        fleet.add_vfo(f, D, L, 0, g, 1, 1, 1, "Q%04d" % i)
        bank.add_vfo(f, D, L, 0, g, 1, 1, 1, "Q%04d" % i)
    fleet.set_mode(a.MODE_TENSOR)
    bank.set_mode(a.MODE_TENSOR)
    fleet.finalize()
    bank.finalize()
    tones = [(vfos[i][0], 0.05) for i in range(0, 70, 7)]
    for k in range(3):
        x = _signal(fs, blk, k, tones, 770)
        fleet.process(x)
        bank.process(x)
        for i in range(70):
            assert fleet.output(i) == bank.output(i), (i, k)
    fleet.close()
    bank.close()
