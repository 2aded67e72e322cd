"""GPU suite: the native multi-GPU layer (aeroddc_fleet_*): VFO sharding over the node's GPUs from one host thread, the raw
block either spread over the GPUs in slices that every kernel reads in place through peer memory (default) or
broadcast with NCCL (AERODDC_EXCHANGE=nccl). On a one-GPU box it checks the degenerate fleet; with two or
more GPUs (gpurun --gpus N) it checks that every VFO's bytes equal the single-bank result and the oracle."""
import numpy as np
import pytest

from oracle_bind import FMT_CF32, FMT_CU8, Oracle, synth_raw, unpack

pytestmark = pytest.mark.gpu


def _ndev():
    import torch

    return torch.cuda.device_count()


def _descs(fs, n, seed):
    rng = np.random.default_rng(seed)
    return [dict(mixer=float(rng.integers(int(-0.4 * fs), int(0.4 * fs))), D=[5, 6, 7][i % 3], gain=float(rng.uniform(0.05, 0.4))) for i in range(n)]


@pytest.mark.parametrize("ndev,exchange", [(1, "single"), (2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer"), (8, "nccl")])
def test_fleet_matches_single_bank_and_oracle(ndev, exchange, monkeypatch):
    import aeroddc

    if _ndev() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    if exchange == "nccl":
        monkeypatch.setenv("AERODDC_EXCHANGE", "nccl")
    fs, blk = 1536000, 384000
    descs = _descs(fs, 37, 5)
    fleet = aeroddc.Fleet(fs, blk, aeroddc.CU8, tuple(range(ndev)))
    bank = aeroddc.Bank(fs, blk, aeroddc.CU8, 0)
    for i, d in enumerate(descs):
        assert fleet.add_vfo(d["mixer"], d["D"], 0, 0, d["gain"], 1, 1, 1, "F%04d" % i) == i
        bank.add_vfo(d["mixer"], d["D"], 0, 0, d["gain"], 1, 1, 1, "F%04d" % i)
    fleet.finalize()
    bank.finalize()
    assert fleet.num_devices == ndev and fleet.exchange == exchange
    assert sorted(set(fleet.device_of(i) for i in range(len(descs)))) == list(range(ndev))
    picks = [0, 1, 17, 36]
    oracles = [Oracle(fs, blk, descs[i]["D"], 0, descs[i]["mixer"], descs[i]["gain"]) for i in picks]
    raws = [synth_raw(FMT_CU8, b * blk, blk, seed=6, amp=0.8) for b in range(4)]
    # pipelined: two blocks in flight
    fleet.submit(raws[0])
    for b in range(4):
        if b + 1 < 4:
            fleet.submit(raws[b + 1])
        fleet.wait()
        bank.process(raws[b])
        x = unpack(FMT_CU8, raws[b])
        for i in range(len(descs)):
            assert fleet.output(i) == bank.output(i), (i, b)
        for j, i in enumerate(picks):
            assert fleet.output(i)[0] == oracles[j].process(x), (i, b)
    fleet.close()
    bank.close()


@pytest.mark.parametrize("ndev", [1, 2])
def test_fleet_keeps_sub_vfos_with_their_main_vfo(ndev):
    import aeroddc

    if _ndev() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    fs, blk = 1920000, 480000
    fleet = aeroddc.Fleet(fs, blk, aeroddc.CF32, tuple(range(ndev)))
    m0 = fleet.add_vfo(-250000.0, 3, 0, 0, 0.01, 0, 1, 1, "MAIN0")
    m1 = fleet.add_vfo(400000.0, 3, 0, 0, 0.01, 0, 1, 1, "MAIN1")
    s0 = fleet.add_vfo(20000.0, 0, 5, 0, 0.5, 1, 1, 1, "SUB00", parent=m0)
    s1 = fleet.add_vfo(-30000.0, 1, 5, 0, 0.5, 1, 1, 1, "SUB01", parent=m1)
    fleet.finalize()
    assert fleet.device_of(s0) == fleet.device_of(m0) and fleet.device_of(s1) == fleet.device_of(m1)
    if ndev > 1:
        assert fleet.device_of(m0) != fleet.device_of(m1)
    mains = [Oracle(fs, blk, 3, 0, -250000.0, 0.01, 0, 0, 1, 1), Oracle(fs, blk, 3, 0, 400000.0, 0.01, 0, 0, 1, 1)]
    subs = [Oracle(fs >> 3, blk >> 3, 0, 5, 20000.0, 0.5), Oracle(fs >> 3, blk >> 3, 1, 5, -30000.0, 0.5)]
    for b in range(3):
        x = synth_raw(FMT_CF32, b * blk, blk, seed=2, amp=0.7)
        fleet.process(x)
        for m, s_, idx in ((mains[0], subs[0], s0), (mains[1], subs[1], s1)):
            m.process(x)
            assert fleet.output(idx)[0] == s_.process(m.stage(3)), (idx, b)
    fleet.close()


def test_bank_reads_its_block_from_a_peer_gpu():
    """aeroddc_bank_submit_device with an address in ANOTHER GPU's HBM: the kernel's TMA tile loads cross NVLink.
    Bytes must equal those of the same bank fed from local memory."""
    import aeroddc

    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    fs, blk = 1536000, 384000
    descs = _descs(fs, 9, 11)
    banks = []
    for dev in (1, 1):
        b = aeroddc.Bank(fs, blk, aeroddc.CF32, dev)
        for i, d in enumerate(descs):
            b.add_vfo(d["mixer"], d["D"], 0, 0, d["gain"], 1, 1, 1, "P%04d" % i)
        b.finalize()
        banks.append(b)
    aeroddc.enable_peer(1, 0)
    remote = aeroddc.dev_alloc(0, blk * 8)          # lives on GPU 0, read by the bank on GPU 1
    for k in range(3):
        x = synth_raw(FMT_CF32, k * blk, blk, seed=4, amp=0.8)
        aeroddc.dev_upload(0, remote, x)
        banks[0].submit_device(remote)
        banks[0].wait()
        banks[1].process(x)
        for i in range(len(descs)):
            assert banks[0].output(i) == banks[1].output(i), (i, k)
    aeroddc.dev_free(0, remote)
    for b in banks:
        b.close()
