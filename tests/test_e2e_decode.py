"""End-to-end check (BASELINE.json configs[4]): the bank's payloads go into
the reference's UNCHANGED decoder and must yield the same decoded ACARS set as the CPU chain's payloads.

  tools/aerol_frames.py        ACARS text -> ISU/SSU signal units -> scrambler, K=7 code, interleaver, UW + header
  tools/synth_iq.py            channel bits -> 600 bit/s MSK carriers inside a 288 kS/s cu8 capture
  aero-publish-b200 --dump     the product (GPU) : capture -> per-topic int16 payloads            [-m gpu]
  tests/tools/oracle_payloads  the CPU chain     : capture -> per-topic int16 payloads
  tests/tools/ref_decode.py    oracle/_ref/libref_decode.so = reference MskDemodulator + SignalHunter + AeroL,
                               compiled unmodified (Qt and libcorrect replaced by shims, see oracle/ref_decode_harness.cpp)

The decoder library is built in this container from /root/reference/decode and travels to the GPU box as a built
file; where it is absent (a checkout without the reference tree) these tests skip and say so.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import aerol_frames as af  # noqa: E402
import ref_decode  # noqa: E402

BIN = os.path.join(ROOT, "aero-cli_b200", "aero-publish-b200")
INI = os.path.join(ROOT, "tests", "data", "e2e_288k.ini")

pytestmark = pytest.mark.skipif(not os.path.exists(ref_decode.LIB),
                                reason="oracle/_ref/libref_decode.so not built (needs /root/reference/decode: make -C oracle refdecode)")

# topic -> (carrier offset from the capture centre: VFO frequency - centre + audio offset, amplitude, message seed)
CHANNELS = {"PCH01": (50000 + 650, 0.20, 1), "PCH02": (-70000 + 1100, 0.15, 2), "PCH03": (101000 + 650, 0.25, 3)}


def expected_records(msgs):
    """What the reference's ParserISU must hand over for the messages aerol_frames.messages_to_sus sends."""
    want = set()
    for k, (aes, ges, reg, label, text) in enumerate(msgs):
        head = "AES=%06X|GES=%02X|QNO=%02X|REFNO=%02X|MODE=32|TAK=15|BI=31|DL=0|MORE=0|NONACARS=0|LABEL=%s|" % (aes, ges, (k % 15) + 1, k % 16, label)
        full = reg.rjust(7, ".")
        want.add("FRAGMENT|" + head + "REG=%s|TEXT=%s" % (full, text))
        want.add("ACARS|" + head + "REG=%s|TEXT=%s" % (full.lstrip("."), text))     # acarslookupresult strips the dots
    return want


@pytest.mark.parametrize("bitrate", [600, 1200])
def test_frame_generator_is_the_inverse_of_the_reference_frame_decoder(bitrate):
    """Channel bits (hard decisions as soft values 0 / 255) straight into the unmodified AeroL::processDemodulatedSoftBits."""
    msgs = af.example_messages(5, seed=7)
    bits = af.PChannelFramer(bitrate).stream(af.messages_to_sus(msgs))
    assert bits.size % 1200 == 0
    dec = ref_decode.RefDecoder(bitrate)
    soft = bits.astype(np.int16) * 255
    for i in range(0, soft.size, 12):          # the demodulators emit 12 soft bits at a time (mskdemodulator.cpp:423-426)
        dec.feed_softbits(soft[i:i + 12])
    got = dec.records()
    dec.close()
    assert set(got) == expected_records(msgs) and len(got) == 2 * len(msgs)


def test_a_broken_crc_is_rejected_by_the_reference_decoder():
    """Negative control: the checker is live. One signal unit of the second message gets a bad CRC; the reference drops
    that message and still delivers the others."""
    msgs = af.example_messages(3, seed=9)
    sus = af.messages_to_sus(msgs)
    first = af.isu_signal_units(msgs[0][0], msgs[0][1], 1, 0, af.acars_user_data(msgs[0][2], msgs[0][3], msgs[0][4]))
    k = len(first) + 1                          # first SSU of message 2
    sus[k] = sus[k][:11] + bytes([sus[k][11] ^ 0x40])
    bits = af.PChannelFramer(600).stream(sus)
    dec = ref_decode.RefDecoder(600)
    soft = bits.astype(np.int16) * 255
    for i in range(0, soft.size, 12):
        dec.feed_softbits(soft[i:i + 12])
    got = set(dec.records())
    dec.close()
    assert got == expected_records([msgs[0]]) | {r.replace("QNO=01|REFNO=00", "QNO=03|REFNO=02") for r in expected_records([msgs[2]])}


def _make_capture(tmp):
    sent = {}
    carriers = []
    for topic, (offset, amp, seed) in CHANNELS.items():
        msgs = af.example_messages(3, seed=seed)
        sent[topic] = msgs
        bits = af.PChannelFramer(600).stream(af.messages_to_sus(msgs), lead_frames=6, tail_frames=2)
        path = tmp / (topic + ".bits")
        bits.tofile(path)
        carriers.append("--carrier=%d:600:msk:%g:bits=%s" % (offset, amp, path))
    iq = tmp / "cap.cu8"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "synth_iq.py"), str(iq), "--format", "cu8", "--rate", "288000", "--seconds", "28",
                    "--noise", "0.02"] + carriers, check=True, capture_output=True)
    return iq, sent


def _cpu_dump(iq, out):
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "oracle_payloads.py"), INI, str(iq), "cu8", str(out)], check=True, capture_output=True)


def test_cpu_chain_payloads_decode_to_every_sent_message(tmp_path):
    """frames -> MSK -> wideband capture -> reference-exact CPU chain -> reference demodulator and decoder -> the messages.
    Needs aero-publish-b200 only for --plan (settings-file arithmetic, no device)."""
    iq, sent = _make_capture(tmp_path)
    _cpu_dump(iq, tmp_path / "cpu")
    got = ref_decode.decode_dump(str(tmp_path / "cpu"), 600)
    assert sorted(got) == sorted(CHANNELS)
    for topic, msgs in sent.items():
        assert set(got[topic]) == expected_records(msgs), topic


@pytest.mark.gpu
def test_gpu_payloads_decode_to_the_identical_acars_set(tmp_path):
    """The same capture through the product (CUDA bank behind Publisher) and through the CPU chain: the unchanged decoder
    must produce identical record lists per topic - and they must be the messages that were sent."""
    iq, sent = _make_capture(tmp_path)
    gpu, cpu = tmp_path / "gpu", tmp_path / "cpu"
    gpu.mkdir()
    subprocess.run([BIN, "-d", "file=%s,format=cu8" % iq, "--dump", str(gpu), INI], check=True, capture_output=True)
    _cpu_dump(iq, cpu)
    from_gpu = ref_decode.decode_dump(str(gpu), 600)
    from_cpu = ref_decode.decode_dump(str(cpu), 600)
    assert from_gpu == from_cpu
    for topic, msgs in sent.items():
        assert set(from_gpu[topic]) == expected_records(msgs), topic
