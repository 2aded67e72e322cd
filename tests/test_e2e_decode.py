"""End-to-end check (BASELINE.json configs[4]): the bank's payloads go into
the reference's UNCHANGED decoder and must yield the same decoded ACARS set as the CPU chain's payloads.

  tools/aerol_frames.py        ACARS text -> ISU/SSU signal units -> scrambler, K=7 code, interleaver, UW + header
  tools/synth_iq.py            channel bits -> 600 bit/s MSK or 10500 bit/s offset-QPSK carriers inside a 288 kS/s cu8 capture
  aero-publish-b200 --dump     the product (GPU) : capture -> per-topic int16 payloads            [-m gpu]
  tests/tools/oracle_payloads  the CPU chain     : capture -> per-topic int16 payloads
  tests/tools/ref_decode.py    oracle/_ref/libref_decode.so = reference Msk/OqpskDemodulator + SignalHunter + AeroL,
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
DATA = os.path.join(ROOT, "tests", "data")

pytestmark = pytest.mark.skipif(not os.path.exists(ref_decode.LIB),
                                reason="oracle/_ref/libref_decode.so not built (needs /root/reference/decode: make -C oracle refdecode)")

# topic -> (carrier offset from the capture centre: VFO frequency - centre + audio offset, amplitude, message seed)
# The 1200 bit/s channel type is left out on purpose: Decoder::Decoder leaves the MSK demodulator's fb at its 600 bit/s
# default and only switches Fs to 24000 (decode.cpp:141-145), so the unmodified chain does not lock to such a carrier.
SCENARIOS = {
    "msk600": dict(bitrate=600, kind="msk", ini=os.path.join(DATA, "e2e_288k.ini"), seconds=28, lead=6, messages=3,
                   channels={"PCH01": (50000 + 650, 0.20, 1), "PCH02": (-70000 + 1100, 0.15, 2), "PCH03": (101000 + 650, 0.25, 3)}),
    "oqpsk10500": dict(bitrate=10500, kind="aoqpsk", ini=os.path.join(DATA, "e2e_288k_oqpsk.ini"), seconds=8, lead=6, messages=8,
                       channels={"WCH01": (60000 + 8000, 0.20, 4), "WCH02": (-80000 + 9000, 0.15, 5)}),
}


def expected_records(msgs):
    """What the reference's ParserISU must hand over for the messages aerol_frames.messages_to_sus sends."""
    want = set()
    for k, (aes, ges, reg, label, text) in enumerate(msgs):
        head = "AES=%06X|GES=%02X|QNO=%02X|REFNO=%02X|MODE=32|TAK=15|BI=31|DL=0|MORE=0|NONACARS=0|LABEL=%s|" % (aes, ges, (k % 15) + 1, k % 16, label)
        full = reg.rjust(7, ".")
        want.add("FRAGMENT|" + head + "REG=%s|TEXT=%s" % (full, text))
        want.add("ACARS|" + head + "REG=%s|TEXT=%s" % (full.lstrip("."), text))     # acarslookupresult strips the dots
    return want


@pytest.mark.parametrize("bitrate", [600, 1200, 10500])
def test_frame_generator_is_the_inverse_of_the_reference_frame_decoder(bitrate):
    """Channel bits (hard decisions as soft values 0 / 255) straight into the unmodified AeroL::processDemodulatedSoftBits."""
    msgs = af.example_messages(5, seed=7)
    bits = af.PChannelFramer(bitrate).stream(af.messages_to_sus(msgs))
    assert bits.size % (5250 if bitrate == 10500 else 1200) == 0
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


# The tensor mode applies to raw-fed VFOs with at least six half-band stages on cf32 input: three 600 bit/s channels as flat
# VFOs straight off a 1.536 MS/s cf32 capture (seven stages -> 12 kHz audio).
TENSOR_SCENARIO = dict(bitrate=600, kind="msk", ini=os.path.join(DATA, "e2e_1536k_flat.ini"), seconds=24, lead=6, messages=2, rate=1536000, format="cf32",
                       channels={"TCH01": (250000 + 650, 0.20, 1), "TCH02": (-370000 + 1100, 0.15, 2), "TCH03": (501000 + 650, 0.25, 3)})


def _make_capture(tmp, sc):
    sent = {}
    carriers = []
    for topic, (offset, amp, seed) in sc["channels"].items():
        msgs = af.example_messages(sc["messages"], seed=seed)
        sent[topic] = msgs
        bits = af.PChannelFramer(sc["bitrate"]).stream(af.messages_to_sus(msgs), lead_frames=sc["lead"], tail_frames=3)
        path = tmp / (topic + ".bits")
        bits.tofile(path)
        carriers.append("--carrier=%d:%d:%s:%g:bits=%s" % (offset, sc["bitrate"], sc["kind"], amp, path))
    iq = tmp / ("cap." + sc.get("format", "cu8"))
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "synth_iq.py"), str(iq), "--format", sc.get("format", "cu8"), "--rate", str(sc.get("rate", 288000)), "--seconds",
                    str(sc["seconds"]), "--noise", "0.02"] + carriers, check=True, capture_output=True)
    return iq, sent


def _cpu_dump(ini, iq, out, fmt="cu8"):
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "oracle_payloads.py"), ini, str(iq), fmt, str(out)], check=True, capture_output=True)


def _decode(dump, bitrate):
    """A fresh process per dump: the reference demodulators keep a few function-level statics (oqpskdemodulator.cpp:398-410),
    so two decodes only start from the same state in separate processes - as two aero-decode runs would."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ref_decode.py"), str(dump), str(bitrate)], check=True, capture_output=True,
                       text=True)
    out = {}
    for line in r.stdout.splitlines():
        topic, rec = line.split(" ", 1)
        out.setdefault(topic, []).append(rec)
    return out


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_cpu_chain_payloads_decode_to_every_sent_message(tmp_path, name):
    """frames -> modulator -> wideband capture -> reference-exact CPU chain -> reference demodulator and decoder -> the
    messages. Needs aero-publish-b200 only for --plan (settings-file arithmetic, no device)."""
    sc = SCENARIOS[name]
    iq, sent = _make_capture(tmp_path, sc)
    _cpu_dump(sc["ini"], iq, tmp_path / "cpu")
    got = _decode(tmp_path / "cpu", sc["bitrate"])
    assert sorted(got) == sorted(sc["channels"])
    for topic, msgs in sent.items():
        assert set(got[topic]) == expected_records(msgs), topic


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_gpu_payloads_decode_to_the_identical_acars_set(tmp_path, name):
    """The same capture through the product (CUDA bank behind Publisher) and through the CPU chain: the unchanged decoder
    must produce identical record lists per topic - and they must be the messages that were sent."""
    sc = SCENARIOS[name]
    iq, sent = _make_capture(tmp_path, sc)
    gpu, cpu = tmp_path / "gpu", tmp_path / "cpu"
    gpu.mkdir()
    subprocess.run([BIN, "-d", "file=%s,format=cu8" % iq, "--dump", str(gpu), sc["ini"]], check=True, capture_output=True)
    _cpu_dump(sc["ini"], iq, cpu)
    from_gpu = _decode(gpu, sc["bitrate"])
    from_cpu = _decode(cpu, sc["bitrate"])
    assert from_gpu == from_cpu
    for topic, msgs in sent.items():
        assert set(from_gpu[topic]) == expected_records(msgs), topic


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_tolerance_mode_payloads_decode_to_the_identical_acars_set(tmp_path, name):
    """AERODDC_MODE_FAST (`mode=fast`: fused arithmetic, at most 1 LSB away from the reference's int16) through the
    unchanged decoder: the record list per topic must equal the CPU chain's and the messages that were sent - the
    requirement BASELINE.json puts on any output that is not byte-identical."""
    sc = SCENARIOS[name]
    iq, sent = _make_capture(tmp_path, sc)
    gpu, cpu = tmp_path / "gpu", tmp_path / "cpu"
    gpu.mkdir()
    subprocess.run([BIN, "-d", "file=%s,format=cu8,mode=fast" % iq, "--dump", str(gpu), sc["ini"]], check=True, capture_output=True)
    _cpu_dump(sc["ini"], iq, cpu)
    worst = 0
    for topic in sc["channels"]:
        a = np.fromfile(gpu / (topic + ".i16"), np.int16).astype(np.int32)
        b = np.fromfile(cpu / (topic + ".i16"), np.int16).astype(np.int32)
        assert a.size == b.size and a.size > 0
        worst = max(worst, int(np.abs(a - b).max()))
        err, sig = float(((a - b) ** 2).sum()), float((b ** 2).sum())
        assert err == 0 or 10 * np.log10(sig / err) >= 80.0, topic      # signals of normal level: the stated SNR bound, unconditionally
    assert worst <= 3                                                   # 1e-4 of full scale = 3.27 LSB
    from_gpu = _decode(gpu, sc["bitrate"])
    from_cpu = _decode(cpu, sc["bitrate"])
    assert from_gpu == from_cpu
    for topic, msgs in sent.items():
        assert set(from_gpu[topic]) == expected_records(msgs), topic


@pytest.mark.gpu
def test_tensor_mode_payloads_decode_to_the_identical_acars_set(tmp_path):
    """AERODDC_MODE_TENSOR (`mode=tensor`: mix + five half-band stages as a tcgen05 GEMM, later stages fused in its epilogue)
    on a geometry it applies to, through the unchanged decoder: payloads within the stated tolerance of the CPU chain's
    (max |err| <= 1e-4 FS, SNR >= 80 dB, unconditionally), identical record lists, and the messages that were sent."""
    sc = TENSOR_SCENARIO
    iq, sent = _make_capture(tmp_path, sc)
    gpu, cpu = tmp_path / "gpu", tmp_path / "cpu"
    gpu.mkdir()
    subprocess.run([BIN, "-d", "file=%s,format=cf32,mode=tensor" % iq, "--dump", str(gpu), sc["ini"]], check=True, capture_output=True)
    _cpu_dump(sc["ini"], iq, cpu, "cf32")
    worst = 0
    for topic in sc["channels"]:
        a = np.fromfile(gpu / (topic + ".i16"), np.int16).astype(np.int32)
        b = np.fromfile(cpu / (topic + ".i16"), np.int16).astype(np.int32)
        assert a.size == b.size and a.size > 0
        assert np.any(a != b)                                           # really the tolerance path, not the exact one
        worst = max(worst, int(np.abs(a - b).max()))
        err, sig = float(((a - b) ** 2).sum()), float((b ** 2).sum())
        assert 10 * np.log10(sig / err) >= 80.0, topic
    assert worst <= 3                                                   # 1e-4 of full scale = 3.27 LSB
    from_gpu = _decode(gpu, sc["bitrate"])
    from_cpu = _decode(cpu, sc["bitrate"])
    assert from_gpu == from_cpu
    for topic, msgs in sent.items():
        assert set(from_gpu[topic]) == expected_records(msgs), topic


def test_restated_viterbi_round_trip_and_error_correction():
    """oracle/viterbi_restated.c stands in for libcorrect (parity unpinned). Known-answer properties a K=7 rate-1/2 decoder
    must have: encode -> decode is the identity (hard and soft), and isolated channel errors are corrected."""
    import ctypes
    lib = ctypes.CDLL(ref_decode.LIB)
    lib.correct_convolutional_create.restype = ctypes.c_void_p
    lib.correct_convolutional_create.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    lib.correct_convolutional_encode_len.restype = ctypes.c_size_t
    lib.correct_convolutional_encode_len.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.correct_convolutional_encode.restype = ctypes.c_size_t
    lib.correct_convolutional_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    for f in (lib.correct_convolutional_decode, lib.correct_convolutional_decode_soft):
        f.restype = ctypes.c_ssize_t
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    poly = np.array([109, 79], np.uint16)
    conv = lib.correct_convolutional_create(2, 7, poly.ctypes.data)
    rng = np.random.default_rng(3)
    msg = rng.integers(0, 256, 64, dtype=np.uint8)
    nbits = lib.correct_convolutional_encode_len(conv, msg.size)
    assert nbits == 2 * (8 * msg.size + 8)
    enc = np.zeros((nbits + 7) // 8, np.uint8)
    lib.correct_convolutional_encode(conv, msg.ctypes.data, msg.size, enc.ctypes.data)
    # the same code as the transmit side of tools/aerol_frames.py
    fr = af.PChannelFramer(600)
    mine = fr._encode(np.unpackbits(msg))
    assert np.array_equal(np.unpackbits(enc)[:mine.size], mine)
    out = np.zeros(msg.size + 2, np.uint8)
    assert lib.correct_convolutional_decode(conv, enc.ctypes.data, nbits, out.ctypes.data) > 0
    assert np.array_equal(out[:msg.size], msg)
    bits = np.unpackbits(enc)[:nbits].copy()
    for pos in range(20, nbits - 40, 37):        # one flipped channel bit every 37: far apart relative to the free distance 10
        bits[pos] ^= 1
    hard = np.packbits(bits)
    out[:] = 0
    lib.correct_convolutional_decode(conv, hard.ctypes.data, nbits, out.ctypes.data)
    assert np.array_equal(out[:msg.size], msg)
    soft = np.where(bits > 0, 200, 55).astype(np.uint8)
    soft[::11] = 128                              # erasures
    out[:] = 0
    lib.correct_convolutional_decode_soft(conv, soft.ctypes.data, nbits, out.ctypes.data)
    assert np.array_equal(out[:msg.size], msg)


@pytest.mark.gpu
def test_publisher_over_zeromq_into_the_decoder(tmp_path):
    """The literal drop-in path: aero-publish-b200 (Publisher -> CUDA bank -> vfo::transmitData -> ZmqPublisher) publishes
    on a real ZeroMQ socket; a subscriber collects the three-frame messages exactly as aero-decode's consumer does
    (decode.cpp:318-347) and hands the payloads to the unchanged decoder. Payload bytes must equal the CPU chain's."""
    zmq = pytest.importorskip("zmq")
    import socket
    import struct
    import time
    sc = SCENARIOS["oqpsk10500"]
    iq, sent = _make_capture(tmp_path, sc)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ini = tmp_path / "zmq.ini"
    ini.write_text(open(sc["ini"]).read().replace("tcp://*:6003", "tcp://127.0.0.1:%d" % port))
    ctx = zmq.Context.instance()
    sub = ctx.socket(zmq.SUB)
    for topic in sc["channels"]:
        sub.setsockopt(zmq.SUBSCRIBE, topic.encode())
    sub.setsockopt(zmq.RCVHWM, 100000)
    sub.setsockopt(zmq.RCVTIMEO, 500)
    import glob
    libs = glob.glob(os.path.join(os.path.dirname(zmq.__file__), "..", "pyzmq.libs", "libzmq*.so*"))   # pyzmq's bundled libzmq
    env = dict(os.environ, AERODDC_LIBZMQ=os.path.abspath(libs[0])) if libs else dict(os.environ)
    proc = subprocess.Popen([BIN, "-d", "file=%s,format=cu8,delay=2,throttle=8" % iq, str(ini)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    got = {}
    try:
        sub.connect("tcp://127.0.0.1:%d" % port)
        quiet = 0
        while quiet < 4:                       # until the publisher has exited and the socket stayed silent for 2 s
            try:
                topic, rate, payload = sub.recv_multipart()
                got.setdefault(topic.decode(), []).append((struct.unpack("<I", rate)[0], payload))
                quiet = 0
            except zmq.Again:
                quiet = quiet + 1 if proc.poll() is not None else 0
    finally:
        if proc.poll() is None:
            proc.kill()
        sub.close(0)
    assert proc.returncode == 0, proc.stderr.read().decode()[-400:]
    assert sorted(got) == sorted(sc["channels"])
    cpu, zdump = tmp_path / "cpu", tmp_path / "zmq"
    zdump.mkdir()
    _cpu_dump(sc["ini"], iq, cpu)
    for topic, msgs in got.items():
        assert {r for r, _ in msgs} == {48000} and len({len(p) for _, p in msgs}) == 1
        (zdump / (topic + ".i16")).write_bytes(b"".join(p for _, p in msgs))
        (zdump / (topic + ".meta")).write_text("%d %d\n" % (msgs[0][0], len(msgs[0][1])))
        assert (zdump / (topic + ".i16")).read_bytes() == (cpu / (topic + ".i16")).read_bytes(), topic
    decoded = _decode(zdump, sc["bitrate"])
    for topic, msgs in sent.items():
        assert set(decoded[topic]) == expected_records(msgs), topic


def test_frame_generator_pieces():
    """Internal invariants of tools/aerol_frames.py that do not need the decoder: CRC residue, parity, interleaver bijection,
    scrambler period, frame lengths, ISU segmentation limits."""
    su = af.signal_unit(range(10))
    assert len(su) == 12 and af.crc16(su[:10]) == su[10] | (su[11] << 8)
    assert af.crc16(b"123456789") == 0x906E                       # CRC-16/X-25 check value: the variant AeroLcrc16 computes
    assert all(bin(af.odd_parity(c)).count("1") % 2 == 1 for c in range(128))
    for rate, cols, nbits, nsu in ((600, 6, 1200, 6), (1200, 9, 1200, 6), (10500, 78, 5250, 26)):
        fr = af.PChannelFramer(rate)
        assert sorted(fr.tx_pos) == list(range(64 * cols))        # the interleaver is a permutation of one block
        assert fr.frame([af.fill_in_su()] * nsu).size == nbits == fr.frame_bits
    seq = af.scrambler_sequence(2 * 32767 + 10)
    assert np.array_equal(seq[:32767 + 10], seq[32767:2 * 32767 + 10]) and 0 < seq[:32767].sum() < 32767   # x^15 + x^14 + 1: maximal length
    ud = af.acars_user_data(".N123AB", "H1", "X" * 230)
    sus = af.isu_signal_units(0x123456, 0x90, 3, 4, ud)
    assert len(sus) == 1 + (len(ud) - 2 + 7) // 8 and sus[0][0] == 0x71 and sus[0][6] == len(sus) - 1
    assert [s[0] & 0x3F for s in sus[1:]] == list(range(len(sus) - 2, -1, -1))     # SSU sequence numbers count down to 0
    with pytest.raises(AssertionError):
        af.isu_signal_units(1, 2, 3, 4, bytes(2 + 8 * 64))        # more than 63 SSUs do not fit the 6-bit count


def test_long_and_multi_block_messages_through_the_reference_decoder():
    """A 210-character text (28 SSUs, spanning five 600 bit/s frames) and an ETB-terminated first block of a multi-block
    message: the reference reassembles the first and reports the second as a fragment with 'more to come'."""
    long_text = " ".join("WPT%02d N%04d W%05d" % (i, 4000 + 7 * i, 7000 + 13 * i) for i in range(11))[:210]
    sus = af.isu_signal_units(0xABCDEF, 0x85, 5, 6, af.acars_user_data(".D-AIXY", "H1", long_text))
    sus += af.isu_signal_units(0x400A0B, 0x85, 6, 7, af.acars_user_data(".G-XLEA", "B6", "PART ONE OF TWO", more=True))
    dec = ref_decode.RefDecoder(600)
    soft = af.PChannelFramer(600).stream(sus).astype(np.int16) * 255
    for i in range(0, soft.size, 12):
        dec.feed_softbits(soft[i:i + 12])
    got = dec.records()
    dec.close()
    assert "FRAGMENT|AES=ABCDEF|GES=85|QNO=05|REFNO=06|MODE=32|TAK=15|BI=31|DL=0|MORE=0|NONACARS=0|LABEL=H1|REG=.D-AIXY|TEXT=" + long_text in got
    assert "ACARS|AES=ABCDEF|GES=85|QNO=05|REFNO=06|MODE=32|TAK=15|BI=31|DL=0|MORE=0|NONACARS=0|LABEL=H1|REG=D-AIXY|TEXT=" + long_text in got
    assert "FRAGMENT|AES=400A0B|GES=85|QNO=06|REFNO=07|MODE=32|TAK=15|BI=31|DL=0|MORE=1|NONACARS=0|LABEL=B6|REG=.G-XLEA|TEXT=PART ONE OF TWO" in got
    assert not any(r.startswith("ACARS|AES=400A0B") for r in got)      # the defragmenter waits for the closing block


def test_replayed_cpu_payloads_decode_over_zeromq(tmp_path):
    """The CPU-only twin of test_publisher_over_zeromq_into_the_decoder: tools/replay_payloads.py puts the CPU chain's dump on
    a real ZeroMQ socket in the reference's wire format; what a subscriber receives decodes to the sent messages."""
    zmq = pytest.importorskip("zmq")
    import socket
    import struct
    sc = SCENARIOS["oqpsk10500"]
    iq, sent = _make_capture(tmp_path, sc)
    cpu = tmp_path / "cpu"
    _cpu_dump(sc["ini"], iq, cpu)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = zmq.Context.instance()
    sub = ctx.socket(zmq.SUB)
    sub.setsockopt(zmq.SUBSCRIBE, b"")
    sub.setsockopt(zmq.RCVHWM, 100000)
    sub.setsockopt(zmq.RCVTIMEO, 500)
    proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "replay_payloads.py"), str(cpu), "--bind", "tcp://127.0.0.1:%d" % port, "--settle", "1.5"],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    got = {}
    try:
        sub.connect("tcp://127.0.0.1:%d" % port)
        quiet = 0
        while quiet < 3:
            try:
                topic, rate, payload = sub.recv_multipart()
                got.setdefault(topic.rstrip(b"\0").decode(), []).append((struct.unpack("<I", rate)[0], payload))
                quiet = 0
            except zmq.Again:
                quiet = quiet + 1 if proc.poll() is not None else 0
    finally:
        if proc.poll() is None:
            proc.kill()
        sub.close(0)
    assert proc.returncode == 0 and sorted(got) == sorted(sc["channels"])
    z = tmp_path / "zmq"
    z.mkdir()
    for topic, msgs in got.items():
        (z / (topic + ".i16")).write_bytes(b"".join(p for _, p in msgs))
        (z / (topic + ".meta")).write_text("%d %d\n" % (msgs[0][0], len(msgs[0][1])))
        assert (z / (topic + ".i16")).read_bytes() == (cpu / (topic + ".i16")).read_bytes()
    decoded = _decode(z, sc["bitrate"])
    for topic, msgs in sent.items():
        assert set(decoded[topic]) == expected_records(msgs), topic
