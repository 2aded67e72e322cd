"""The C++ host side above the C ABI (aero-cli_b200/host: Publisher, vfo, ZmqPublisher, ini reader,
IQ file source), driven through the aero-publish-b200 shell.

CPU part: the settings-file semantics of Publisher::loadSettings (publisher.cpp:55-227) - buffer
split, main-VFO matching, decimation counts, late decimation, mixer arithmetic - against an
independent restatement of those rules in this file.
GPU part (-m gpu): stand-alone `vfo` objects reproduce the SURVEY anchors, and a whole settings file
run from an IQ file produces, per ZeroMQ topic, exactly the bytes of the oracle chain."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

from oracle_bind import FMT_CF32, FMT_CU8, Oracle, fnv1a64, synth_anchor, synth_raw, unpack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "aero-cli_b200", "aero-publish-b200")
DATA = os.path.join(ROOT, "tests", "data")


def _build():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "aero-cli_b200", "csrc")], check=True)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "aero-cli_b200", "host")], check=True)


def read_ini(path):
    top, arrays, group = {}, {}, None
    for line in open(path):
        line = line.strip()
        if not line or line[0] in ";#":
            continue
        if line.startswith("["):
            group = line[1:line.index("]")]
            group = None if group == "General" else group
            continue
        k, v = line.split("=", 1)
        v = v.strip().strip('"')
        if group is None:
            top[k.strip()] = v
        else:
            arrays.setdefault(group, {})[k.strip()] = v
    def arr(name):
        a = arrays.get(name, {})
        n = int(a.get("size", 0))
        return [{k.split("\\", 1)[1]: v for k, v in a.items() if k.startswith("%d\\" % (i + 1))} for i in range(n)]
    return top, arr("main_vfos"), arr("vfos")


def expected_plan(path):
    """publisher.cpp:55-227 restated (integer divisions as written there)."""
    top, mains, vfos = read_ini(path)
    Fs = int(top["sample_rate"])
    center = int(top.get("center_frequency", 0))
    mix_offset = int(top.get("mix_offset", 0))
    bufsplit, buflen = 4, (2 * Fs) // 4
    if ((2 * Fs) // 4) % 512 > 0:
        bufsplit, buflen = 5, (2 * Fs) // 5
    out = []
    M = []
    for m in mains:
        f, r = int(m["frequency"]), int(m["out_rate"])
        D = 0 if Fs // r == 1 else int(math.log2(Fs // r))
        M.append(dict(mixer=center - f, out_rate=int(Fs / 2 ** D)))
        out.append(dict(kind="main", parent=-1, fs=Fs, decim=D, late=0, mixer=float(center - f), block=buflen // 2, usb=0, topic=""))
    subs = {i: [] for i in range(len(M))}
    flat = []
    for v in vfos:
        f = int(v["frequency"]) + mix_offset
        dr, orate = int(v.get("data_rate", 0)), int(v.get("out_rate", 0))
        if orate == 0 and dr > 0:
            orate = {600: 12000, 1200: 24000}.get(dr, 48000)
        mf, mr, mi = 0, Fs, -1
        for a, m in enumerate(M):
            if abs((center - m["mixer"]) - f) < m["out_rate"]:
                mi, mf, mr = a, m["mixer"], m["out_rate"]
                break
        late = 0
        if mr // 48000 == 5:
            D, late = int(math.log2(mr // (5 * orate))), 5
        elif mr // 48000 == 6:
            D, late = int(math.log2(mr // (6 * orate))), 6
        else:
            D = int(math.log2(Fs // orate)) - int(math.log2(Fs // mr))
        rec = dict(kind="sub" if mi >= 0 else "flat", parent=mi, fs=mr, decim=D, late=late, mixer=float((center - mf) - f), block=mr // bufsplit,
                   usb=1, topic=v.get("topic", ""), gain=np.float32(np.float32(float(v.get("gain", 0))) / np.float32(100)),
                   filter_bw=int(v.get("filter_bandwidth", 0)))
        (subs[mi] if mi >= 0 else flat).append(rec)
    ordered = []
    mains_out = [o for o in out if o["kind"] == "main"]
    for i, m in enumerate(mains_out):
        ordered.append(m)
        ordered += subs[i]
    return dict(sample_rate=Fs, block=buflen // 2), ordered + flat


@pytest.mark.parametrize("ini", ["sdr_54W_style_1536k.ini", "two_mains_1920k.ini", "flat_2400k.ini"])
def test_settings_plan_matches_reference_rules(ini):
    _build()
    got = json.loads(subprocess.run([BIN, "--plan", os.path.join(DATA, ini)], check=True, capture_output=True, text=True).stdout)
    head, want = expected_plan(os.path.join(DATA, ini))
    assert got["sample_rate"] == head["sample_rate"] and got["block"] == head["block"]
    assert len(got["vfos"]) == len(want)
    for g, w in zip(got["vfos"], want):
        for k in ("kind", "parent", "fs", "decim", "late", "block", "usb", "topic"):
            assert g[k] == w[k], (k, g, w)
        assert g["mixer"] == w["mixer"]
        if "gain" in w:
            assert np.float32(g["gain"]) == w["gain"] and g["filter_bw"] == w["filter_bw"]


def test_settings_errors(tmp_path):
    _build()
    def plan(text):
        p = tmp_path / "x.ini"
        p.write_text(text)
        r = subprocess.run([BIN, "--plan", str(p)], capture_output=True, text=True)
        return r.returncode, r.stdout
    assert plan("center_frequency=1\n")[0] == 1                         # no sample_rate (publisher.cpp:66-70)
    rc, out = plan("sample_rate=1000000\n")
    assert rc == 1 and "not supported" in out                            # rate whitelist (publisher.cpp:72-75)
    rc, out = plan("sample_rate=288000\ncenter_frequency=100\n[main_vfos]\nsize=1\n1\\frequency=100\n1\\out_rate=288000\n[vfos]\nsize=1\n1\\frequency=900000\n1\\data_rate=600\n1\\topic=FAR01\n")
    assert rc == 1 and "matches no main VFO" in out
    assert subprocess.run([BIN, "--plan", str(tmp_path / "missing.ini")], capture_output=True).returncode == 1


def test_settings_reader_dialect(tmp_path):
    """QSettings IniFormat details the deployed files rely on: an explicit [General] group is the top level, values may be
    quoted, ';' and '#' start comment lines, CRLF line ends, unparsable integers read as 0 (QVariant::toInt)."""
    _build()
    plain = ("sample_rate=288000\ncenter_frequency=10000000\n[main_vfos]\nsize=1\n1\\frequency=10000000\n1\\out_rate=288000\n"
             "[vfos]\nsize=1\n1\\frequency=10050000\n1\\data_rate=600\n1\\gain=150\n1\\filter_bandwidth=0\n1\\topic=AAA01\n")
    fancy = ("; a comment\r\n# another\r\n[General]\r\n sample_rate = \"288000\" \r\ncenter_frequency=10000000\r\n\r\n[main_vfos]\r\nsize=1\r\n"
             "1\\frequency=10000000\r\n1\\out_rate=288000\r\n[vfos]\r\n1\\topic=\"AAA01\"\r\nsize=1\r\n1\\frequency=10050000\r\n"
             "1\\data_rate=600\r\n1\\gain=150\r\n1\\filter_bandwidth=wide\r\nnot a key line\r\n")
    def plan(text):
        p = tmp_path / "d.ini"
        p.write_bytes(text.encode())
        r = subprocess.run([BIN, "--plan", str(p)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        return json.loads(r.stdout)
    assert plan(plain) == plan(fancy)
    got = plan(plain)
    assert [v["kind"] for v in got["vfos"]] == ["main", "sub"] and got["vfos"][1]["topic"] == "AAA01"


def test_settings_main_vfo_limit(tmp_path):
    """The reference holds its main VFOs' children in VFOsub[3] (publisher.h:50): a fourth main VFO is refused up front
    instead of writing past the array."""
    _build()
    text = "sample_rate=1920000\ncenter_frequency=10000000\n[main_vfos]\nsize=4\n" + "".join(
        "%d\\frequency=%d\n%d\\out_rate=240000\n" % (i, 9400000 + 300000 * i, i) for i in range(1, 5))
    p = tmp_path / "m.ini"
    p.write_text(text)
    r = subprocess.run([BIN, "--plan", str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "more than 3 main VFOs" in r.stdout + r.stderr


def test_source_errors_and_no_cpu_fallback(tmp_path):
    """Device-string errors surface before any device work, like a failed SoapySDR::Device::make (publisher.cpp:27-31:
    the constructor logs and leaves isRunning() false, main.cpp:52-55 exits 1). With a valid source and no CUDA device
    the run must fail loudly - there is no CPU path behind Publisher."""
    _build()
    ini = os.path.join(DATA, "flat_2400k.ini")
    def run(dev):
        r = subprocess.run([BIN, "-d", dev, "--hash", ini], capture_output=True, text=True)
        return r.returncode, r.stdout + r.stderr
    rc, out = run("rtlsdr=0")
    assert rc == 1 and "failed to find device" in out
    rc, out = run("file=%s,format=cu8" % (tmp_path / "missing.cu8"))
    assert rc == 1 and "cannot open IQ file" in out
    (tmp_path / "x.cu8").write_bytes(b"\x80" * 1000)
    rc, out = run("file=%s,format=cu4" % (tmp_path / "x.cu8"))
    assert rc == 1 and "unknown IQ format" in out
    rc, out = run("file=%s,format=cu8,throttle=-1" % (tmp_path / "x.cu8"))
    assert rc == 1 and "must not be negative" in out
    import torch
    if not torch.cuda.is_available():
        rc, out = run("synthetic=1,format=cu8,blocks=1")
        assert rc == 1 and "CUDA" in out.upper()


def test_zmq_wire_format_over_a_real_socket():
    """ZmqPublisher through libzmq (dlopen) to a pyzmq SUB socket: three frames - 5 topic bytes, uint32 LE
    rate, payload - as aero-decode's consumer expects (zmqpublisher.cpp:61-73, decode/decode.cpp:283-366)."""
    zmq = pytest.importorskip("zmq")
    _build()
    import glob
    import socket
    import struct

    libs = glob.glob(os.path.join(os.path.dirname(zmq.__file__), "..", "pyzmq.libs", "libzmq*.so*"))
    if not libs:
        pytest.skip("no bundled libzmq to dlopen")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, AERODDC_LIBZMQ=os.path.abspath(libs[0]))
    proc = subprocess.Popen([BIN, "--zmq-selftest", "tcp://127.0.0.1:%d" % port], env=env)
    try:
        ctx = zmq.Context.instance()
        sub = ctx.socket(zmq.SUB)
        sub.setsockopt(zmq.SUBSCRIBE, b"VFO42")
        sub.setsockopt(zmq.RCVTIMEO, 5000)
        sub.connect("tcp://127.0.0.1:%d" % port)
        frames = sub.recv_multipart()
        assert len(frames) == 3
        assert frames[0] == b"VFO42"                       # exactly the first five bytes of the topic
        assert struct.unpack("<I", frames[1])[0] == 48000
        p = np.frombuffer(frames[2], np.int16)
        assert p.size == 600 and np.array_equal(np.diff(p), np.full(599, 7, np.int16))
        sub.close(0)
    finally:
        proc.kill()
        proc.wait()


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_standalone_vfo_objects_reproduce_survey_anchors():
    _build()
    got = json.loads(subprocess.run([BIN, "--anchors"], check=True, capture_output=True, text=True).stdout)
    want = {"ANC00": ("4435d1a5843eed39", 144000, 48000), "ANC01": ("78e7d3dd6abf468a", 76800, 24000),
            "ANC02": ("4d86251b7e306418", 153600, 48000), "ANC03": ("c315edd09d50c16e", 144000, 48000)}
    for t, (fnv, nbytes, rate) in want.items():
        assert got[t]["fnv1a64"] == fnv and got[t]["bytes"] == nbytes and got[t]["rate"] == rate, (t, got[t])


@pytest.mark.gpu
@pytest.mark.parametrize("ini,fmt,blocks", [("two_mains_1920k.ini", FMT_CF32, 5), ("sdr_54W_style_1536k.ini", FMT_CU8, 5), ("flat_2400k.ini", FMT_CU8, 6)])
def test_publisher_from_iq_file_matches_oracle_per_topic(tmp_path, ini, fmt, blocks):
    """Settings file -> Publisher -> IQ file source -> GPU bank -> ZmqPublisher frames, per topic,
    against the oracle driven with the same VFO tree."""
    _build()
    path = os.path.join(DATA, ini)
    head, plan = expected_plan(path)
    Fs, B = head["sample_rate"], head["block"]
    raws = [synth_anchor(b * B, B) if fmt == FMT_CF32 else synth_raw(fmt, b * B, B, seed=91, amp=0.8) for b in range(blocks)]
    f = tmp_path / "iq.bin"
    with open(f, "wb") as fh:
        for r in raws:
            fh.write(r.tobytes())
    dev = "file=%s,format=%s" % (f, {FMT_CF32: "cf32", FMT_CU8: "cu8"}[fmt])
    res = subprocess.run([BIN, "-d", dev, "--hash", path], check=True, capture_output=True, text=True)
    got = json.loads(res.stdout)
    assert "processed %d blocks" % blocks in res.stderr
    # oracle: mains first, then their subs on the main's stage-D stream; flat VFOs on the raw stream
    mains, want = {}, {}
    for i, v in enumerate([p for p in plan if p["kind"] == "main"]):
        mains[i] = Oracle(Fs, B, v["decim"], 0, v["mixer"], 0.01, 0, 0, 1, 1)
    leaves = [(p, Oracle(p["fs"], p["block"], p["decim"], p["late"], p["mixer"], float(p["gain"]), p["filter_bw"])) for p in plan if p["kind"] != "main"]
    acc = {p["topic"]: b"" for p, _ in leaves}
    for r in raws:
        x = r if fmt == FMT_CF32 else unpack(fmt, r)
        mid = {}
        for i, m in mains.items():
            m.process(x)
            mid[i] = m.stage(m.D)
        for p, o in leaves:
            acc[p["topic"]] += o.process(mid[p["parent"]] if p["parent"] >= 0 else x)
    for p, o in leaves:
        t = p["topic"][:5]
        assert got[t]["bytes"] == len(acc[p["topic"]]) and got[t]["rate"] == o.out_rate and got[t]["messages"] == blocks
        assert got[t]["fnv1a64"] == "%016x" % fnv1a64(acc[p["topic"]]), t
    assert len(got) == len(leaves)


@pytest.mark.gpu
def test_publisher_on_two_gpus_matches_one_gpu(tmp_path):
    """`gpus=2` in the device string: VFO subtrees sharded over two GPUs, NCCL broadcast; same bytes per topic."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _build()
    path = os.path.join(DATA, "two_mains_1920k.ini")
    head, _ = expected_plan(path)
    B = head["block"]
    f = tmp_path / "iq.bin"
    with open(f, "wb") as fh:
        for b in range(5):
            fh.write(synth_raw(FMT_CU8, b * B, B, seed=17, amp=0.8).tobytes())
    one = json.loads(subprocess.run([BIN, "-d", "file=%s,format=cu8" % f, "--hash", path], check=True, capture_output=True, text=True).stdout)
    out = subprocess.run([BIN, "-d", "file=%s,format=cu8,gpus=2" % f, "--hash", path], check=True, capture_output=True, text=True).stdout
    two = json.loads(out.strip().splitlines()[-1])      # NCCL may print its version banner on stdout first
    assert one == two and len(one) == 5
