"""the end-to-end comparison tooling.
GPU: a synthetic capture with in-band OQPSK/MSK-like carriers -> aero-publish-b200 --dump and the CPU chain
(tests/tools/oracle_payloads.py) produce byte-identical per-topic payload files, i.e. an unchanged aero-decode
fed through tools/replay_payloads.py would see identical input.
CPU: the replay tool emits the reference's 3-frame wire format."""
import filecmp
import os
import socket
import struct
import subprocess
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "aero-cli_b200", "aero-publish-b200")
INI = os.path.join(ROOT, "tests", "data", "two_mains_1920k.ini")


def _synth(path):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "synth_iq.py"), str(path), "--format", "cu8", "--rate", "1920000", "--seconds", "1.5",
                    "--carrier=-379850:10500:oqpsk:0.2", "--carrier=-460850:10500:oqpsk:0.15", "--carrier=507150:1200:msk:0.1", "--noise", "0.03"], check=True)


@pytest.mark.gpu
def test_gpu_and_cpu_payload_dumps_are_identical(tmp_path):
    iq = tmp_path / "cap.cu8"
    _synth(iq)
    gpu, cpu = tmp_path / "gpu", tmp_path / "cpu"
    gpu.mkdir()
    subprocess.run([BIN, "-d", "file=%s,format=cu8" % iq, "--dump", str(gpu), INI], check=True, capture_output=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "oracle_payloads.py"), INI, str(iq), "cu8", str(cpu)], check=True)
    names = sorted(f for f in os.listdir(cpu))
    assert names == sorted(os.listdir(gpu)) and len(names) == 10
    for f in names:
        assert filecmp.cmp(cpu / f, gpu / f, shallow=False), f
    a = np.fromfile(gpu / "AAA01.i16", np.int16).astype(np.float64)
    assert np.sqrt((a * a).mean()) > 3000      # the carrier really is in this VFO's passband


def test_replay_tool_emits_reference_wire_format(tmp_path):
    zmq = pytest.importorskip("zmq")
    d = tmp_path / "dump"
    d.mkdir()
    payload = (np.arange(2400, dtype=np.int16) - 1200)
    (d / "VFO07.i16").write_bytes(payload.tobytes() * 3)
    (d / "VFO07.meta").write_text("12000 4800\n")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = zmq.Context.instance()
    sub = ctx.socket(zmq.SUB)
    sub.setsockopt(zmq.SUBSCRIBE, b"VFO07")
    sub.setsockopt(zmq.RCVTIMEO, 8000)
    proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "replay_payloads.py"), str(d), "--bind", "tcp://127.0.0.1:%d" % port, "--settle", "1.5"])
    try:
        time.sleep(0.3)
        sub.connect("tcp://127.0.0.1:%d" % port)
        got = [sub.recv_multipart() for _ in range(3)]
    finally:
        proc.wait(timeout=20)
        sub.close(0)
    for fr in got:
        assert fr[0] == b"VFO07" and struct.unpack("<I", fr[1])[0] == 12000 and fr[2] == payload.tobytes()
