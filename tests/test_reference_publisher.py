"""The reference's own Publisher, compiled unmodified (oracle/_ref/ref_publish = publish/publisher.cpp + the vfo chain
behind Qt / SoapySDR / libzmq stand-ins, see oracle/ref_publisher_harness.cpp), pins the host-side rows of SURVEY.md
section 8: a1 (demodData + DC correction), the settings-file semantics of Publisher::loadSettings (8b "config arithmetic
to mirror") and the main -> sub VFO tree (8f items 1 and 3).

CPU: the product's settings reader (`aero-publish-b200 --plan`) builds the same tree as the reference, and the oracle
chain that the GPU tests compare against (tests/tools/oracle_payloads.py) is byte-identical to the reference Publisher's
ZeroMQ payloads, with and without DC correction.
GPU: the product's payloads equal the reference Publisher's directly.
"""
import filecmp
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "aero-cli_b200", "aero-publish-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_publish")
DATA = os.path.join(ROOT, "tests", "data")
ORACLE_DUMP = os.path.join(ROOT, "tests", "tools", "oracle_payloads.py")

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_publish not built (needs /root/reference/publish: make -C oracle refpublish)")

INIS = ["sdr_54W_style_1536k.ini", "two_mains_1920k.ini", "e2e_288k.ini", "e2e_288k_oqpsk.ini"]


def _capture(path, fs, fmt, blocks, seed=5, dc=0.0):
    """Noise + a few tones, `blocks` settings-file blocks long (+ a partial block that both sides must drop)."""
    n = blocks * ((2 * fs) // 4 // 2) + 1000
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = 0.25 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for k, f in enumerate((0.031, -0.117, 0.204, -0.333)):
        x += 0.15 * np.exp(2j * np.pi * (f * t + 0.1 * k))
    x += dc
    iq = np.empty(2 * n)
    iq[0::2], iq[1::2] = x.real, x.imag
    if fmt == "cf32":
        iq.astype(np.float32).tofile(path)
    elif fmt == "cs16":
        np.clip(np.round(iq * 20000), -32768, 32767).astype(np.int16).tofile(path)
    else:
        np.clip(np.round(iq * 100 + 127.4), 0, 255).astype(np.uint8).tofile(path)


def _ref_run(ini, iq, fmt, dcc, out):
    out.mkdir()
    r = subprocess.run([REF, ini, str(iq), fmt, "1" if dcc else "0", str(out)], capture_output=True, text=True)
    return r.returncode, json.loads(r.stdout.strip().splitlines()[-1])


def _fs(ini):
    for line in open(ini):
        if line.startswith("sample_rate="):
            return int(line.split("=")[1])


@pytest.mark.parametrize("name", INIS)
def test_settings_tree_equals_the_reference_publishers(tmp_path, name):
    """Publisher::loadSettings on the real code: same main / sub VFOs, parents, mixer frequencies, output rates and
    block length as the product's reader."""
    ini = os.path.join(DATA, name)
    iq = tmp_path / "c.cu8"
    _capture(iq, _fs(ini), "cu8", 1)
    rc, ref = _ref_run(ini, iq, "cu8", False, tmp_path / "ref")
    assert rc == 0
    got = json.loads(subprocess.run([BIN, "--plan", ini], check=True, capture_output=True, text=True).stdout)
    assert got["sample_rate"] == ref["sample_rate"] and 2 * got["block"] == ref["buflen"]
    assert len(got["vfos"]) == len(ref["vfos"])
    for g, w in zip(got["vfos"], ref["vfos"]):
        assert (g["kind"], g["parent"], g["mixer"], g["usb"]) == (w["kind"], w["parent"], w["mixer"], w["usb"]), (g, w)
        assert g["fs"] >> g["decim"] == w["out_rate"], (g, w)        # vfo::getOutRate() = Fs / 2^D
    # every sub-VFO topic published one message per block at the rate the product's plan implies
    leaves = [v for v in got["vfos"] if v["kind"] != "main"]
    assert sorted(ref["topics"]) == sorted(v["topic"][:5] for v in leaves)
    for v in leaves:
        assert ref["topics"][v["topic"][:5]]["rate"] == (v["fs"] >> v["decim"]) // max(v["late"], 1)


def test_reference_rate_whitelist_and_errors(tmp_path):
    """The reference accepts 288000 / 1536000 / 1920000 only (publisher.h:32); the product adds 2400000 and 61440000
    for BASELINE's configurations and says so. Anything else is rejected by both."""
    iq = tmp_path / "c.cu8"
    _capture(iq, 288000, "cu8", 1)
    rc, ref = _ref_run(os.path.join(DATA, "flat_2400k.ini"), iq, "cu8", False, tmp_path / "a")
    assert rc == 1 and "error" in ref
    bad = tmp_path / "bad.ini"
    bad.write_text("sample_rate=1000000\n")
    rc, ref = _ref_run(str(bad), iq, "cu8", False, tmp_path / "b")
    assert rc == 1
    assert subprocess.run([BIN, "--plan", str(bad)], capture_output=True).returncode == 1
    rc, ref = _ref_run(os.path.join(DATA, "e2e_288k.ini"), tmp_path / "missing.cu8", "cu8", False, tmp_path / "c")
    assert rc == 1


@pytest.mark.parametrize("name,fmt,dcc", [("e2e_288k.ini", "cu8", False), ("e2e_288k.ini", "cf32", True), ("e2e_288k_oqpsk.ini", "cs16", True),
                                          ("two_mains_1920k.ini", "cu8", False), ("two_mains_1920k.ini", "cf32", True),
                                          ("sdr_54W_style_1536k.ini", "cu8", False)])
def test_oracle_chain_equals_the_reference_publisher(tmp_path, name, fmt, dcc):
    """The CPU chain the GPU tests are checked against, pinned on the whole reference Publisher: byte-identical payloads
    per topic, including the DC-correction recurrence (publisher.cpp:292-296) whose state runs across blocks."""
    ini = os.path.join(DATA, name)
    iq = tmp_path / ("c." + fmt)
    _capture(iq, _fs(ini), fmt, 3, dc=0.08 if dcc else 0.0)
    rc, ref = _ref_run(ini, iq, fmt, dcc, tmp_path / "ref")
    assert rc == 0 and ref["dcc"] == int(dcc)
    subprocess.run([sys.executable, ORACLE_DUMP, ini, str(iq), fmt, str(tmp_path / "cpu")] + (["dcc"] if dcc else []), check=True, capture_output=True)
    names = sorted(os.listdir(tmp_path / "ref"))
    assert names == sorted(os.listdir(tmp_path / "cpu")) and names
    for f in names:
        assert filecmp.cmp(tmp_path / "ref" / f, tmp_path / "cpu" / f, shallow=False), f
    assert all(t["messages"] == 3 for t in ref["topics"].values())      # the partial fourth block was dropped


@pytest.mark.gpu
@pytest.mark.parametrize("name,fmt,dcc", [("e2e_288k.ini", "cu8", True), ("two_mains_1920k.ini", "cf32", False), ("sdr_54W_style_1536k.ini", "cs16", True)])
def test_product_equals_the_reference_publisher(tmp_path, name, fmt, dcc):
    """aero-publish-b200 (CUDA bank behind the Publisher mirror) against the unmodified reference Publisher, directly."""
    ini = os.path.join(DATA, name)
    iq = tmp_path / ("c." + fmt)
    _capture(iq, _fs(ini), fmt, 4, dc=0.08 if dcc else 0.0)
    rc, ref = _ref_run(ini, iq, fmt, dcc, tmp_path / "ref")
    assert rc == 0
    gpu = tmp_path / "gpu"
    gpu.mkdir()
    subprocess.run([BIN, "-d", "file=%s,format=%s" % (iq, fmt)] + (["--enable-dcc"] if dcc else []) + ["--dump", str(gpu), ini], check=True, capture_output=True)
    names = sorted(os.listdir(tmp_path / "ref"))
    assert names == sorted(os.listdir(gpu)) and names
    for f in names:
        assert filecmp.cmp(tmp_path / "ref" / f, gpu / f, shallow=False), f


GPUVFO = os.path.join(ROOT, "oracle", "_ref", "ref_publish_gpuvfo")


@pytest.mark.skipif(not os.path.exists(GPUVFO), reason="oracle/_ref/ref_publish_gpuvfo not built (make -C oracle refpublish_gpuvfo)")
def test_reference_publisher_on_the_product_vfo_class_has_no_cpu_path(tmp_path):
    """oracle/_ref/ref_publish_gpuvfo = the reference's unmodified publisher.cpp linked against the PRODUCT's vfo class.
    Without a CUDA device the first vfo::process must fail loudly (no CPU fallback behind the class surface)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the GPU variant of this test covers the path")
    ini = os.path.join(DATA, "e2e_288k.ini")
    iq = tmp_path / "c.cu8"
    _capture(iq, 288000, "cu8", 1)
    (tmp_path / "out").mkdir()
    r = subprocess.run([GPUVFO, ini, str(iq), "cu8", "0", str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not os.listdir(tmp_path / "out")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GPUVFO), reason="oracle/_ref/ref_publish_gpuvfo not built (make -C oracle refpublish_gpuvfo)")
@pytest.mark.parametrize("name,fmt,dcc", [("e2e_288k.ini", "cu8", True), ("two_mains_1920k.ini", "cf32", False), ("two_mains_1920k.ini", "cu8", True),
                                          ("sdr_54W_style_1536k.ini", "cs16", False)])
def test_reference_publisher_source_drives_the_product_vfo_class(tmp_path, name, fmt, dcc):
    """Drop-in at link level: the reference's own Publisher (settings, reader loop, DC correction - all its code) calls
    setFs / setDecimationCount / ... / init / process on the product's vfo objects, one private GPU bank per main VFO.
    The payloads must equal those of the all-reference build."""
    ini = os.path.join(DATA, name)
    iq = tmp_path / ("c." + fmt)
    _capture(iq, _fs(ini), fmt, 4, dc=0.08 if dcc else 0.0)
    rc, ref = _ref_run(ini, iq, fmt, dcc, tmp_path / "ref")
    assert rc == 0
    gpu = tmp_path / "gpu"
    gpu.mkdir()
    r = subprocess.run([GPUVFO, ini, str(iq), fmt, "1" if dcc else "0", str(gpu)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr[-400:]
    names = sorted(os.listdir(tmp_path / "ref"))
    assert names == sorted(os.listdir(gpu)) and names
    for f in names:
        assert filecmp.cmp(tmp_path / "ref" / f, gpu / f, shallow=False), f
