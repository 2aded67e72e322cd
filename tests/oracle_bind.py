"""ctypes bindings for the two CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  - oracle/libddc_oracle.so, the C restatement (always available once built)
* ``RefVfo``  - oracle/_ref/libref_vfo.so, the unmodified reference ``vfo`` class compiled from
  /root/reference/publish (present only where that tree was available at build time)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libddc_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_vfo.so")

FMT_CU8, FMT_CS16, FMT_CF32 = 0, 1, 2


def build_oracle():
    """Build the checkers (does nothing for _ref when /root/reference is absent)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = ctypes.CDLL(ORACLE_SO)
        L.ddc_oracle_create.restype = ctypes.c_void_p
        L.ddc_oracle_create.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_float] + [ctypes.c_int] * 4
        L.ddc_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.ddc_oracle_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ddc_oracle_process_repeat.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ddc_oracle_out_bytes.argtypes = [ctypes.c_void_p]
        L.ddc_oracle_out_rate.argtypes = [ctypes.c_void_p]
        L.ddc_oracle_stage.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ddc_lowpass_taps.argtypes = [ctypes.c_double] * 4 + [ctypes.c_void_p, ctypes.c_int]
        L.ddc_hilbert_taps.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ddc_nco_table.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        L.ddc_unpack.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
        L.ddc_dc_correct.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
        _oracle = L
    return _oracle


def ref_lib():
    """The compiled unmodified reference, or None when it was never built here."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        L = ctypes.CDLL(REF_SO)
        L.refvfo_create.restype = ctypes.c_void_p
        L.refvfo_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_float, ctypes.c_double,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.refvfo_add_sub.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.refvfo_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.refvfo_process_repeat.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.refvfo_pop.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p, ctypes.c_int]
        L.refvfo_stage.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.refvfo_destroy.argtypes = [ctypes.c_void_p]
        L.ref_lowpass.argtypes = [ctypes.c_double] * 4 + [ctypes.c_void_p, ctypes.c_int]
        L.ref_hilbert.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ref_nco.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
        _ref = L
    return _ref


class Oracle:
    """One VFO of the restated chain. process(block cf32 interleaved) -> payload bytes."""

    def __init__(self, Fs, B, D, L, mixer, gain, filter_bw=0, demod_usb=1, cstyle=1, scalecomp=1):
        self.lib = oracle_lib()
        self.h = self.lib.ddc_oracle_create(Fs, B, D, L, float(mixer), float(gain), int(filter_bw), demod_usb, cstyle, scalecomp)
        if not self.h:
            raise ValueError("oracle rejected the configuration")
        self.B, self.D = B, D
        self.out_bytes = self.lib.ddc_oracle_out_bytes(self.h)
        self.out_rate = self.lib.ddc_oracle_out_rate(self.h)
        self._out = np.empty(self.out_bytes, np.uint8)

    def process(self, iq_f32):
        iq_f32 = np.ascontiguousarray(iq_f32, dtype=np.float32)
        assert iq_f32.size == 2 * self.B
        n = self.lib.ddc_oracle_process(self.h, iq_f32.ctypes.data, self.B, self._out.ctypes.data)
        assert n == self.out_bytes, n
        return self._out.tobytes()

    def process_repeat(self, iq_f32, n_blocks):
        iq_f32 = np.ascontiguousarray(iq_f32, dtype=np.float32)
        self.lib.ddc_oracle_process_repeat(self.h, iq_f32.ctypes.data, self.B, n_blocks, self._out.ctypes.data)

    def stage(self, s):
        n = self.B >> s
        out = np.empty(2 * n, np.float32)
        self.lib.ddc_oracle_stage(self.h, s, out.ctypes.data, n)
        return out

    def close(self):
        if self.h:
            self.lib.ddc_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefVfo:
    """One reference `vfo` object (optionally with sub-VFOs attached by add_sub)."""

    _count = 0

    def __init__(self, Fs, B, D, L, mixer, gain, filter_bw=0, demod_usb=1, cstyle=1, scalecomp=1, topic=None):
        self.lib = ref_lib()
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libref_vfo.so not built")
        if topic is None:
            RefVfo._count += 1
            topic = "V%04d" % (RefVfo._count % 10000)
        self.topic = topic
        self.h = self.lib.refvfo_create(Fs, D, float(mixer), float(gain), float(filter_bw), demod_usb, cstyle, scalecomp,
                                        topic.encode(), B, L)
        self.B, self.D = B, D
        self.parented = False
        self._buf = (ctypes.c_ubyte * 400000)()

    def add_sub(self, sub):
        self.lib.refvfo_add_sub(self.h, sub.h)
        sub.parented = True

    def process(self, iq_f32):
        """Returns {topic5: (rate, payload bytes)} for every message emitted by this call."""
        iq_f32 = np.ascontiguousarray(iq_f32, dtype=np.float32)
        self.lib.refvfo_process(self.h, iq_f32.ctypes.data, iq_f32.size // 2)
        out = {}
        topic = ctypes.create_string_buffer(8)
        rate = ctypes.c_uint32()
        while True:
            n = self.lib.refvfo_pop(topic, ctypes.byref(rate), self._buf, len(self._buf))
            if n < 0:
                break
            out[topic.value.decode()] = (rate.value, bytes(self._buf[:n]))
        return out

    def process_repeat(self, iq_f32, n_blocks):
        iq_f32 = np.ascontiguousarray(iq_f32, dtype=np.float32)
        self.lib.refvfo_process_repeat(self.h, iq_f32.ctypes.data, iq_f32.size // 2, n_blocks)

    def stage(self, s):
        n = self.B >> s
        out = np.empty(2 * n, np.float32)
        self.lib.refvfo_stage(self.h, s, out.ctypes.data, n)
        return out

    def close(self):
        if self.h and not self.parented:
            self.lib.refvfo_destroy(self.h)
        self.h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------------------------------------
# deterministic inputs
# ---------------------------------------------------------------------------------------------
def synth_anchor(n0, n):
    """The SURVEY.md section-8c anchor input: interleaved cf32 for samples n0 .. n0+n-1."""
    k = np.arange(n0, n0 + n, dtype=np.uint64)
    re = (((k * np.uint64(7919) + np.uint64(13)) % np.uint64(2001)).astype(np.int64) - 1000).astype(np.float32) / np.float32(1000.0)
    im = (((k * np.uint64(104729) + np.uint64(7)) % np.uint64(2001)).astype(np.int64) - 1000).astype(np.float32) / np.float32(1000.0)
    out = np.empty(2 * n, np.float32)
    out[0::2] = re
    out[1::2] = im
    return out


def synth_raw(fmt, n0, n, seed=1234, amp=0.6):
    """Seeded noise + a few tones, as raw bytes of format fmt, for samples n0..n0+n-1.

    Counter-based (each sample depends only on its absolute index) so any block can be
    regenerated independently on the CPU and GPU sides."""
    k = np.arange(n0, n0 + n, dtype=np.uint64)
    def h(x, salt):
        x = (x + np.uint64(salt)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
        return x
    with np.errstate(over="ignore"):
        u1 = (h(k, seed) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
        u2 = (h(k, seed + 7777) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    ph = 2 * np.pi * ((k.astype(np.float64) * 0.0173) % 1.0)
    re = amp * (0.5 * (u1 - 0.5) * 2 + 0.4 * np.cos(ph))
    im = amp * (0.5 * (u2 - 0.5) * 2 + 0.4 * np.sin(ph))
    if fmt == FMT_CF32:
        out = np.empty(2 * n, np.float32)
        out[0::2] = re
        out[1::2] = im
        return out
    if fmt == FMT_CS16:
        out = np.empty(2 * n, np.int16)
        out[0::2] = np.clip(np.round(re * 32767), -32768, 32767)
        out[1::2] = np.clip(np.round(im * 32767), -32768, 32767)
        return out
    out = np.empty(2 * n, np.uint8)
    out[0::2] = np.clip(np.round(re * 127 + 127.4), 0, 255)
    out[1::2] = np.clip(np.round(im * 127 + 127.4), 0, 255)
    return out


def unpack(fmt, raw):
    """raw numpy array of format fmt -> interleaved float32, via the oracle's conversion."""
    raw = np.ascontiguousarray(raw)
    n = raw.size // 2
    out = np.empty(2 * n, np.float32)
    oracle_lib().ddc_unpack(fmt, raw.ctypes.data, n, out.ctypes.data)
    return out


def dc_correct(iq_f32, state):
    """In-place DC removal (publisher.cpp:292-296) of an interleaved float32 block; state = float32[2] running average."""
    oracle_lib().ddc_dc_correct(iq_f32.ctypes.data, iq_f32.size // 2, state.ctypes.data)
    return iq_f32


def fnv1a64(data, h=1469598103934665603):
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h
