#!/usr/bin/env python
"""TEST INFRASTRUCTURE: feed a per-topic payload dump (the layout written by `aero-publish-b200 --dump DIR` and by
tests/tools/oracle_payloads.py: TOPIC.i16 + TOPIC.meta "rate bytes_per_block") through the reference's UNMODIFIED
decoder chain (oracle/_ref/libref_decode.so: MskDemodulator / OqpskDemodulator -> SignalHunter -> AeroL, see
oracle/ref_decode_harness.cpp) one ZeroMQ-message-sized block at a time, exactly as aero-decode would receive it,
and print / return the decoded records.

    tests/tools/ref_decode.py dump_dir 600            # every topic at 600 bit/s
    tests/tools/ref_decode.py dump_dir 600 > cpu.log ; tools/compare_frames.py cpu.log gpu.log
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_decode.so")


def load():
    lib = ctypes.CDLL(LIB)
    lib.refdec_create.restype = ctypes.c_void_p
    lib.refdec_create.argtypes = [ctypes.c_int]
    lib.refdec_feed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32]
    lib.refdec_feed_softbits.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    for f in (lib.refdec_output, lib.refdec_log):
        f.restype = ctypes.c_size_t
        f.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.refdec_demod_freq.restype = ctypes.c_double
    lib.refdec_demod_freq.argtypes = [ctypes.c_void_p]
    lib.refdec_destroy.argtypes = [ctypes.c_void_p]
    return lib


class RefDecoder:
    """One aero-decode instance (one topic, one bit rate)."""

    def __init__(self, bitrate, lib=None):
        self.lib = lib or load()
        self.h = self.lib.refdec_create(bitrate)
        if not self.h:
            raise ValueError("bit rate %d is not a continuous channel type of the reference decoder" % bitrate)

    def feed(self, pcm, rate):
        pcm = np.ascontiguousarray(pcm, np.int16)
        self.lib.refdec_feed(self.h, pcm.ctypes.data, pcm.size, rate)

    def feed_softbits(self, bits):
        bits = np.ascontiguousarray(bits, np.int16)
        self.lib.refdec_feed_softbits(self.h, bits.ctypes.data, bits.size)

    def _drain(self, fn):
        n = fn(self.h, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        fn(self.h, buf, n + 1)
        return buf.value.decode("latin-1")

    def records(self):
        return [r for r in self._drain(self.lib.refdec_output).split("\n") if r]

    def log(self):
        return [r for r in self._drain(self.lib.refdec_log).split("\n") if r]

    def close(self):
        if self.h:
            self.lib.refdec_destroy(self.h)
            self.h = None


def decode_dump(dump_dir, bitrate, lib=None):
    """{topic: [records]} for every TOPIC.i16 in the dump."""
    lib = lib or load()
    out = {}
    for name in sorted(os.listdir(dump_dir)):
        if not name.endswith(".i16"):
            continue
        topic = name[:-4]
        rate, nbytes = (int(x) for x in open(os.path.join(dump_dir, topic + ".meta")).read().split())
        pcm = np.fromfile(os.path.join(dump_dir, name), np.int16)
        per = nbytes // 2
        dec = RefDecoder(bitrate, lib)
        for i in range(0, pcm.size - per + 1, per):
            dec.feed(pcm[i:i + per], rate)
        out[topic] = dec.records()
        dec.close()
    return out


def main():
    dump, bitrate = sys.argv[1], int(sys.argv[2])
    for topic, recs in decode_dump(dump, bitrate).items():
        for r in recs:
            print("%s %s" % (topic, r))


if __name__ == "__main__":
    main()
