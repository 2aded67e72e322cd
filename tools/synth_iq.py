#!/usr/bin/env python
"""Synthetic IQ capture generator (replaces SDR hardware for tests and benchmarks).

Writes interleaved I,Q samples (cu8 / cs16 / cf32) containing AWGN plus MSK / OQPSK-like carriers at
given offsets from the centre frequency. By default a carrier carries pseudo-random symbols (in-band
energy of the right bandwidth, so that parity is measured on realistic levels). An MSK carrier can
instead carry a channel bit stream written by tools/aerol_frames.py (one byte per bit): that is a
valid Aero-L P channel which the unchanged aero-decode demodulates and decodes; `aoqpsk` is the 10500 bit/s
root-raised-cosine offset QPSK of the wide P channel.

    tools/synth_iq.py out.cu8 --format cu8 --rate 2400000 --seconds 2 \\
        --carrier 123456:10500:oqpsk:0.1 --carrier -400000:600:msk:0.05 --noise 0.05
    tools/aerol_frames.py p600.bits --bitrate 600 --messages 3
    tools/synth_iq.py out.cu8 --rate 288000 --seconds 20 --carrier 50650:600:msk:0.2:bits=p600.bits
"""
import argparse

import numpy as np


def carrier(n0, n, fs, offset_hz, baud, kind, amp, seed, state, data_bits=None):
    """Samples n0..n0+n-1 of one constant-envelope carrier; `state` carries the MSK phase across calls.
    `data_bits` (0/1 array) replaces the pseudo-random symbols for as long as it lasts."""
    t = np.arange(n0, n0 + n, dtype=np.float64)
    sym = (t * baud / fs).astype(np.int64)
    # symbol k -> +-1 from a counter hash, so any block can be generated independently
    h = (sym.astype(np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    h ^= h >> np.uint64(31)
    bits = ((h >> np.uint64(17)) & np.uint64(1)).astype(np.float64) * 2 - 1
    if data_bits is not None and len(data_bits):
        have = sym < len(data_bits)
        bits[have] = data_bits[sym[have]].astype(np.float64) * 2 - 1
    frac = t * baud / fs - sym
    if kind == "aoqpsk":
        # Aero 10500 bit/s offset QPSK as the reference's OqpskDemodulator expects it: bit n is a root-raised-cosine
        # pulse (roll-off 1, matched to RootRaisedCosine::design(1.0, ..., fb / 2), oqpskdemodulator.cpp:167-173)
        # centred on n / baud, even bits on the I rail and odd bits on the Q rail, i.e. 2 bits per symbol with the Q
        # rail half a symbol late. The receiver resolves rail order and polarity from the unique word.
        def rrc(x):                                   # x in symbols (1 symbol = 2 bit periods), roll-off 1
            d = 1.0 - 16.0 * x * x
            safe = np.where(np.abs(d) < 1e-9, 1.0, d)
            return np.where(np.abs(d) < 1e-9, 1.0, (4.0 / np.pi) * np.cos(2.0 * np.pi * x) / safe)
        u = t * baud / fs                             # time in bit periods
        k0 = np.floor(u).astype(np.int64)
        re = np.zeros(n)
        im = np.zeros(n)
        for dk in range(-9, 11):
            k = k0 + dk
            hk = (k.astype(np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
            hk ^= hk >> np.uint64(31)
            a = ((hk >> np.uint64(17)) & np.uint64(1)).astype(np.float64) * 2 - 1
            a[k < 0] = 0.0
            if data_bits is not None and len(data_bits):
                have = (k >= 0) & (k < len(data_bits))
                a[have] = data_bits[k[have]].astype(np.float64) * 2 - 1
            p = a * rrc((u - k) / 2.0)
            even = (k % 2) == 0
            re += np.where(even, p, 0.0)
            im += np.where(even, 0.0, p)
        w = 2 * np.pi * ((offset_hz / fs * t) % 1.0)
        return amp * (re + 1j * im) * np.exp(1j * w) / 1.8   # 1.8: keeps the peak of the shaped signal near `amp`
    if kind == "msk":
        # MSK: the phase advances by +-pi/2 per symbol; a one raises the frequency, which is the sense the
        # reference's MskDemodulator + differential decoder expect behind the USB demodulator
        acc = state.get("acc", 0.0) + np.cumsum(bits) / max(fs / baud, 1.0)
        state["acc"] = float(acc[-1]) if n else state.get("acc", 0.0)
        ph = 0.5 * np.pi * acc
    else:  # offset QPSK-like: I and Q half-sine pulses, Q delayed by half a symbol
        ph = np.arctan2(bits * np.sin(np.pi * frac), np.roll(bits, int(fs / baud / 2)) * np.cos(np.pi * frac))
    w = 2 * np.pi * ((offset_hz / fs * t) % 1.0)
    return amp * np.exp(1j * (w + ph))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--format", default="cu8", choices=["cu8", "cs16", "cf32"])
    ap.add_argument("--rate", type=int, default=2400000)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--noise", type=float, default=0.05, help="AWGN RMS per rail")
    ap.add_argument("--carrier", action="append", default=[], help="offset_hz:baud:msk|oqpsk|aoqpsk:amplitude[:bits=FILE]")
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    total = int(args.rate * args.seconds)
    blk = 1 << 20
    rng = np.random.default_rng(args.seed)
    states = [dict() for _ in args.carrier]
    payloads = []
    for c in args.carrier:
        extra = c.split(":")[4:]
        payloads.append(np.fromfile(extra[0][5:], np.uint8) if extra and extra[0].startswith("bits=") else None)
    with open(args.out, "wb") as f:
        for n0 in range(0, total, blk):
            n = min(blk, total - n0)
            x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * args.noise
            for i, c in enumerate(args.carrier):
                off, baud, kind, amp = c.split(":")[:4]
                x += carrier(n0, n, args.rate, float(off), float(baud), kind, float(amp), args.seed * 1000 + i, states[i], payloads[i])
            iq = np.empty(2 * n, np.float64)
            iq[0::2], iq[1::2] = x.real, x.imag
            if args.format == "cf32":
                f.write(iq.astype(np.float32).tobytes())
            elif args.format == "cs16":
                f.write(np.clip(np.round(iq * 32767), -32768, 32767).astype(np.int16).tobytes())
            else:
                f.write(np.clip(np.round(iq * 127 + 127.4), 0, 255).astype(np.uint8).tobytes())
    print("wrote %d samples (%s) to %s" % (total, args.format, args.out))


if __name__ == "__main__":
    main()
