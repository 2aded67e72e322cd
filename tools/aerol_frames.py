#!/usr/bin/env python
"""Aero-L P-channel frame generator: the transmit side that aero-decode's frame decoder undoes.

The reference ships only the receiver (decode/aerol.cpp). For the end-to-end check of the down-converter bank
 the synthetic capture has to carry frames that receiver accepts, so this module is the
inverse of each receive step, in transmit order:

    ACARS text  -> ISU user data with odd parity          (inverse of ParserISU::parse,        aerol.cpp:333-470)
    user data   -> 0x71 ISU + 0xC0.. SSUs, CRC-16 each    (inverse of ISUData::update,         aerol.cpp:158-227;
                                                           CRC = AeroLcrc16::calcusingbytes,   aerol.h:335-368)
    6 SUs/frame -> 576 info bits, LSB first               (inverse of the byte packing,        aerol.cpp:1511-1523)
    scrambler   -> x^15 PRBS restarted every frame        (AeroLScrambler,                     aerol.h:406-440)
    FEC         -> rate 1/2 K=7 convolutional code, polynomials 109 / 79 with the newest bit in the LSB, never
                   flushed between frames                 (AeroL::AeroL SetCode(2, 7, {109, 79}), aerol.cpp:910-914)
    interleaver -> 64 x N block, rows permuted by 27 i mod 64 (inverse of deinterleave_ba,     aerol.cpp:594-611;
                   N = 6 at 600 bit/s, 9 at 1200 bit/s, 78 at 10500 bit/s, aerol.cpp:980-1021)
    framing     -> 600 / 1200 bit/s (MSK): 32-bit unique word 0xE15AE893, 16-bit header (format id 1, super-frame
                   marker, frame counter twice), 1152 channel bits = 1200 bits (aerol.cpp:916-917, 1182-1222);
                   10500 bit/s (OQPSK): the unique word on both rails (64 bits, aerol.cpp:1091-1122,
                   927-931), 16-bit header + 178 dummy bits, 4992 channel bits = 5250 bits (aerol.cpp:1012-1021)

The channel bits then go to the modulators in tools/synth_iq.py (`--carrier ...:msk|aoqpsk:...:bits=FILE`).
"""
import numpy as np

UNIQUE_WORD = 0xE15AE893
ROWS = 64


def crc16(data):
    """CRC-16 as AeroLcrc16::calcusingbytes computes it (reflected 0x1021, preset 0xFFFF, inverted)."""
    crc = 0xFFFF
    for byte in data:
        for k in range(8):
            bit = (byte >> k) & 1
            low = crc & 1
            crc >>= 1
            if low ^ bit:
                crc ^= 0x8408
    return (~crc) & 0xFFFF


def signal_unit(first10):
    """10 octets + CRC (low octet first: the receiver reads crc = su[11] << 8 | su[10])."""
    body = bytes(first10)
    assert len(body) == 10
    c = crc16(body)
    return body + bytes([c & 0xFF, c >> 8])


def fill_in_su():
    return signal_unit([0x01] + [0] * 9)


def odd_parity(ch):
    ch &= 0x7F
    return ch | (0x80 if bin(ch).count("1") % 2 == 0 else 0)


def acars_user_data(reg, label, text, mode="2", tak=0x15, block_id="1", more=False):
    """The user-data octets of an ACARS block as carried in an ISU (layout: aerol.cpp:354-372,383-440)."""
    reg = reg.rjust(7, ".")[:7]
    assert len(label) == 2
    ud = [0xFF, 0xFF, 0x01, odd_parity(ord(mode))]
    ud += [odd_parity(ord(c)) for c in reg]
    ud += [odd_parity(tak)] + [odd_parity(ord(c)) for c in label] + [odd_parity(ord(block_id))]
    if text:
        ud += [0x02] + [odd_parity(ord(c)) for c in text]
    ud += [0x97 if more else 0x83]              # ETB / ETX with parity
    ud += [0x93, 0xAB]                          # block check sequence (not verified by the receiver)
    ud += [0x7F]
    return bytes(ud)


def isu_signal_units(aes, ges, qno, refno, user_data):
    """0x71 initial signal unit + subsequent signal units carrying `user_data`."""
    rest = user_data[2:]
    assert len(user_data) >= 3, "at least one SSU is needed for the receiver to complete the ISU"
    groups = [rest[i:i + 8] for i in range(0, len(rest), 8)]
    n = len(groups)
    assert 1 <= n <= 63
    last = len(groups[-1])
    sus = [signal_unit([0x71, (aes >> 16) & 0xFF, (aes >> 8) & 0xFF, aes & 0xFF, ges, ((qno & 15) << 4) | (refno & 15), n & 0x3F,
                        (last & 15) << 4, user_data[0], user_data[1]])]
    for i, g in enumerate(groups):
        seq = n - 1 - i
        sus.append(signal_unit([0xC0 | seq, ((qno & 15) << 4) | (refno & 15)] + list(g) + [0] * (8 - len(g))))
    return sus


def scrambler_sequence(n):
    state = [1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1]
    out = np.empty(n, np.uint8)
    for a in range(n):
        v = state[0] ^ state[14]
        out[a] = v
        state = [v] + state[:-1]
    return out


class PChannelFramer:
    """Turns a queue of 12-octet signal units into the channel bit stream of a 600, 1200 or 10500 bit/s P channel."""

    def __init__(self, bitrate):
        assert bitrate in (600, 1200, 10500)
        self.cols = {600: 6, 1200: 9, 10500: 78}[bitrate]
        self.wide = bitrate == 10500
        self.sus_per_frame = 26 if self.wide else 6          # 4992 / 2 / 96 or 1152 / 2 / 96
        self.frame_bits = 5250 if self.wide else 1200
        self.reg = 0                  # convolutional encoder register, continuous across frames
        self.frame_no = 0
        self.prbs = scrambler_sequence(96 * self.sus_per_frame)
        perm = (np.arange(ROWS) * 27) % ROWS
        k = np.arange(ROWS * self.cols)
        # receiver: out[j*64 + i] = block[perm[i]*cols + j]  ->  transmitter: block[perm[i]*cols + j] = coded[j*64 + i]
        i, j = k % ROWS, k // ROWS
        self.tx_pos = perm[i] * self.cols + j

    def _encode(self, bits):
        out = np.empty(2 * len(bits), np.uint8)
        reg = self.reg
        for n, b in enumerate(bits):
            reg = ((reg << 1) | int(b)) & 0x7F
            out[2 * n] = bin(reg & 109).count("1") & 1
            out[2 * n + 1] = bin(reg & 79).count("1") & 1
        self.reg = reg
        return out

    def frame(self, sus):
        """One frame from exactly `sus_per_frame` signal units."""
        assert len(sus) == self.sus_per_frame and all(len(s) == 12 for s in sus)
        info = np.unpackbits(np.frombuffer(b"".join(sus), np.uint8), bitorder="little")
        coded = self._encode(info ^ self.prbs)
        data = np.empty_like(coded)
        blk = ROWS * self.cols
        for b in range(0, len(coded), blk):
            seg = np.empty(blk, np.uint8)
            seg[self.tx_pos] = coded[b:b + blk]
            data[b:b + blk] = seg
        fc = self.frame_no & 15
        header = (1 << 12) | ((1 if fc == 0 else 0) << 8) | (fc << 4) | fc
        self.frame_no += 1
        uw = [(UNIQUE_WORD >> (31 - n)) & 1 for n in range(32)]
        hd = [(header >> (15 - n)) & 1 for n in range(16)]
        if self.wide:
            uw = [b for b in uw for _ in (0, 1)]             # the same word on the I and on the Q rail
            hd += [(n * 7 + n // 3) & 1 for n in range(178)]  # dummy bits (the receiver drops them)
        return np.concatenate([np.array(uw + hd, np.uint8), data])

    def stream(self, sus, lead_frames=2, tail_frames=2):
        """Channel bits for all `sus` (padded with fill-in units), with fill-in frames before and after: the receiver
        hands a frame over only while it receives the next one (Viterbi + delay line = one frame, aerol.cpp:1497-1509)."""
        n = self.sus_per_frame
        q = [fill_in_su()] * (n * lead_frames) + list(sus)
        while len(q) % n:
            q.append(fill_in_su())
        q += [fill_in_su()] * (n * tail_frames)
        return np.concatenate([self.frame(q[i:i + n]) for i in range(0, len(q), n)])


def example_messages(n, seed=1):
    """Deterministic ACARS uplinks used by the tests and the README walk-through: (aes, ges, reg, label, text)."""
    rng = np.random.default_rng(seed)
    words = ["WX", "ETA", "FUEL", "GATE", "RWY", "POS", "REQ", "ATIS", "OPS", "MSG", "FL350", "N4512", "W07355", "CLR", "ROGER"]
    out = []
    for k in range(n):
        aes = int(rng.integers(0x400000, 0xAFFFFF))
        reg = "".join(rng.choice(list("ABCDEFGHJKLMNPRSTUVWXYZ0123456789"), 5))
        label = str(rng.choice(["H1", "Q0", "10", "5Z", "B6", "SA"]))
        text = " ".join(str(w) for w in rng.choice(words, int(rng.integers(2, 12)))) + " #%d" % k
        out.append((aes, 0x90 + (k % 4), "." + reg, label, text))
    return out


def messages_to_sus(messages):
    sus = []
    for k, (aes, ges, reg, label, text) in enumerate(messages):
        sus += isu_signal_units(aes, ges, qno=(k % 15) + 1, refno=k % 16, user_data=acars_user_data(reg, label, text))
    return sus


def main():
    import argparse
    ap = argparse.ArgumentParser(description="write the channel bits (one byte per bit) of a synthetic P channel")
    ap.add_argument("out")
    ap.add_argument("--bitrate", type=int, default=600, choices=[600, 1200, 10500])
    ap.add_argument("--messages", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--lead", type=int, default=6, help="fill-in frames before the first message: the receiver needs "
                    "up to ~8 s to hunt to the carrier and lock (SignalHunter steps 450 Hz every 15 spectra)")
    ap.add_argument("--tail", type=int, default=2)
    args = ap.parse_args()
    msgs = example_messages(args.messages, args.seed)
    bits = PChannelFramer(args.bitrate).stream(messages_to_sus(msgs), args.lead, args.tail)
    bits.tofile(args.out)
    for m in msgs:
        print("AES=%06X GES=%02X REG=%s LABEL=%s TEXT=%s" % m)
    print("%d bits = %.1f s at %d bit/s" % (len(bits), len(bits) / args.bitrate, args.bitrate))


if __name__ == "__main__":
    main()
