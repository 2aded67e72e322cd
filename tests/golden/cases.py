"""Golden case list shared by make_golden.py (runs the REFERENCE) and the tests (run oracle / GPU).

Each case is one `vfo` of the reference chain driven block by block:
  name, Fs, B, D, L, mixer Hz, gain, filter_bw, demod_usb, cstyle, scalecomp, blocks, input
`input` is "anchor" (the SURVEY.md section-8c cf32 pattern) or ("raw", fmt, seed, amp): seeded
cu8/cs16/cf32 samples converted to float exactly as the product does (oracle_bind.unpack).
The first four are the SURVEY.md anchors with their FNV-1a-64 hashes.
"""
FMT_CU8, FMT_CS16, FMT_CF32 = 0, 1, 2

CASES = [
    # name            Fs       B       D  L  mixer      gain  bw    usb cs sc blocks input
    ("anchor_1536k_d5", 1536000, 384000, 5, 0, 123456.0, 0.5, 0, 1, 1, 1, 6, "anchor"),
    ("anchor_288k_d1_l6", 288000, 57600, 1, 6, -34567.0, 0.5, 0, 1, 1, 1, 8, "anchor"),
    ("anchor_288k_d0_l6_bw", 288000, 57600, 0, 6, 20000.0, 0.25, 3000, 1, 1, 1, 8, "anchor"),
    ("anchor_1920k_d3_l5", 1920000, 480000, 3, 5, -250000.0, 0.5, 0, 1, 1, 1, 6, "anchor"),
    # 600 / 1200 bps style channels of the ini bank (D = 7 / 6)
    ("bank_1536k_d7", 1536000, 384000, 7, 0, -601234.0, 7.5, 0, 1, 1, 1, 5, ("raw", FMT_CF32, 11, 0.6)),
    ("bank_1536k_d6_bw", 1536000, 384000, 6, 0, 455001.0, 5.0, 6000, 1, 1, 1, 5, ("raw", FMT_CF32, 12, 0.6)),
    # the wideband topology (D=8, late /5) at a tenth of the rate
    ("wide_6144k_d8_l5", 6144000, 1536000, 8, 5, 2345678.0, 0.05, 0, 1, 1, 1, 5, ("raw", FMT_CF32, 13, 0.9)),
    ("wide_6144k_d8_l5_bw", 6144000, 1536000, 8, 5, -1999999.0, 0.05, 1500, 1, 1, 1, 5, ("raw", FMT_CF32, 14, 0.9)),
    # config A: 2.4 MS/s cu8, and a cs16 variant (B = Fs/5)
    ("cfgA_2400k_cu8_d5", 2400000, 480000, 5, 0, 123456.0, 0.05, 0, 1, 1, 1, 6, ("raw", FMT_CU8, 15, 0.8)),
    ("cs16_2400k_d4", 2400000, 480000, 4, 0, -777777.0, 0.05, 0, 1, 1, 1, 6, ("raw", FMT_CS16, 16, 0.8)),
    # compressed-IQ outputs of a main VFO without sub-VFOs (vfo::compress)
    ("iq_nibble_288k_d2", 288000, 57600, 2, 0, 30000.0, 1.0, 0, 0, 1, 2, 4, ("raw", FMT_CF32, 17, 0.9)),
    ("iq_int8_288k_d3", 288000, 57600, 3, 0, -41000.0, 1.0, 0, 0, 0, 1, 4, ("raw", FMT_CF32, 18, 0.9)),
    # zero mixer frequency (rotation exactly (1, 0)) and a negative-frequency D=2 case
    ("zero_mixer_288k_d2", 288000, 57600, 2, 0, 0.0, 0.5, 0, 1, 1, 1, 6, "anchor"),
]


def case_dict(c):
    keys = ["name", "Fs", "B", "D", "L", "mixer", "gain", "filter_bw", "demod_usb", "cstyle", "scalecomp", "blocks", "input"]
    return dict(zip(keys, c))


# Nested main -> sub topologies as Publisher::loadSettings builds them (publisher.cpp:118-219): the main VFO
# (demod_usb = 0, compress style 1) only feeds its sub-VFOs, which mix/decimate its stage-D stream at its
# output rate. (name, Fs, B, main (mixer, D), subs [(mixer, D, L, gain, filter_bw)], blocks, input)
NESTED = [
    # config B shape: main at the centre with out_rate = Fs (D = 0), 600/1200/10500-style subs
    ("nested_1536k_main_d0", 1536000, 384000, (0.0, 0),
     [(-601234.0, 7, 0, 0.075, 0), (455001.0, 6, 0, 0.05, 0), (123456.0, 5, 0, 0.1, 0), (-33000.0, 7, 0, 0.08, 0)], 5, ("raw", FMT_CF32, 31, 0.6)),
    # main decimating by 8 to 240 kHz, subs use the late /5 stage (main_out_rate / 48000 == 5)
    ("nested_1920k_main_d3_late5", 1920000, 480000, (-250000.0, 3),
     [(20000.0, 0, 5, 0.5, 0), (-61000.5, 0, 5, 0.4, 3000), (7000.0, 1, 5, 0.5, 0)], 5, "anchor"),
    # main to 288 kHz (late /6), cu8 input
    ("nested_2304k_main_d3_late6_cu8", 2304000, 576000, (400000.0, 3),
     [(-30000.0, 0, 6, 0.5, 0), (45000.0, 1, 6, 0.5, 0)], 5, ("raw", FMT_CU8, 32, 0.8)),
]


def nested_dict(c):
    keys = ["name", "Fs", "B", "main", "subs", "blocks", "input"]
    return dict(zip(keys, c))
