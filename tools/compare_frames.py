#!/usr/bin/env python
"""Compare two aero-decode text logs (e.g. from the CPU-reference payloads and from the GPU payloads)
as SETS of decoded frames, ignoring the wall-clock timestamps aero-decode prints
(/root/reference/decode/output.cpp:35 uses currentDateTimeUtc). Exit code 0 iff identical.

    aero-decode ... > cpu.log   (fed by tools/replay_payloads.py cpu_dump)
    aero-decode ... > gpu.log   (fed by tools/replay_payloads.py gpu_dump)
    tools/compare_frames.py cpu.log gpu.log
"""
import re
import sys

STAMP = re.compile(r"\b\d{2}:\d{2}:\d{2}(\.\d+)?\b|\b\d{2}-\d{2}-\d{2,4}\b|\b\d{4}-\d{2}-\d{2}[T ]\d{2}:\d{2}:\d{2}(\.\d+)?Z?\b")


def frames(path):
    out = []
    for line in open(path, errors="replace"):
        line = STAMP.sub("<t>", line.rstrip())
        if line:
            out.append(line)
    return out


def main():
    a, b = frames(sys.argv[1]), frames(sys.argv[2])
    sa, sb = set(a), set(b)
    only_a, only_b = sorted(sa - sb), sorted(sb - sa)
    print("%d / %d lines, %d distinct in common" % (len(a), len(b), len(sa & sb)))
    for tag, rows in (("only in " + sys.argv[1], only_a), ("only in " + sys.argv[2], only_b)):
        for r in rows[:20]:
            print("%s: %s" % (tag, r))
    sys.exit(0 if not only_a and not only_b and len(a) == len(b) else 1)


if __name__ == "__main__":
    main()
