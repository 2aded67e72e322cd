#!/usr/bin/env python
"""Replay recorded per-topic payload dumps to ZeroMQ in aero-publish's wire format, so that an
UNCHANGED `aero-decode -p tcp://host:port -t <topic> -b <rate>` can consume them on a machine that has it
(SURVEY.md section 8d item E / 8f-4). Frames: [first 5 bytes of topic][uint32 LE rate][payload]
(/root/reference/publish/zmqpublisher.cpp:61-73; the reference's own test harness tools/audio-publisher:126-128 sends the same three frames).

Dump directory layout (written by `aero-publish-b200 --dump DIR` or tools/oracle_payloads.py):
    DIR/<topic>.i16      concatenated payloads of that topic
    DIR/<topic>.meta     "rate bytes_per_message"

    tools/replay_payloads.py DIR --bind tcp://*:6003 [--realtime]
"""
import argparse
import glob
import os
import struct
import time

import zmq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--bind", default="tcp://*:6003")
    ap.add_argument("--realtime", action="store_true", help="pace messages at the output sample rate")
    ap.add_argument("--settle", type=float, default=1.0, help="seconds to wait for subscribers before sending")
    args = ap.parse_args()
    ctx = zmq.Context.instance()
    pub = ctx.socket(zmq.PUB)
    pub.bind(args.bind)
    time.sleep(args.settle)
    streams = []
    for meta in sorted(glob.glob(os.path.join(args.dir, "*.meta"))):
        topic = os.path.basename(meta)[:-5]
        rate, per_msg = [int(v) for v in open(meta).read().split()]
        data = open(os.path.join(args.dir, topic + ".i16"), "rb").read()
        streams.append((topic, rate, per_msg, data))
    n_msgs = max(len(d) // m for _, _, m, d in streams) if streams else 0
    for k in range(n_msgs):
        t0 = time.time()
        dur = 0.0
        for topic, rate, per_msg, data in streams:
            chunk = data[k * per_msg:(k + 1) * per_msg]
            if not chunk:
                continue
            pub.send_multipart([topic.encode()[:5].ljust(5, b"\0"), struct.pack("<I", rate), chunk])
            dur = max(dur, len(chunk) / 2 / rate)
        if args.realtime:
            time.sleep(max(0.0, dur - (time.time() - t0)))
    print("sent %d messages per topic on %d topics" % (n_msgs, len(streams)))


if __name__ == "__main__":
    main()
