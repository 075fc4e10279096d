#!/usr/bin/env python
"""tests/tools/latency_trace.py — where one small-message round trip spends its time.

Runs the 64 B device-buffer ping-pong (Server + Client on one GPU) with STARWAY_TRACE enabled and
interleaves the progress thread's pipeline events with time stamps taken in the Python coroutine
(both are CLOCK_MONOTONIC).  Prints the median gap between consecutive events of a round trip.

  STARWAY_TRACE=/tmp/sw_trace python tests/tools/latency_trace.py [--bytes 64] [--iters 300]
"""
from __future__ import annotations

import argparse
import asyncio
import collections
import glob
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("STARWAY_QUIET", "1")
os.environ.setdefault("STARWAY_TRACE", "/tmp/sw_latency_trace")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=64)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--warm", type=int, default=100)
    ap.add_argument("--sim", action="store_true", help="engine on the CPU stand-in backend (tests/hostsim): host-side costs only")
    args = ap.parse_args()
    if args.sim:
        import numpy as np

        from tests import hostsim

        sw = hostsim.load()
        mk = lambda: np.ones(args.bytes, dtype=np.uint8)  # noqa: E731
        sync = lambda: None  # noqa: E731
    else:
        import torch

        import starway_b200 as sw

        dev = torch.device("cuda", 0)
        mk = lambda: torch.ones(args.bytes, dtype=torch.uint8, device=dev)  # noqa: E731
        sync = torch.cuda.synchronize
    marks = []

    async def run():
        for kv in filter(None, os.environ.get("SW_OPTS", "").split(",")):   # e.g. SW_OPTS=resident_puts=0,linger_us=300
            k, v = kv.split("=")
            sw.get_context().set_option(k, int(v))
        server, client = sw.Server(), sw.Client()
        await client.aconnect_address(server.listen_address())
        while not server.list_clients():
            await asyncio.sleep(0.005)
        ep = next(iter(server.list_clients()))
        ping, pong, rping, rpong = mk(), mk(), mk(), mk()
        sync()
        now = time.monotonic
        for i in range(args.warm + args.iters):
            rec = i >= args.warm
            f = server.arecv(rping, 1, 0xFFFF)
            t0 = now()
            s = client.asend(ping, 1)
            t1 = now()
            await s
            t2 = now()
            await f
            t3 = now()
            f = client.arecv(rpong, 2, 0xFFFF)
            await server.asend(ep, pong, 2)
            await f
            t4 = now()
            if rec:
                marks.extend([(t0, "py_asend_call"), (t1, "py_asend_returned"), (t2, "py_send_future_done"),
                              (t3, "py_recv_future_done"), (t4, "py_pong_done")])
        await client.aclose()
        await server.aclose()

    try:
        import uvloop

        uvloop.run(run())
    except ImportError:
        asyncio.run(run())
    sw.shutdown()  # writes the engine trace
    events = list(marks)
    for path in glob.glob(os.environ["STARWAY_TRACE"] + "*"):
        with open(path) as f:
            for line in f:
                p = line.split()
                if len(p) >= 2:
                    try:
                        events.append((float(p[0]), p[1]))
                    except ValueError:
                        pass
    events.sort()
    lo, hi = marks[0][0], marks[-1][0]
    events = [e for e in events if lo <= e[0] <= hi]
    # split into round trips at py_asend_call; report the median offset of the k-th event of each name
    trips, cur = [], None
    for t, name in events:
        if name == "py_asend_call":
            if cur:
                trips.append(cur)
            cur = [(t, name)]
        elif cur is not None:
            cur.append((t, name))
    shape = collections.Counter(tuple(n for _, n in tr) for tr in trips).most_common(1)[0][0]
    sel = [tr for tr in trips if tuple(n for _, n in tr) == shape]
    print(f"{len(sel)}/{len(trips)} round trips share the modal event sequence ({args.bytes} B device buffers)")
    prev = 0.0
    for k, name in enumerate(shape):
        off = statistics.median((tr[k][0] - tr[0][0]) * 1e6 for tr in sel)
        print(f"  +{off:8.2f} us  (d {off - prev:6.2f})  {name}")
        prev = off
    rtt = statistics.median((b[0][0] - a[0][0]) * 1e6 for a, b in zip(trips, trips[1:]))
    print(f"  median round trip {rtt:.2f} us")


if __name__ == "__main__":
    main()
