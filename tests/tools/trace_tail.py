#!/usr/bin/env python
"""tests/tools/trace_tail.py <trace files...> [-n N] — the last N events of a STARWAY_TRACE file, microseconds
relative to the first one printed (progress-thread pipeline events; see trace() in starway_b200/csrc/engine.cpp)."""
import sys

n = 120
args = sys.argv[1:]
if "-n" in args:
    i = args.index("-n")
    n = int(args[i + 1])
    del args[i:i + 2]
ev = []
for path in args:
    with open(path) as f:
        for line in f:
            p = line.split()
            try:
                ev.append((float(p[0]), p[1], " ".join(p[2:])))
            except (ValueError, IndexError):
                pass
ev.sort()
ev = ev[-n:]
t0, prev = ev[0][0], ev[0][0]
for t, name, rest in ev:
    print(f"+{(t - t0) * 1e6:9.2f} (d {(t - prev) * 1e6:7.2f})  {name:<14} {rest}")
    prev = t
