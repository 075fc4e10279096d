"""`python -m starway_b200.bench_cli` — the reference's benchmark CLI, re-implemented for this engine.

Mirrors the reference's harness (reference ``src/starway/bench.py`` + ``src/starway/benchmarks/
scenarios.py``; catalogue in ``benchmark.md:46-102``): the same four scenarios with the same
metric definitions, defaults and tag constants, so results compare scenario for scenario:

  large-array       GB/s of asend + aflush, client -> server   (1 GiB, 1 warm-up + 3 timed)
  small-messages    msg/s over batches of concurrent asend + one aflush (1 KiB x 64, 2 + 10 batches)
  pingpong-flag     RTT of a 1-byte tagged ping/pong (100 warm-up + 1000 timed)
  streaming-duplex  aggregate GB/s, both directions at once (4 MiB, 8 warm-up + 64 timed)

Roles: ``loopback`` (Server + Client in this process, the shape of the reference's tests),
``server`` / ``client`` (two processes; the client needs ``--address HEX`` printed by the server, or
``--host/--port``).  ``--buffers device`` (default when a GPU is visible) uses CUDA tensors,
``--buffers host`` uses NumPy arrays like the reference.  Control frames (READY / DONE) travel over
the tagged channel itself, as in the reference (CONTROL_TAG / READY_TAG / DONE_TAG).
"""
from __future__ import annotations

import argparse
import asyncio
import json
import sys
import time
from typing import Any

import numpy as np

TAG_MASK = (1 << 64) - 1
CONTROL_TAG, READY_TAG, DONE_TAG = 0x1AA0, 0x1AA1, 0x1AA2
LARGE_DATA_TAG, SMALL_DATA_TAG = 0x2B00, 0x2B10
FLAG_PING_TAG, FLAG_PONG_TAG = 0x2B20, 0x2B21
STREAM_UP_TAG, STREAM_DOWN_TAG = 0x2B30, 0x2B31

DEFAULTS = {
    "large-array": {"message_bytes": 1 << 30, "warmup": 1, "iterations": 3},
    "small-messages": {"message_bytes": 1024, "concurrency": 64, "warmup_batches": 2, "iterations": 10},
    "pingpong-flag": {"warmup": 100, "iterations": 1000},
    "streaming-duplex": {"message_bytes": 4 << 20, "warmup": 8, "iterations": 64},
}


def list_scenarios() -> list[str]:
    return list(DEFAULTS)


def parse_size(text: str) -> int:
    t = text.strip().lower().replace("_", "")
    for suf, mul in (("gib", 1 << 30), ("gb", 1 << 30), ("g", 1 << 30), ("mib", 1 << 20), ("mb", 1 << 20), ("m", 1 << 20),
                     ("kib", 1 << 10), ("kb", 1 << 10), ("k", 1 << 10)):
        if t.endswith(suf):
            return int(float(t[: -len(suf)]) * mul)
    return int(float(t))


class Bufs:
    def __init__(self, kind: str):
        self.kind = kind
        if kind == "device":
            import torch

            self.torch = torch

    def filled(self, n: int, value: int):
        if self.kind == "device":
            t = self.torch.full((n,), value, dtype=self.torch.uint8, device="cuda")
            self.torch.cuda.synchronize()
            return t
        return np.full(n, value, dtype=np.uint8)

    def empty(self, n: int):
        return self.filled(n, 0)


# ------------------------------------------------------------------------------- scenario bodies
async def large_array_client(client, cfg, bufs):
    n, warm, iters = cfg["message_bytes"], cfg["warmup"], cfg["iterations"]
    payload = bufs.filled(n, 0x5A)
    durations = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        await client.asend(payload, LARGE_DATA_TAG)
        await client.aflush()
        if i >= warm:
            durations.append(time.perf_counter() - t0)
    total = sum(durations)
    per = [n / d / 1e9 for d in durations]
    # metric keys and samples as the reference's scenario reports them (benchmarks/scenarios.py:109-120)
    return {"total_seconds": total, "avg_seconds_per_iter": total / iters, "avg_gbps": n * iters / total / 1e9,
            "best_gbps": max(per), "worst_gbps": min(per)}, {"duration_seconds": durations, "per_iter_gbps": per}


async def large_array_server(server, ep, cfg, bufs):
    buf = bufs.empty(cfg["message_bytes"])
    for _ in range(cfg["warmup"] + cfg["iterations"]):
        await server.arecv(buf, LARGE_DATA_TAG, TAG_MASK)
    await server.aflush_ep(ep)


async def small_messages_client(client, cfg, bufs):
    n, conc = cfg["message_bytes"], cfg["concurrency"]
    warm, iters = cfg["warmup_batches"], cfg["iterations"]
    payloads = [bufs.filled(n, i % 251) for i in range(conc)]
    durations = []
    for b in range(warm + iters):
        t0 = time.perf_counter()
        await asyncio.gather(*[client.asend(p, SMALL_DATA_TAG) for p in payloads])
        await client.aflush()
        if b >= warm:
            durations.append(time.perf_counter() - t0)
    total = sum(durations)
    per_msg = [d / conc for d in durations]
    lat = np.array(per_msg) * 1e6
    # reference benchmarks/scenarios.py:184-196
    return ({"total_seconds": total, "messages_per_second": iters * conc / total, "bandwidth_gbps": n * iters * conc / total / 1e9,
             "latency_p50_us": float(np.percentile(lat, 50)), "latency_p95_us": float(np.percentile(lat, 95))},
            {"batch_duration_seconds": durations, "avg_latency_seconds": per_msg})


async def small_messages_server(server, ep, cfg, bufs):
    n, conc = cfg["message_bytes"], cfg["concurrency"]
    bufs_ = [bufs.empty(n) for _ in range(conc)]
    for _ in range(cfg["warmup_batches"] + cfg["iterations"]):
        await asyncio.gather(*[server.arecv(b, SMALL_DATA_TAG, TAG_MASK) for b in bufs_])
    await server.aflush_ep(ep)


async def pingpong_client(client, cfg, bufs):
    ping, pong = bufs.filled(1, 1), bufs.empty(1)
    rtts = []
    for i in range(cfg["warmup"] + cfg["iterations"]):
        t0 = time.perf_counter()
        f = client.arecv(pong, FLAG_PONG_TAG, TAG_MASK)
        await client.asend(ping, FLAG_PING_TAG)
        await f
        if i >= cfg["warmup"]:
            rtts.append(time.perf_counter() - t0)
    us = np.array(rtts) * 1e6
    # reference benchmarks/scenarios.py:261-272 (plus the 95th percentile)
    return ({"avg_rtt_us": float(us.mean()), "median_rtt_us": float(np.median(us)), "min_rtt_us": float(us.min()),
             "max_rtt_us": float(us.max()), "avg_one_way_us": float(us.mean()) / 2.0, "p95_rtt_us": float(np.percentile(us, 95))},
            {"rtt_seconds": rtts})


async def pingpong_server(server, ep, cfg, bufs):
    ping, pong = bufs.empty(1), bufs.filled(1, 2)
    for _ in range(cfg["warmup"] + cfg["iterations"]):
        await server.arecv(ping, FLAG_PING_TAG, TAG_MASK)
        await server.asend(ep, pong, FLAG_PONG_TAG)
    await server.aflush_ep(ep)


async def streaming_client(client, cfg, bufs):
    n, warm, iters = cfg["message_bytes"], cfg["warmup"], cfg["iterations"]
    up, down = bufs.filled(n, 0x11), bufs.empty(n)
    durations = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        await asyncio.gather(client.arecv(down, STREAM_DOWN_TAG, TAG_MASK), client.asend(up, STREAM_UP_TAG))
        if i >= warm:
            durations.append(time.perf_counter() - t0)
    total = sum(durations)
    # reference benchmarks/scenarios.py:319-336
    return ({"total_seconds": total, "avg_seconds_per_iter": total / iters, "client_to_server_gbps": n * iters / total / 1e9,
             "server_to_client_gbps": n * iters / total / 1e9, "aggregate_gbps": 2 * n * iters / total / 1e9},
            {"iteration_seconds": durations})


async def streaming_server(server, ep, cfg, bufs):
    n = cfg["message_bytes"]
    up, down = bufs.empty(n), bufs.filled(n, 0x22)
    for _ in range(cfg["warmup"] + cfg["iterations"]):
        await asyncio.gather(server.arecv(up, STREAM_UP_TAG, TAG_MASK), server.asend(ep, down, STREAM_DOWN_TAG))
    await server.aflush_ep(ep)


SCENARIOS = {
    "large-array": (large_array_client, large_array_server),
    "small-messages": (small_messages_client, small_messages_server),
    "pingpong-flag": (pingpong_client, pingpong_server),
    "streaming-duplex": (streaming_client, streaming_server),
}


# ------------------------------------------------------------------------------- control channel
def _frame(obj: Any) -> np.ndarray:
    return np.frombuffer(json.dumps(obj, separators=(",", ":"), sort_keys=True).encode(), dtype=np.uint8).copy()


async def _recv_frame(worker, tag) -> Any:
    buf = np.zeros(4096, dtype=np.uint8)
    _, length = await worker.arecv(buf, tag, TAG_MASK)
    return json.loads(bytes(buf[:length]).decode())


async def run_server_side(server, ep, plan, bufs):
    for name, cfg in plan:
        await server.asend(ep, _frame({"scenario": name, "ready": True}), READY_TAG)
        await SCENARIOS[name][1](server, ep, cfg, bufs)
        await _recv_frame(server, DONE_TAG)


async def run_client_side(client, plan, bufs):
    results = []
    for name, cfg in plan:
        await _recv_frame(client, READY_TAG)
        metrics, samples = await SCENARIOS[name][0](client, cfg, bufs)
        await client.asend(_frame({"scenario": name, "done": True}), DONE_TAG)
        # one entry = the reference's ScenarioResult.to_dict() (benchmarks/scenarios.py:49-57)
        results.append({"name": name, "metrics": metrics, "config": cfg, "samples": samples})
    return results


def build_plan(args) -> list[tuple[str, dict]]:
    names = list_scenarios() if (not args.scenarios or args.scenarios == ["all"]) else args.scenarios
    plan = []
    for name in names:
        if name not in DEFAULTS:
            raise SystemExit(f"Unknown scenario '{name}'. Available: {', '.join(list_scenarios())}")
        cfg = dict(DEFAULTS[name])
        over = {
            "large-array": {"message_bytes": args.large_bytes, "iterations": args.large_iterations, "warmup": args.large_warmup},
            "small-messages": {"message_bytes": args.small_bytes, "iterations": args.small_iterations,
                               "warmup_batches": args.small_warmup, "concurrency": args.small_concurrency},
            "pingpong-flag": {"iterations": args.flag_iterations, "warmup": args.flag_warmup},
            "streaming-duplex": {"message_bytes": args.stream_bytes, "iterations": args.stream_iterations, "warmup": args.stream_warmup},
        }[name]
        cfg.update({k: v for k, v in over.items() if v is not None})
        plan.append((name, cfg))
    return plan


async def amain(args):
    import starway_b200 as sw

    kind = args.buffers
    if kind == "auto":
        kind = "device" if sw.device_count() > 0 else "host"
    bufs = Bufs(kind)
    plan = build_plan(args)
    results = None
    if args.role == "loopback":
        server, client = sw.Server(), sw.Client()
        addr = server.listen_address()
        await client.aconnect_address(addr)
        for _ in range(1000):
            if server.list_clients():
                break
            await asyncio.sleep(0.005)
        ep = next(iter(server.list_clients()))
        srv = asyncio.ensure_future(run_server_side(server, ep, plan, bufs))
        results = await run_client_side(client, plan, bufs)
        await srv
        await client.aclose()
        await server.aclose()
    elif args.role == "server":
        server = sw.Server()
        accepted = asyncio.Event()
        loop = asyncio.get_running_loop()
        server.set_accept_cb(lambda ep: loop.call_soon_threadsafe(accepted.set))
        if args.port:
            server.listen(args.host, args.port)
            print(f"listening on {args.host}:{args.port}", flush=True)
        else:
            print("address " + server.listen_address().hex(), flush=True)
        await accepted.wait()
        ep = next(iter(server.list_clients()))
        await run_server_side(server, ep, plan, bufs)
        await server.aclose()
    else:
        client = sw.Client()
        if args.address:
            await client.aconnect_address(bytes.fromhex(args.address.replace(":", "").strip()))
        else:
            await client.aconnect(args.host, args.port)
        results = await run_client_side(client, plan, bufs)
        await client.aclose()
    if results is not None:
        # the reference's report (src/starway/bench.py:383-405): timestamp, transport, scenarios[]; `transport` is the
        # UCX_TLS variable there -- here the backend and where the buffers live
        report = {"timestamp": time.time(), "transport": f"{sw.backend_name()}:{kind}", "role": args.role,
                  "scenarios": [dict(r) if args.store_trace else {k: v for k, v in r.items() if k != "samples"} for r in results]}
        print("\n=== Benchmark Results ===")
        for r in results:
            print(f"\n[{r['name']}]")
            for k, v in r["metrics"].items():
                print(f"  {k}: {v:.6f}" if isinstance(v, float) else f"  {k}: {v}")
        if args.output:
            import os

            os.makedirs(os.path.dirname(os.path.abspath(args.output)), exist_ok=True)
            with open(args.output, "w") as f:
                json.dump(report, f, indent=2)
            print(f"\nJSON results written to {args.output}")
    sw.shutdown()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m starway_b200.bench_cli", description=__doc__.split("\n\n")[0])
    ap.add_argument("--role", choices=["loopback", "server", "client"], default="loopback")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("--address", default=None, help="hex worker address printed by the server role")
    ap.add_argument("--buffers", choices=["auto", "device", "host"], default="auto")
    ap.add_argument("--scenario", dest="scenarios", action="append", help="repeatable; default: all")
    ap.add_argument("--output", default=None, help="write a JSON report here")
    ap.add_argument("--store-trace", action="store_true", help="keep the per-iteration samples in the JSON report")
    ap.add_argument("--large-bytes", type=parse_size)
    ap.add_argument("--large-iterations", type=int)
    ap.add_argument("--large-warmup", type=int)
    ap.add_argument("--small-bytes", type=parse_size)
    ap.add_argument("--small-iterations", type=int)
    ap.add_argument("--small-warmup", type=int)
    ap.add_argument("--small-concurrency", type=int)
    ap.add_argument("--flag-iterations", type=int)
    ap.add_argument("--flag-warmup", type=int)
    ap.add_argument("--stream-bytes", type=parse_size)
    ap.add_argument("--stream-iterations", type=int)
    ap.add_argument("--stream-warmup", type=int)
    args = ap.parse_args(argv)
    try:
        import uvloop

        uvloop.run(amain(args))
    except ImportError:
        asyncio.run(amain(args))


if __name__ == "__main__":
    main(sys.argv[1:])
