"""First-contact check of the resident progress path on a GPU: a few sizes in both directions of one
in-process connection, bit-exact compare, then the engine counters.  Run under `timeout`."""
import asyncio
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("STARWAY_QUIET", "1")
import starway_b200 as sw  # noqa: E402


async def main():
    for kv in filter(None, os.environ.get("RS_OPTS", "").split(",")):
        k, v = kv.split("=")
        sw.get_context().set_option(k, int(v))
    server, client = sw.Server(), sw.Client()
    await client.aconnect_address(server.listen_address())
    for _ in range(400):
        if server.list_clients():
            break
        await asyncio.sleep(0.005)
    ep = next(iter(server.list_clients()))
    sizes = () if os.environ.get("RS_ONLY_PINGPONG") else (0, 1, 64, 300, 4096, 8128, 8129, 65536, (1 << 20) + 5, 64 << 20)
    for n in sizes:
        for rnd in range(3):
            src = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
            dst = torch.full((n + 32,), 0xEE, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f = server.arecv(dst, 5, 0xFFFF)
            await asyncio.wait_for(client.asend(src, 0x10005), 20)
            got = await asyncio.wait_for(f, 20)
            dt = time.perf_counter() - t0
            torch.cuda.synchronize()
            assert got == (0x10005, n), got
            assert torch.equal(dst[:n], src) and bool((dst[n:] == 0xEE).all()), f"payload mismatch n={n}"
            # and back: server -> client, receive posted AFTER the send (unexpected path)
            s2 = client.arecv(dst, 9, (1 << 64) - 1)
            dst.fill_(0xEE)
            torch.cuda.synchronize()
            await asyncio.sleep(0.002)
            snd = server.asend(ep, src, 9)
            assert await asyncio.wait_for(s2, 20) == (9, n)
            await asyncio.wait_for(snd, 20)
            torch.cuda.synchronize()
            assert torch.equal(dst[:n], src), f"payload mismatch (reverse) n={n}"
        print(f"n={n:>9} ok   last one-way {dt * 1e6:8.1f} us", flush=True)
    # ping-pong latency, 64 B
    a, b = torch.zeros(64, dtype=torch.uint8, device="cuda"), torch.zeros(64, dtype=torch.uint8, device="cuda")
    lat = []
    for i in range(300):
        t0 = time.perf_counter()
        f = server.arecv(a, 1, 0xFFFF)
        await client.asend(b, 1)
        await f
        f = client.arecv(b, 2, 0xFFFF)
        await server.asend(ep, a, 2)
        await f
        lat.append(time.perf_counter() - t0)
    lat = sorted(lat[50:])
    print(f"64 B ping-pong RTT median {lat[len(lat) // 2] * 1e6:.1f} us  min {lat[0] * 1e6:.1f} us", flush=True)
    await asyncio.gather(client.aflush(), server.aflush_ep(ep))
    await client.aclose()
    await server.aclose()


asyncio.run(asyncio.wait_for(main(), 240))
st = sw.get_context().stats()
print({k: v for k, v in st.items() if v})
sw.shutdown()
print("resident smoke ok")
