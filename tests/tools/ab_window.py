"""A/B harness: one window workload (loopback, device buffers) under engine options from SW_OPTS.
  SW_OPTS=yield_us=0,resident_puts=0 python tests/tools/ab_window.py 65536 [window] [iters]"""
import asyncio
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("STARWAY_QUIET", "1")
import starway_b200 as sw  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
window = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100


async def main():
    for kv in filter(None, os.environ.get("SW_OPTS", "").split(",")):
        k, v = kv.split("=")
        sw.get_context().set_option(k, int(v))
    server, client = sw.Server(), sw.Client()
    await client.aconnect_address(server.listen_address())
    while not server.list_clients():
        await asyncio.sleep(0.005)
    src = torch.randint(0, 256, (window * n,), dtype=torch.uint8, device="cuda")
    dst = torch.zeros(window * n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    srcs = [src[j * n:(j + 1) * n] for j in range(window)]
    dsts = [dst[j * n:(j + 1) * n] for j in range(window)]

    async def one():
        recvs = [server.arecv(d, 1, 0xFFFF) for d in dsts]
        sends = [client.asend(x, 1) for x in srcs]
        for f in sends:
            await f
        await client.aflush()
        for f in recvs:
            await f

    for _ in range(5):
        await one()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(iters):
            await one()
        best = min(best, time.perf_counter() - t0)
    torch.cuda.synchronize()
    assert torch.equal(src, dst)
    print(f"{os.environ.get('SW_OPTS', 'default'):40s} n={n:9d} window={window}: {window * iters / best / 1e6:7.4f} Mmsg/s "
          f"{window * iters * n / best / 1e9:9.2f} GB/s  {best / iters * 1e6:8.1f} us/window", flush=True)
    await client.aclose()
    await server.aclose()


try:
    import uvloop

    uvloop.run(main())
except ImportError:
    asyncio.run(main())
sw.shutdown()
