"""world_size-2 on CPU (torch.distributed / gloo): the multi-rank plumbing bench.py uses —
address-blob all_gather, ring connect (rank r -> r+1), eager + rendezvous traffic in both
directions at once, barrier-ordered close — driven through the host-logic simulator."""
import multiprocessing as mp
import os
import random

import pytest


def _rank_main(rank, world, port, q):
    try:
        os.environ["SW_SIM_DEVICE"] = str(rank)
        os.environ["SW_SIM_DEVICES"] = str(world)
        import asyncio

        import numpy as np
        import torch.distributed as dist

        from tests.hostsim import load

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        api = load()

        async def main():
            server = api.Server()
            addr = server.listen_address()
            addrs = [None] * world
            dist.all_gather_object(addrs, addr)
            client = api.Client()
            await client.aconnect_address(addrs[(rank + 1) % world])
            for _ in range(1000):
                if server.list_clients():
                    break
                await asyncio.sleep(0.005)
            assert len(server.list_clients()) == 1
            sizes = [1, 4096, 8128, 8129, 300000, 1 << 21]
            src = [np.random.default_rng(1000 * rank + i).integers(0, 256, n, dtype=np.uint8) for i, n in enumerate(sizes)]
            dst = [np.zeros(n, dtype=np.uint8) for n in sizes]
            recvs = [server.arecv(d, i, 0xFFFF) for i, d in enumerate(dst)]
            await asyncio.gather(*[client.asend(s, (rank << 32) | i) for i, s in enumerate(src)])
            await client.aflush()
            res = await asyncio.gather(*recvs)
            prev = (rank - 1) % world
            for i, (tag, length) in enumerate(res):
                assert tag == (prev << 32) | i and length == sizes[i]
                want = np.random.default_rng(1000 * prev + i).integers(0, 256, sizes[i], dtype=np.uint8)
                np.testing.assert_array_equal(dst[i], want)
            await asyncio.get_running_loop().run_in_executor(None, dist.barrier)
            await client.aclose()
            await asyncio.get_running_loop().run_in_executor(None, dist.barrier)
            await server.aclose()

        asyncio.run(asyncio.wait_for(main(), 60))
        api.shutdown()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        q.put((rank, f"fail: {e!r}"))


@pytest.mark.parametrize("world", [2, 3])
def test_ring_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = random.randint(20000, 40000)
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(30)
    assert all(v == "ok" for v in results.values()), results
