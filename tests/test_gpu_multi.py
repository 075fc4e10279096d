"""Two GPUs, two processes: bit-exact eager and rendezvous transfers over NVLink peer access
(the single-GPU suite covers the same protocol through CUDA-IPC on one device).

Opt-in (`STARWAY_TEST_MULTI_GPU=1`) and skipped with fewer than two devices: the round-end GPU run
uses a one-GPU box; `gpurun --gpus 2 -- 'STARWAY_TEST_MULTI_GPU=1 python -m pytest tests/test_gpu_multi.py -m gpu'`
runs it."""
import asyncio
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [1, 64, 8128, 8129, 65536 + 16, (1 << 20) + 3, 32 << 20]
ADDR = "127.0.0.1"


def _pattern(i, n):
    return ((np.arange(n, dtype=np.uint64) * 2654435761 + i * 131) >> 9).astype(np.uint8)


def _sender(port, device):
    os.environ["STARWAY_DEVICE"] = str(device)
    os.environ["STARWAY_QUIET"] = "1"
    import torch

    torch.cuda.set_device(device)
    import starway_b200 as sw

    async def inner():
        client = sw.Client()
        await client.aconnect(ADDR, port)
        bufs = [torch.from_numpy(_pattern(i, n)).cuda(device) for i, n in enumerate(SIZES)]
        torch.cuda.synchronize()
        verdict = torch.zeros(1, dtype=torch.uint8, device=f"cuda:{device}")
        fv = client.arecv(verdict, 0x77, (1 << 64) - 1)
        for rnd in range(3):
            for i, b in enumerate(bufs):
                await client.asend(b, 100 * rnd + i)
            await client.aflush()
        assert await asyncio.wait_for(fv, 120) == (0x77, 1)
        await client.aclose()

    asyncio.run(inner())
    sw.shutdown()


@pytest.mark.skipif(os.environ.get("STARWAY_TEST_MULTI_GPU") != "1", reason="opt-in: STARWAY_TEST_MULTI_GPU=1")
def test_two_gpus_bit_exact(port):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    os.environ.setdefault("STARWAY_DEVICE", "0")
    torch.cuda.set_device(0)
    import starway_b200 as sw

    async def go():
        server = sw.Server()
        server.listen(ADDR, port)
        connected = asyncio.Event()
        loop = asyncio.get_running_loop()
        server.set_accept_cb(lambda _: loop.call_soon_threadsafe(connected.set))
        child = mp.get_context("spawn").Process(target=_sender, args=(port, 1))
        child.start()
        try:
            await asyncio.wait_for(connected.wait(), 180)
            ep = next(iter(server.list_clients()))
            for rnd in range(3):
                dsts = [torch.full((n + 32,), 0xEE, dtype=torch.uint8, device="cuda:0") for n in SIZES]
                torch.cuda.synchronize()
                futs = [server.arecv(d, 100 * rnd + i, (1 << 64) - 1) for i, d in enumerate(dsts)]
                for i, (f, d, n) in enumerate(zip(futs, dsts, SIZES)):
                    assert await asyncio.wait_for(f, 120) == (100 * rnd + i, n)
                    torch.cuda.synchronize()
                    got = d.cpu().numpy()
                    np.testing.assert_array_equal(got[:n], _pattern(i, n))
                    assert (got[n:] == 0xEE).all()
            await server.asend(ep, torch.ones(1, dtype=torch.uint8, device="cuda:0"), 0x77)
            await server.aflush()
            for _ in range(600):
                if not child.is_alive():
                    break
                await asyncio.sleep(0.1)
            assert child.exitcode == 0
        finally:
            if child.is_alive():
                child.kill()
            child.join()
        await server.aclose()

    asyncio.run(asyncio.wait_for(go(), 600))
