"""The restated-reference CPU engine (oracle/cpu_engine.cpp) must itself satisfy the reference's
test assertions — it is the timed CPU baseline, so it has to be a correct one."""
import asyncio

import numpy as np
import pytest

from oracle import starway_cpu as cpu


def run(coro):
    return asyncio.run(asyncio.wait_for(coro, 60))


@pytest.mark.parametrize("size", [1, 10, 4096, 8192, 8193, 1 << 20])
def test_roundtrip(port, size):
    async def go():
        s, c = cpu.make_pair(port)
        await c.aconnect("127.0.0.1", port)
        send = np.random.randint(0, 256, size, dtype=np.uint8)
        recv = np.zeros(size, dtype=np.uint8)
        f = s.arecv(recv, 1, 0xFFFF)
        await c.asend(send, 1)
        assert await f == (1, size)
        np.testing.assert_array_equal(send, recv)
        recv.fill(0)
        snd = s.asend(0, send, 0x20002)           # unexpected first (a rendezvous send completes at match time)
        await asyncio.sleep(0.01)
        assert await c.arecv(recv, 2, 0xFFFF) == (0x20002, size)
        await snd
        np.testing.assert_array_equal(send, recv)
        await asyncio.gather(c.aflush(), s.aflush())
        pending = c.arecv(recv, 999, (1 << 64) - 1)
        await c.aclose()
        with pytest.raises(Exception, match="cancel"):
            await pending
        await s.aclose()

    run(go())


def test_many_messages_tagset(port):
    async def go():
        s, c = cpu.make_pair(port)
        await c.aconnect("127.0.0.1", port)
        n = 500
        sends = [c.asend(np.array([i]), i) for i in range(n)]
        recvs = [s.arecv(np.zeros(1, dtype=np.uint8), 0, 0) for _ in range(n)]
        res = await asyncio.gather(*sends, *recvs)
        assert {r[0] for r in res if r is not None} == set(range(n))
        await c.aclose()
        await s.aclose()

    run(go())
