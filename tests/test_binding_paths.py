"""The binding layer has two implementations of the same steps (resolve buffer -> sw_post_* ->
resolve futures from sw_poll): the CPython fast path (starway_b200/csrc/fastpath.c) and the pure
ctypes code in _core.py.  The rest of the CPU suite runs on the fast path; this module re-runs a
representative subset on the ctypes fallback and checks fast-path specifics (buffer cache, explicit
loop argument, cancelled futures)."""
import asyncio

import numpy as np
import pytest

from tests import cases_basic as cb


@pytest.fixture(scope="module")
def ctypes_api():
    from tests.hostsim import load

    api = load(use_fastpath=False)
    assert api.get_context()._fp is None
    yield api
    api.shutdown()


def run(coro):
    return asyncio.run(asyncio.wait_for(coro, 120))


CASES = [
    cb.case_worker_address_connection_roundtrip, cb.case_client_to_server_send_recv, cb.case_multiple_clients,
    cb.case_concurrent_send_recv, cb.case_bidirectional_traffic, cb.case_rapid_connect_close_client,
    cb.case_shutdown_with_in_flight_ops, cb.case_double_close, cb.case_client_op_before_connect, cb.case_readme_quickstart,
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_ctypes_fallback(ctypes_api, port, case):
    run(case(ctypes_api, port))


@pytest.mark.parametrize("size", [1, 8129, 1 << 20])
def test_ctypes_fallback_integrity(ctypes_api, port, size):
    run(cb.case_message_integrity(ctypes_api, port, size))


def test_fastpath_is_active(sim_api):
    from starway_b200 import _core

    assert _core._fastpath is not None, "starway_b200/_fastpath.so missing: run `make lib`"
    assert sim_api.get_context()._fp is not None


def test_fastpath_buffer_cache_and_conversion(sim_api, port):
    async def go():
        async with cb.gen_server_client(sim_api, port) as (server, client):
            buf = np.zeros(8, dtype=np.uint8)
            src = np.arange(8, dtype=np.uint8)
            for i in range(5):  # same objects every time: cache hits after the first call
                src[:] = i
                f = server.arecv(buf, 0, 0)
                await client.asend(src, i)
                assert await f == (i, 8) and (buf == i).all()
            # int64 source is converted (reference nanobind behaviour): 1 byte on the wire
            f = server.arecv(buf, 0, 0)
            await client.asend(np.array([300]), 9)
            assert await f == (9, 1) and buf[0] == 300 & 0xFF
            # a read-only array cached by a send must still be refused as a receive buffer
            ro = np.arange(4, dtype=np.uint8)
            ro.flags.writeable = False
            f = server.arecv(buf, 0, 0)
            await client.asend(ro, 3)
            await f
            with pytest.raises(TypeError):
                server.arecv(ro, 0, 0)
            with pytest.raises(TypeError):
                server.arecv(np.zeros(4, dtype=np.int32), 0, 0)
            with pytest.raises(TypeError):
                client.asend("not a buffer", 0)

    run(go())


def test_cancelled_future_and_explicit_loop(sim_api, port):
    async def go():
        async with cb.gen_server_client(sim_api, port) as (server, client):
            buf = np.zeros(4, dtype=np.uint8)
            f = server.arecv(buf, 5, (1 << 64) - 1)
            f.cancel()  # the completion that arrives later must not blow up the drain
            await client.asend(np.arange(4, dtype=np.uint8), 5)
            await client.aflush()
            await asyncio.sleep(0.05)
            # explicit loop argument goes through the ctypes path of the same context
            f2 = server.arecv(buf, 6, (1 << 64) - 1, loop=asyncio.get_running_loop())
            await client.asend(np.arange(4, dtype=np.uint8), 6, loop=asyncio.get_running_loop())
            assert await f2 == (6, 4)

    run(go())


# ---------------------------------------------------------------- adaptive completion polling
def test_spin_disabled_uses_eventfd_only(sim_api, port):
    ctx = sim_api.get_context()
    saved, ctx._spin_ns = ctx._spin_ns, 0
    try:
        run(cb.case_concurrent_send_recv(sim_api, port))
        assert ctx._spin_loop is None
    finally:
        ctx._spin_ns = saved


def test_spin_times_out_and_eventfd_takes_over(sim_api, port):
    """A receive that stays pending longer than the polling budget must still complete (the loop has
    gone back to sleeping on the eventfd by then)."""
    async def go():
        ctx = sim_api.get_context()
        async with cb.gen_server_client(sim_api, port) as (server, client):
            buf = np.zeros(16, dtype=np.uint8)
            f = server.arecv(buf, 7, 0xFF)
            await asyncio.sleep(0.05)           # >> STARWAY_SPIN_US
            assert ctx._spin_loop is None       # polling has stopped, the op is still pending
            assert not f.done()
            await client.asend(np.full(16, 9, dtype=np.uint8), 7)
            assert await f == (7, 16) and (buf == 9).all()

    run(go())


def test_loop_stopped_while_polling(sim_api, port):
    """The loop returns while its polling callback is still queued; the next loop on the same context
    must receive its completions (suppressed wake-ups are re-enabled by the watchdog / new reader)."""
    ctx = sim_api.get_context()
    state = {}

    async def first():
        server = sim_api.Server()
        server.listen("127.0.0.1", port)
        client = sim_api.Client()
        await client.aconnect("127.0.0.1", port)
        buf = np.zeros(4, dtype=np.uint8)
        state.update(server=server, client=client, buf=buf, fut_loop=asyncio.get_running_loop())
        server.arecv(buf, 1, 0xFF)              # leaves an op pending -> the loop is polling when it returns
        assert ctx._spin_loop is not None or ctx._spin_ns <= 0

    async def second():
        server, client = state["server"], state["client"]
        buf2 = np.zeros(4, dtype=np.uint8)
        await client.asend(np.arange(4, dtype=np.uint8), 1)     # matches the receive posted by the first loop
        f = server.arecv(buf2, 2, 0xFF)
        await client.asend(np.arange(4, dtype=np.uint8) + 1, 2)
        assert await f == (2, 4)
        await client.aclose()
        await server.aclose()

    loop1 = asyncio.new_event_loop()
    try:
        loop1.run_until_complete(asyncio.wait_for(first(), 30))
    finally:
        pass  # loop1 is left un-closed on purpose (stopped, polling callback still queued)
    run(second())
    assert (state["buf"] == np.arange(4)).all()
    loop1.close()


def test_numa_helpers_are_noops_without_topology(sim_api):
    # the CPU stand-in backend has no PCI topology: nothing is known, nothing is changed
    import os

    before = os.sched_getaffinity(0)
    assert sim_api.local_cpus(0) == set()
    assert sim_api.bind_to_device_numa(0) is False
    assert os.sched_getaffinity(0) == before


@pytest.mark.parametrize("fast", [True, False], ids=["fastpath", "ctypes"])
def test_two_loops_in_two_threads_share_one_context(sim_api, ctypes_api, fast):
    """Either loop may drain completions that belong to the other (one completion queue per context):
    they are handed over with call_soon_threadsafe; batch buffers are per draining thread."""
    import threading

    from tests.conftest import free_port

    api = sim_api if fast else ctypes_api
    errors = []

    async def traffic(port, base):
        async with cb.gen_server_client(api, port) as (server, client):
            for rnd in range(30):
                bufs = [np.zeros(32, dtype=np.uint8) for _ in range(40)]
                recvs = [server.arecv(b, base + i, (1 << 64) - 1) for i, b in enumerate(bufs)]
                for i in range(40):
                    await client.asend(np.full(32, (base + i + rnd) & 0xFF, dtype=np.uint8), base + i)
                res = await asyncio.gather(*recvs)
                for i, (r, b) in enumerate(zip(res, bufs)):
                    assert r == (base + i, 32) and (b == ((base + i + rnd) & 0xFF)).all()

    def worker(base):
        try:
            asyncio.run(asyncio.wait_for(traffic(free_port(), base), 120))
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(1000 * (k + 1),)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(180)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads)


def test_uvloop_event_loop(sim_api):
    """bench.py and the scenario scripts run on uvloop: eventfd reader, polling callback and the fast
    path's future handling must behave the same there."""
    uvloop = pytest.importorskip("uvloop")
    from tests.conftest import free_port
    from tests.hostsim import SimDev

    for case in (cb.case_concurrent_send_recv, cb.case_bidirectional_traffic, cb.case_shutdown_with_in_flight_ops):
        uvloop.run(asyncio.wait_for(case(sim_api, free_port()), 120))
    uvloop.run(asyncio.wait_for(cb.case_chaos(sim_api, free_port(), 3), 120))
    uvloop.run(asyncio.wait_for(cb.case_random_schedule_vs_oracle(sim_api, free_port(), 4, SimDev), 120))


@pytest.mark.parametrize("fast", [True, False], ids=["fastpath", "ctypes"])
def test_second_thread_flush_close_and_raw_callbacks(sim_api, ctypes_api, fast):
    """Operations posted through the ctypes path (aflush, aflush_ep, aclose, aconnect, the raw callback
    forms) release the GIL inside sw_post_*: another loop's drain may poll their completion before the
    posting thread has registered the operation.  The completion must not be dropped (round-1 advisor
    finding: aflush from a second thread timed out on the fast path)."""
    import threading

    from tests.conftest import free_port

    api = sim_api if fast else ctypes_api
    errors = []
    stop = threading.Event()

    async def traffic(port):
        async with cb.gen_server_client(api, port) as (server, client):
            buf = np.zeros(32, dtype=np.uint8)
            src = np.full(32, 7, dtype=np.uint8)
            while not stop.is_set():
                f = server.arecv(buf, 5, (1 << 64) - 1)
                await client.asend(src, 5)
                await f

    async def flusher(port):
        loop = asyncio.get_running_loop()
        async with cb.gen_server_client(api, port) as (server, client):
            ep = next(iter(server.list_clients()))
            for i in range(400):
                await asyncio.wait_for(client.aflush(), 5)
                await asyncio.wait_for(server.aflush_ep(ep), 5)
                done = loop.create_future()
                client.flush(lambda: loop.call_soon_threadsafe(done.set_result, None),
                             lambda why: loop.call_soon_threadsafe(done.set_exception, Exception(why)))
                await asyncio.wait_for(done, 5)
                if i % 40 == 0:   # connect / close cycles post through the same path
                    c2 = api.Client()
                    await asyncio.wait_for(c2.aconnect("127.0.0.1", port), 5)
                    await asyncio.wait_for(c2.aclose(), 5)

    def worker(fn):
        try:
            asyncio.run(asyncio.wait_for(fn(free_port()), 120))
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))
        finally:
            stop.set()

    t1 = threading.Thread(target=worker, args=(traffic,))
    t2 = threading.Thread(target=worker, args=(flusher,))
    t1.start()
    t2.start()
    t2.join(150)
    stop.set()
    t1.join(30)
    assert not errors, errors
    assert not t1.is_alive() and not t2.is_alive()


def test_bound_post_callables(sim_api, port):
    """`Client.asend` / `arecv` and `Server.arecv` are C callables bound to the worker (fastpath.c Post): the plain call
    shape posts without a Python frame, every other shape (explicit loop, keywords) reaches the Python method, and the
    callable does not keep its object alive."""
    import gc
    import weakref

    from starway_b200 import _core

    post_type = type(sim_api.get_context()._fp.bound(0, 0, False, lambda *a, **k: None))

    async def go():
        async with cb.gen_server_client(sim_api, port) as (server, client):
            assert isinstance(client.asend, post_type) and isinstance(client.arecv, post_type) and isinstance(server.arecv, post_type)
            loop = asyncio.get_running_loop()
            src, dst = np.arange(64, dtype=np.uint8), np.zeros(64, dtype=np.uint8)
            f = server.arecv(dst, 7, 0xFFFF)
            await client.asend(src, 7)
            assert await f == (7, 64) and np.array_equal(src, dst)
            dst[:] = 0
            f = server.arecv(dst, 8, 0xFFFF, loop)            # explicit loop: Python method
            await client.asend(src, 8, loop=loop)              # keyword: Python method
            assert await f == (8, 64) and np.array_equal(src, dst)
            dst[:] = 0
            f = server.arecv(dst, tag=9, tag_mask=0xFFFF)
            await client.asend(buffer=src, tag=9)
            assert await f == (9, 64) and np.array_equal(src, dst)
            with pytest.raises(TypeError):
                client.asend(src)                              # wrong arity: the Python method's own error
            with pytest.raises((TypeError, ValueError, RuntimeError)):
                client.asend(object(), 1)                      # not a buffer

    run(go())
    c = sim_api.Client()
    ref = weakref.ref(c)
    del c
    gc.collect()   # (no cycle is needed to free it; collect only makes the assertion independent of frame locals)
    assert ref() is None
    assert _core._fastpath is not None
