"""One rank of the multi-GPU parity tests (tests/test_gpu_multi.py spawns `world` of these, one per GPU).

Every scenario is a BASELINE.json configuration scaled to test size; every received buffer and every
(sender_tag, length) is checked against the CPU oracle (oracle/tagmatch.c through oracle.tagmatch.COracle):
the rank replays the events it took part in through the oracle -- receives in post order, arrivals in the
order the transport must preserve (per-sender FIFO; across senders the order it observed) -- and compares
the oracle's mirror buffers bit for bit with what landed in device memory.

  S2  config 2: 1 MiB device buffers, tag=1, tag_mask=0xFFFF over NVLink (ring r -> r+1), plus one 256 MiB message
  S3  config 3: sampled sizes 0 B .. 64 MiB, eager and rendezvous, expected and unexpected arrivals,
      random tags / masks / capacities (incl. truncation), both directions of every connection at once
  S4  config 4: all-pairs fan-out, wildcard receives, source identified from sender_tag, full payload compare
  S5  config 5: message storm, 64 B, tags (src << 32) | seq, Server.asend + aflush_ep per peer
"""
from __future__ import annotations

import asyncio
import os
import re
import traceback

import numpy as np

U64 = (1 << 64) - 1
ADDR = "127.0.0.1"
S3_SIZES = [0, 1, 64, 4096, 8128, 8129, 65536, 1 << 20, (1 << 20) + 3, 16 << 20, 64 << 20]
S3_DRAIN_CAP = 20 << 20   # wildcard receives that pick up leftovers: the 64 MiB message is truncated by them
T = 180                   # seconds: any single await


def payload(src: int, dst: int, k: int, n: int) -> np.ndarray:
    """Deterministic bytes of message k from rank src to rank dst (any rank can regenerate them)."""
    salt = (src * 1000003 + dst * 1009 + k * 31 + 7) & 0xFFFFFFFF
    if n <= (1 << 20):
        return np.random.default_rng(salt).integers(0, 256, n, dtype=np.uint8)
    x = np.arange(n, dtype=np.uint32)
    x *= np.uint32(2654435761)
    x += np.uint32(salt)
    x >>= np.uint32(13)
    return x.astype(np.uint8)


def s3_plan(src, dst, direction, seed, sizes=None):
    """Traffic from rank `src` to rank `dst` on one connection (direction 0: client -> server, 1: server ->
    client): receives (tag, mask, cap) in post order and sends (tag, size) in send order.  Both ends compute
    it from the seed."""
    rng = np.random.default_rng((seed * 7919 + src * 131 + dst * 17 + direction) & 0xFFFFFFFF)
    sizes = list(sizes or S3_SIZES)
    rng.shuffle(sizes)
    sends = [((int(rng.integers(0, 4)) | (int(rng.integers(0, 3)) << 8) | (direction << 20)), int(n)) for n in sizes]
    masks = [0, U64, 0xFF, 0xF00FF, 0xFFFF]
    recvs = []
    for tag, n in sends:
        mask = masks[int(rng.integers(0, len(masks)))]
        want = tag if rng.random() < 0.8 else (int(rng.integers(0, 4)) | (direction << 20))
        cap = n if rng.random() < 0.85 else n // 2   # some receives are too small: truncation
        recvs.append((want, mask, cap + int(rng.integers(0, 3)) * 16))
    order = rng.permutation(len(recvs))
    return [recvs[i] for i in order], sends


def s3_outcome(recvs, sends, unexpected_first):
    """What the oracle says about one stream without looking at payloads: (matched receive ops, leftover
    unexpected count).  The sender uses it to learn which of its peer's receives stay posted."""
    from oracle.tagmatch import COracle

    orc = COracle()
    mirrors = [np.zeros(cap, dtype=np.uint8) for _, _, cap in recvs]
    matched = set()
    if not unexpected_first:
        for op, (tag, mask, _) in enumerate(recvs):
            orc.post_recv(op, tag, mask, mirrors[op])
        for tag, n in sends:
            m = orc.arrive(0, tag, None, n)
            if m is not None:
                matched.add(m.op_id)
    else:
        for tag, n in sends:
            orc.arrive(0, tag, None, n)
        for op, (tag, mask, _) in enumerate(recvs):
            if orc.post_recv(op, tag, mask, mirrors[op]) is not None:
                matched.add(op)
    return matched, orc.num_unexpected


class CudaBufs:
    """Device buffers of the product path: torch CUDA tensors on this rank's GPU."""

    def __init__(self, rank):
        import torch

        self.torch, self.dev = torch, torch.device("cuda", rank)

    def from_np(self, a):
        return self.torch.from_numpy(a).to(self.dev)

    def fill(self, n):
        return self.torch.full((n,), 0xEE, dtype=self.torch.uint8, device=self.dev)

    def sync(self):
        self.torch.cuda.synchronize()

    @staticmethod
    def to_np(b):
        return b.cpu().numpy()

    @staticmethod
    def view(b, lo, hi):
        return b[lo:hi]


class SimBufs:
    """'Device' buffers of the CPU stand-in (tests/hostsim): the same scenarios on a machine without GPUs."""

    class View:
        def __init__(self, parent, lo, hi):
            self.parent, self.lo, self.hi = parent, lo, hi
            self.np = parent.np[lo:hi]

        @property
        def __cuda_array_interface__(self):
            return {"shape": (self.hi - self.lo,), "typestr": "|u1", "data": (self.parent.ptr + self.lo, False), "version": 2}

    def __init__(self, rank):
        from tests.hostsim import SimDev

        self.sd = SimDev

    def from_np(self, a):
        return self.sd.from_np(a)

    def fill(self, n):
        return self.sd.alloc(n)

    def sync(self):
        pass

    @staticmethod
    def to_np(b):
        return np.array(b.np, copy=True)

    @classmethod
    def view(cls, b, lo, hi):
        return cls.View(b, lo, hi)


class Rank:
    def __init__(self, rank, world, base_port, barrier, backend="cuda", scale=1.0):
        self.rank, self.world, self.base_port, self._barrier = rank, world, base_port, barrier
        self.peers = [p for p in range(world) if p != rank]
        self.nxt, self.prv = (rank + 1) % world, (rank - 1) % world
        self.backend, self.scale = backend, scale
        # the CPU stand-in runs the same scenarios with the largest sizes left out
        self.sizes = S3_SIZES if scale >= 1 else [n for n in S3_SIZES if n <= (1 << 20) + 3]
        self.drain_cap = S3_DRAIN_CAP if scale >= 1 else 600000   # still truncates the ~1 MiB leftovers

    async def barrier(self):
        await asyncio.get_running_loop().run_in_executor(None, self._barrier.wait, 600)

    async def setup(self):
        if self.backend == "cuda":
            import starway_b200 as sw

            self.bufs = CudaBufs(self.rank)
        else:
            from tests.hostsim import load

            sw = load()
            self.bufs = SimBufs(self.rank)
        self.sw = sw
        self.server = sw.Server()
        self.server.listen(ADDR, self.base_port + self.rank)
        await self.barrier()
        self.clients = {}
        for p in self.peers:
            c = sw.Client()
            await asyncio.wait_for(c.aconnect(ADDR, self.base_port + p), T)
            self.clients[p] = c
        for _ in range(12000):
            if len(self.server.list_clients()) >= len(self.peers):
                break
            await asyncio.sleep(0.005)
        eps = list(self.server.list_clients())
        assert len(eps) == len(self.peers)
        # the endpoint name carries the peer's GPU ordinal == its rank ("starway-ep-3[pid 1234 gpu 5]")
        self.ep_of = {int(re.search(r"gpu (\d+)\]", ep.name).group(1)): ep for ep in eps}
        assert sorted(self.ep_of) == self.peers, [ep.name for ep in eps]
        for ep in eps:
            assert ep.view_transports()[0][1] == "nvlink_ipc", ep.view_transports()   # another device, another process
        await self.barrier()

    # ------------------------------------------------------------------ helpers
    def dev_from(self, a: np.ndarray):
        return self.bufs.from_np(a)

    def dev_fill(self, n: int):
        return self.bufs.fill(n)

    async def expect(self, fut, m, what):
        """Await one receive and compare with the oracle's verdict for it."""
        from oracle.tagmatch import ORC_OK

        if m.status == ORC_OK:
            got = await asyncio.wait_for(fut, T)
            assert got == (m.sender_tag, m.length), (what, got, (m.sender_tag, m.length))
        else:
            try:
                await asyncio.wait_for(fut, T)
            except Exception as e:  # noqa: BLE001
                assert "truncated" in str(e), (what, e)
            else:
                raise AssertionError(f"{what}: expected a truncation error")

    def compare(self, bufs: dict, mirrors: dict, what):
        self.bufs.sync()
        for op, b in bufs.items():
            got = self.bufs.to_np(b)
            if not np.array_equal(got, mirrors[op]):
                bad = np.flatnonzero(got != mirrors[op])
                raise AssertionError(f"rank {self.rank} {what}: receive {op}: {bad.size} of {got.size} bytes differ "
                                     f"from the oracle, first at {bad[0]}")

    # ------------------------------------------------------------------ S2: BASELINE config 2
    async def s2_config2(self, count=8, n=1 << 20):
        from oracle.tagmatch import COracle

        orc = COracle()
        bufs = {k: self.dev_fill(n) for k in range(count)}
        mirrors = {k: np.full(n, 0xEE, dtype=np.uint8) for k in range(count)}
        self.bufs.sync()
        futs = {k: self.server.arecv(bufs[k], 1, 0xFFFF) for k in range(count)}
        for k in range(count):
            assert orc.post_recv(k, 1, 0xFFFF, mirrors[k]) is None
        await asyncio.sleep(0.05)
        await self.barrier()
        out = [self.dev_from(payload(self.rank, self.nxt, k + n, n)) for k in range(count)]
        self.bufs.sync()
        sends = [self.clients[self.nxt].asend(out[k], ((k + 1) << 16) | 1) for k in range(count)]
        for k in range(count):   # what rank-1 sent us, in its send order
            m = orc.arrive(0, ((k + 1) << 16) | 1, payload(self.prv, self.rank, k + n, n))
            assert m is not None and m.op_id == k
            await self.expect(futs[k], m, f"S2 message {k}")
        for f in sends:
            await asyncio.wait_for(f, T)
        await asyncio.wait_for(self.clients[self.nxt].aflush(), T)
        self.compare(bufs, mirrors, f"S2 n={n}")
        await self.barrier()

    # ------------------------------------------------------------------ S3: sizes, masks, both directions
    async def s3_sizes(self, seed, unexpected_first):
        from oracle.tagmatch import COracle

        # inbound streams: on our Server from prv's client (direction 0), on our client from nxt's Server (1)
        streams = []
        for worker, src, direction in ((self.server, self.prv, 0), (self.clients[self.nxt], self.nxt, 1)):
            recvs, sends = s3_plan(src, self.rank, direction, seed, self.sizes)
            streams.append(dict(worker=worker, src=src, direction=direction, recvs=recvs, sends=sends, orc=COracle(),
                                bufs={}, mirrors={}, futs={}, want={}))

        def data_of(src, dst, direction, k, n):
            return payload(src, dst, 100000 * seed + 10 * k + direction, n)

        def post_all():
            for st in streams:
                for op, (_, _, cap) in enumerate(st["recvs"]):
                    st["bufs"][op] = self.dev_fill(cap)
                    st["mirrors"][op] = np.full(cap, 0xEE, dtype=np.uint8)
            self.bufs.sync()
            for st in streams:
                for op, (tag, mask, _) in enumerate(st["recvs"]):
                    st["futs"][op] = st["worker"].arecv(st["bufs"][op], tag, mask)
                    m = st["orc"].post_recv(op, tag, mask, st["mirrors"][op])
                    if m is not None:
                        st["want"][op] = m

        def oracle_arrivals():
            for st in streams:
                for k, (tag, n) in enumerate(st["sends"]):
                    m = st["orc"].arrive(0, tag, data_of(st["src"], self.rank, st["direction"], k, n))
                    if m is not None:
                        st["want"][m.op_id] = m

        # outbound: our client -> nxt's Server (0), our Server -> prv's client (1)
        outbound = []
        for direction, dst in ((0, self.nxt), (1, self.prv)):
            recvs, sends = s3_plan(self.rank, dst, direction, seed, self.sizes)
            outbound.append((direction, dst, recvs, sends,
                             [self.dev_from(data_of(self.rank, dst, direction, k, n)) for k, (_, n) in enumerate(sends)]))
        self.bufs.sync()

        def send_to(direction, dst, buf, tag):
            return self.clients[dst].asend(buf, tag) if direction == 0 else self.server.asend(self.ep_of[dst], buf, tag)

        def send_all():
            return [send_to(direction, dst, outs[k], tag)
                    for direction, dst, _, sends, outs in outbound for k, (tag, _) in enumerate(sends)]

        if not unexpected_first:
            post_all()
            await asyncio.sleep(0.1)
            await self.barrier()
            sfuts = send_all()
            oracle_arrivals()
        else:
            sfuts = send_all()
            await asyncio.sleep(0.3)    # every message (or its RTS) now sits in the peer's unexpected queue
            await self.barrier()
            oracle_arrivals()
            post_all()
        what = f"S3 seed {seed} {'unexpected' if unexpected_first else 'expected'}"
        for st in streams:
            for op, m in st["want"].items():
                await self.expect(st["futs"][op], m, f"{what} dir {st['direction']} receive {op}")
        await asyncio.sleep(0.05)
        for st in streams:
            for op, f in st["futs"].items():
                assert op in st["want"] or not f.done(), f"{what}: receive {op} completed, the oracle left it pending"
            self.compare(st["bufs"], st["mirrors"], f"{what} dir {st['direction']}")
        await self.barrier()
        # ---- drain 1: messages no receive accepted are picked up by wildcard receives (earliest arrival first)
        for st in streams:
            n_unexp = st["orc"].num_unexpected
            base = len(st["recvs"])
            for i in range(n_unexp):
                op = base + i
                st["bufs"][op] = self.dev_fill(self.drain_cap)
                st["mirrors"][op] = np.full(self.drain_cap, 0xEE, dtype=np.uint8)
            self.bufs.sync()
            for i in range(n_unexp):
                op = base + i
                fut = st["worker"].arecv(st["bufs"][op], 0, 0)
                m = st["orc"].post_recv(op, 0, 0, st["mirrors"][op])
                assert m is not None
                await self.expect(fut, m, f"{what} dir {st['direction']} drain {i}")
            assert st["orc"].num_unexpected == 0
            self.compare({op: st["bufs"][op] for op in range(base, base + n_unexp)}, st["mirrors"], f"{what} drain")
        await self.barrier()
        # ---- drain 2: receives nothing matched get one zero-length filler each, in post order (the sender
        # runs the peer's oracle on the plan to learn which ones those are)
        empty = self.dev_from(np.zeros(0, dtype=np.uint8))
        for direction, dst, recvs, sends, _ in outbound:
            matched, _ = s3_outcome(recvs, sends, unexpected_first)
            for op, (tag, _, _) in enumerate(recvs):
                if op not in matched:
                    sfuts.append(send_to(direction, dst, empty, tag))
        for st in streams:
            for op, (tag, _, _) in enumerate(st["recvs"]):
                if op not in st["want"]:
                    m = st["orc"].arrive(0, tag, np.zeros(0, dtype=np.uint8))
                    assert m is not None and m.op_id == op, (what, op, m)
                    await self.expect(st["futs"][op], m, f"{what} filler {op}")
            assert st["orc"].num_posted == 0 and st["orc"].num_unexpected == 0
            self.compare(st["bufs"], st["mirrors"], f"{what} dir {st['direction']} after fillers")
        for f in sfuts:
            await asyncio.wait_for(f, T)
        await asyncio.wait_for(asyncio.gather(self.clients[self.nxt].aflush(), self.server.aflush_ep(self.ep_of[self.prv])), T)
        await self.barrier()

    # ------------------------------------------------------------------ S4: BASELINE config 4
    async def s4_allpairs(self, rounds=4, n=4 << 20):
        from oracle.tagmatch import COracle

        for rnd in range(rounds):
            orc = COracle()
            bufs = {i: self.dev_fill(n) for i in range(len(self.peers))}
            mirrors = {i: np.full(n, 0xEE, dtype=np.uint8) for i in bufs}
            out = {p: self.dev_from(payload(self.rank, p, 5000 + rnd, n)) for p in self.peers}
            self.bufs.sync()
            futs = {i: self.server.arecv(bufs[i], 0, 0) for i in bufs}          # wildcard: any sender, any tag
            for i in bufs:
                assert orc.post_recv(i, 0, 0, mirrors[i]) is None
            sends = [self.clients[p].asend(out[p], self.rank) for p in self.peers]   # tag = source rank
            res = {i: await asyncio.wait_for(futs[i], T) for i in bufs}
            for f in sends:
                await asyncio.wait_for(f, T)
            # The order of arrival across senders is not fixed; wildcard receives complete in post order, so
            # receive i holds the i-th arrival: replay that order through the oracle.
            assert sorted(t for t, _ in res.values()) == self.peers, res
            for i in bufs:
                src = res[i][0]
                m = orc.arrive(src, src, payload(src, self.rank, 5000 + rnd, n))
                assert m is not None and m.op_id == i and (m.sender_tag, m.length) == res[i]
            self.compare(bufs, mirrors, f"S4 round {rnd}")
            await self.barrier()

    # ------------------------------------------------------------------ S5: BASELINE config 5
    async def s5_storm(self, per_pair=3000, n=64):
        from oracle.tagmatch import COracle

        view = self.bufs.view
        # senders: our Server -> each of its endpoints; receivers: our clients (one sender each), wildcard
        slab = {p: self.dev_fill(per_pair * n) for p in self.peers}
        out = {p: self.dev_from(np.concatenate([payload(self.rank, p, 9000 + s, n) for s in range(per_pair)])) for p in self.peers}
        self.bufs.sync()
        views = {p: [view(slab[p], s * n, (s + 1) * n) for s in range(per_pair)] for p in self.peers}
        oviews = {p: [view(out[p], s * n, (s + 1) * n) for s in range(per_pair)] for p in self.peers}
        recvs = {p: [self.clients[p].arecv(views[p][s], 0, 0) for s in range(per_pair)] for p in self.peers}
        await asyncio.sleep(0.05)
        await self.barrier()
        sends = []
        for s in range(per_pair):
            for p in self.peers:
                sends.append(self.server.asend(self.ep_of[p], oviews[p][s], (self.rank << 32) | s))
        for f in sends:
            await asyncio.wait_for(f, T)
        for f in [self.server.aflush_ep(self.ep_of[p]) for p in self.peers]:
            await asyncio.wait_for(f, T)
        for p in self.peers:
            orc = COracle()
            mirror = np.full(per_pair * n, 0xEE, dtype=np.uint8)
            for s in range(per_pair):
                assert orc.post_recv(s, 0, 0, mirror[s * n:(s + 1) * n]) is None
            for s in range(per_pair):   # one sender per client worker: arrival order == send order
                m = orc.arrive(0, (p << 32) | s, payload(p, self.rank, 9000 + s, n))
                got = await asyncio.wait_for(recvs[p][s], T)
                assert m is not None and m.op_id == s and got == (m.sender_tag, m.length), (p, s, got)
            self.bufs.sync()
            assert np.array_equal(self.bufs.to_np(slab[p]), mirror), f"S5: payloads from rank {p} differ from the oracle"
        await self.barrier()

    async def teardown(self):
        await self.barrier()
        for c in self.clients.values():
            await asyncio.wait_for(c.aclose(), T)
        await self.barrier()
        await asyncio.wait_for(self.server.aclose(), T)


async def run_rank(r: Rank, scenarios):
    await r.setup()
    if "s2" in scenarios:
        await r.s2_config2()
        await r.s2_config2(count=1, n=int((256 << 20) * r.scale))
    if "s3" in scenarios:
        for seed, unexpected_first in ((1, False), (2, True), (3, False), (4, True)):
            await r.s3_sizes(seed, unexpected_first)
    if "s4" in scenarios:
        await r.s4_allpairs(n=max(65536, int((4 << 20) * r.scale)))
    if "s5" in scenarios:
        await r.s5_storm(per_pair=max(200, int(3000 * r.scale)))
    await r.teardown()


def rank_main(rank, world, base_port, barrier, queue, scenarios, backend="cuda", scale=1.0):
    os.environ["STARWAY_DEVICE"] = str(rank)
    os.environ["SW_SIM_DEVICE"] = str(rank)
    os.environ["SW_SIM_DEVICES"] = str(max(world, 2))
    os.environ["STARWAY_QUIET"] = "1"
    try:
        if backend == "cuda":
            import torch

            torch.cuda.set_device(rank)
            import starway_b200 as sw

            sw.bind_to_device_numa(rank)
        r = Rank(rank, world, base_port, barrier, backend, scale)
        asyncio.run(asyncio.wait_for(run_rank(r, scenarios), 1500))
        st = r.sw.get_context().stats()
        r.sw.shutdown()
        queue.put((rank, "ok", dict(st)))
    except BaseException:  # noqa: BLE001
        try:
            barrier.abort()
        except Exception:  # noqa: BLE001
            pass
        queue.put((rank, "error", traceback.format_exc()))
