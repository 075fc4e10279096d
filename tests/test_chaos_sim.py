"""Chaos test of the host progress engine (CPU simulator): random mixes of sends (eager and
rendezvous), receives (wildcard / exact / never-matching), flushes and closes at random points,
on several clients.  Invariants: nothing hangs, every future resolves (result or one of the
documented errors), delivered payloads are intact, and aclose() always returns."""
import asyncio
import os

import numpy as np
import pytest

from tests import cases_basic as cb

@pytest.mark.parametrize("seed", range(int(os.environ.get("CHAOS_SEEDS", "12"))))
def test_chaos(sim_api, port, seed):
    asyncio.run(asyncio.wait_for(cb.case_chaos(sim_api, port, seed), 120))


@pytest.mark.parametrize("seed", range(1000, 1000 + int(os.environ.get("CHAOS_SEEDS", "12")) // 2))
def test_chaos_device_buffers(sim_api, port, seed):
    """Same, with the stand-in backend's 'device' buffers: closes and cancellations race with zero-copy
    rendezvous pulls out of user allocations."""
    from tests.hostsim import SimDev

    asyncio.run(asyncio.wait_for(cb.case_chaos(sim_api, port, seed, bufs=SimDev), 120))
