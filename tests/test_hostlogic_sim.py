"""CPU suite: the reference's integration tests (tests/cases_basic.py) driven through the
REAL host progress engine linked against the test-only device simulator (tests/hostsim).
Covers connection, eager / rendezvous protocol, credits, flush, close, cancellation and
world_size-2 (two processes) operation without a GPU."""
import asyncio

import pytest

from tests import cases_basic as cb


def run(coro):
    return asyncio.run(asyncio.wait_for(coro, timeout=120))


@pytest.mark.parametrize("case", cb.SINGLE_PROCESS_CASES, ids=lambda c: c.__name__)
def test_reference_case(sim_api, port, case):
    run(case(sim_api, port))


@pytest.mark.parametrize("size", [1, 1024, 4096, 8128, 8129, 65536, 1 << 20, (1 << 22) + 13])
def test_message_integrity(sim_api, port, size):
    run(cb.case_message_integrity(sim_api, port, size))


@pytest.mark.parametrize("mode", ["flush", "flush_ep"])
def test_two_process_server_send_with_flush_good(sim_api, port, mode):
    run(cb.case_server_send_with_flush_good(sim_api, port, "sim", mode))


def test_two_process_client_send_with_flush_good(sim_api, port):
    run(cb.case_client_send_with_flush_good(sim_api, port, "sim"))


@pytest.mark.parametrize("seed", range(12))
def test_random_schedule_vs_oracle(sim_api, port, seed):
    run(cb.case_random_schedule_vs_oracle(sim_api, port, seed, quiesce=0.002))
