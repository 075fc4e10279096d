"""The reference-shaped benchmark CLI (starway_b200/bench_cli.py; reference src/starway/bench.py +
benchmarks/scenarios.py): scenario bodies + READY/DONE control channel, run on the CPU simulator
with tiny sizes."""
import asyncio

from starway_b200 import bench_cli as bc


def test_scenarios_loopback_on_sim(sim_api):
    plan = [
        ("large-array", {"message_bytes": 3 << 20, "warmup": 1, "iterations": 2}),
        ("small-messages", {"message_bytes": 1024, "concurrency": 16, "warmup_batches": 1, "iterations": 3}),
        ("pingpong-flag", {"warmup": 5, "iterations": 20}),
        ("streaming-duplex", {"message_bytes": 1 << 20, "warmup": 1, "iterations": 4}),
    ]
    bufs = bc.Bufs("host")

    async def go():
        server, client = sim_api.Server(), sim_api.Client()
        await client.aconnect_address(server.listen_address())
        for _ in range(400):
            if server.list_clients():
                break
            await asyncio.sleep(0.005)
        ep = next(iter(server.list_clients()))
        srv = asyncio.ensure_future(bc.run_server_side(server, ep, plan, bufs))
        results = await bc.run_client_side(client, plan, bufs)
        await srv
        await client.aclose()
        await server.aclose()
        return results

    results = asyncio.run(asyncio.wait_for(go(), 120))
    assert [r["name"] for r in results] == [p[0] for p in plan]
    assert results[0]["metrics"]["avg_gbps"] > 0
    assert results[1]["metrics"]["messages_per_second"] > 0
    assert results[2]["metrics"]["median_rtt_us"] > 0
    assert results[3]["metrics"]["aggregate_gbps"] > 0


def test_cli_parsing():
    assert bc.parse_size("4MiB") == 4 << 20 and bc.parse_size("1g") == 1 << 30 and bc.parse_size("512") == 512
    assert bc.list_scenarios() == ["large-array", "small-messages", "pingpong-flag", "streaming-duplex"]


# the keys of the reference's JSON report (reference src/starway/bench.py:383-405, benchmarks/scenarios.py:49-57,
# 109-120, 184-196, 261-272, 319-336)
REFERENCE_METRICS = {
    "large-array": {"total_seconds", "avg_seconds_per_iter", "avg_gbps", "best_gbps", "worst_gbps"},
    "small-messages": {"total_seconds", "messages_per_second", "bandwidth_gbps", "latency_p50_us", "latency_p95_us"},
    "pingpong-flag": {"avg_rtt_us", "median_rtt_us", "min_rtt_us", "max_rtt_us", "avg_one_way_us"},
    "streaming-duplex": {"total_seconds", "avg_seconds_per_iter", "client_to_server_gbps", "server_to_client_gbps", "aggregate_gbps"},
}
REFERENCE_SAMPLES = {
    "large-array": {"duration_seconds", "per_iter_gbps"},
    "small-messages": {"batch_duration_seconds", "avg_latency_seconds"},
    "pingpong-flag": {"rtt_seconds"},
    "streaming-duplex": {"iteration_seconds"},
}


def check_report(report, with_samples):
    assert {"timestamp", "transport", "scenarios"} <= set(report)
    assert [s["name"] for s in report["scenarios"]] == list(REFERENCE_METRICS)
    for s in report["scenarios"]:
        assert {"name", "metrics", "config"} <= set(s)
        assert REFERENCE_METRICS[s["name"]] <= set(s["metrics"]), (s["name"], sorted(s["metrics"]))
        assert all(isinstance(v, float) and v >= 0 for v in s["metrics"].values())
        assert ("samples" in s) == with_samples
        if with_samples:
            assert REFERENCE_SAMPLES[s["name"]] == set(s["samples"])
