"""The multi-rank parity scenarios of tests/multi_gpu_worker.py (BASELINE configs 2-5 vs the oracle) on the
CPU stand-in: world sizes 2 and 3, one process per rank, 'device' buffers of the stand-in.  Covers the
host-side protocol of the N > 1 path (and the scenario code itself) where there is no GPU."""
import pytest

from tests.test_gpu_multi import ALL, run_world


@pytest.mark.parametrize("world", [2, 3])
def test_configs_2_to_5_vs_oracle_on_the_stand_in(world):
    run_world(world, ALL, backend="sim", scale=1 / 64, timeout=600)
