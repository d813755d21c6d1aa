"""world_size-2 `gloo` test of the multi-GPU host logic on the CPU: contiguous blocks of the global
trajectory index per rank, disjoint Philox counters, sum-allreduce of [sum, sumsq, n] (and of tangent sums); the
reduced price must equal the single-process price over all trajectories up to floating-point summation order."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ranks(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gloo") / "res.json")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(HERE, "_gloo_worker.py"), out]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return [json.load(open(f"{out}.{r}")) for r in range(2)]


def test_both_ranks_report_the_same_reduced_results(ranks):
    for key in ("seeds", "basket", "greeks"):
        assert ranks[0][key]["sharded"] == ranks[1][key]["sharded"]
    for key in ("novr", "anti"):
        assert ranks[0][key]["sharded"][:2] == ranks[1][key]["sharded"][:2]


@pytest.mark.parametrize("key", ["novr", "anti"])
def test_sharded_european_price_equals_single_process(ranks, key):
    for r in ranks:
        (p, se, n_local, n_total), (p1, se1) = r[key]["sharded"], r[key]["single"]
        assert n_total == 20_001
        assert abs(p - p1) <= 1e-13 * abs(p1)
        assert abs(se - se1) <= 1e-10 * abs(se1)
    # uneven shards: floor(N r / W) .. floor(N (r+1) / W)
    assert [r[key]["sharded"][2] for r in ranks] == [10_000, 10_001]


def test_per_trajectory_seeds_are_sharded_with_the_trajectories(ranks):
    for r in ranks:
        assert abs(r["seeds"]["sharded"] - r["seeds"]["single"]) <= 1e-13 * abs(r["seeds"]["single"])


def test_strike_grid_and_batch_greeks_reduce_across_ranks(ranks):
    for r in ranks:
        np.testing.assert_allclose(r["basket"]["sharded"], r["basket"]["single"], rtol=1e-13)
        np.testing.assert_allclose(r["greeks"]["sharded"], r["greeks"]["single"], rtol=1e-11, atol=1e-13)


def test_lsm_comm_callback_sums_in_place(ranks):
    want = (np.arange(12, dtype=np.float64) * 3).tolist()  # rank 0: x1, rank 1: x2
    for i, r in enumerate(ranks):
        assert r["comm"] == {"rc": 0, "buf": want, "rank": i, "world": 2}
        assert r["allreduce"] == [[3.0, 4.0], [6.0, 4.0]]


def test_path_dependent_payoffs_reduce_across_ranks(ranks):
    assert ranks[0]["pathdep"]["sharded"] == ranks[1]["pathdep"]["sharded"]
    for r in ranks:
        np.testing.assert_allclose(r["pathdep"]["sharded"], r["pathdep"]["single"], rtol=1e-10)


def test_black_scholes_control_variate_reduces_across_ranks(ranks):
    assert ranks[0]["bs_control"]["sharded"] == ranks[1]["bs_control"]["sharded"]
    for r in ranks:
        np.testing.assert_allclose(r["bs_control"]["sharded"], r["bs_control"]["single"], rtol=1e-10)
