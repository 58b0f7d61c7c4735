"""Multi-process (world_size 2, gloo, CPU) test of the multi-GPU host logic: env sharding by global id
and the all-reduce of per-rank counter sums.  The per-rank "device" sums come from the oracle here
(the CUDA path is exercised by the -m gpu tests); what is under test is the plumbing of
optical_rl_gym_b200.sharding."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL_ENVS, STEPS, SEED = 12, 150, 9


def rank_sums(first, count):
    """What OpticalVecEnv.reduce_counters() returns on a rank that owns envs [first, first+count)."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "optical-rl-gym_b200"), os.path.join(ROOT, "tests")]
    import helpers
    from oracle import oracle

    sums = np.zeros(9, np.int64)
    for i in range(first, first + count):
        e = oracle.OracleEnv("DeepRMSA-v0", helpers.golden_tables(), num_slots=100, episode_length=40)
        e.set_philox(SEED, i)
        e.reset(full=True)
        e.rollout(STEPS, policy=1)
        sums[:8] += e.counters()
        sums[8] += int(e.error() != 0)
    return sums


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path[:0] = [os.path.join(ROOT, "optical-rl-gym_b200")]
    from optical_rl_gym_b200 import sharding

    first, count = sharding.shard_range(TOTAL_ENVS, rank, world)
    stats = sharding.global_statistics(torch.from_numpy(rank_sums(first, count)))
    out[rank] = (first, count, stats)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_job():
    sys.path[:0] = [os.path.join(ROOT, "optical-rl-gym_b200")]
    from optical_rl_gym_b200 import sharding

    for total in (1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_allreduce_equals_single_rank_totals():
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    whole = rank_sums(0, TOTAL_ENVS)
    assert res[0][0] == 0 and res[0][1] + res[1][1] == TOTAL_ENVS and res[1][0] == res[0][1]
    for r in (0, 1):
        s = res[r][2]
        assert s["services_processed"] == whole[0] and s["services_accepted"] == whole[1]
        assert s["bit_rate_requested"] == whole[4] and s["bit_rate_provisioned"] == whole[5]
        assert s["envs_with_errors"] == 0
        assert s["service_blocking_rate"] == (whole[0] - whole[1]) / whole[0]
