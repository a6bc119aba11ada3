"""world_size-2 gloo test of the N > 1 path (sharding + all-gather of end-state records) on CPU.  The shard
compute is the oracle in Philox mode (keyed by global particle id), standing in for the GPU simulator: the test
checks the plumbing -- contiguous ragged shards, first_particle_id offsets, gather order, statistics reduction."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fast_kinematic_simulator_b200 import capi, workloads as W
    from fast_kinematic_simulator_b200.distributed import ShardedForwardSimulator, reduce_statistics, shard_bounds
    import parity

    w = W.se3_narrow_passage(n)
    orc = parity.make_oracle(w, num_threads=1)

    def shard(starts, targets, first_id):
        return orc.forward_simulate(starts, targets, True, capi.NOISE_PHILOX, None, first_id)

    sh = ShardedForwardSimulator(shard, orc.dtype.itemsize, "cpu")
    out = sh.forward_simulate_robots(w.starts, w.targets)
    stats = reduce_statistics(orc.statistics())
    if rank == 0:
        q.put((out.numpy().tobytes(), stats.tolist(), shard_bounds(n, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    n, world = 37, 2  # ragged: 19 + 18
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    raw, stats, bounds = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert bounds == [0, 19, 37]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fast_kinematic_simulator_b200 import capi, workloads as W
    import parity

    w = W.se3_narrow_passage(n)
    orc = parity.make_oracle(w, num_threads=2)
    ref = orc.forward_simulate(w.starts, w.targets, True, capi.NOISE_PHILOX)
    got = np.frombuffer(raw, dtype=orc.dtype)
    assert np.array_equal(got, ref)
    assert stats == orc.statistics().tolist()


def test_shard_bounds():
    from fast_kinematic_simulator_b200.distributed import shard_bounds

    assert shard_bounds(8, 8) == list(range(9))
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(0, 2) == [0, 0, 0]
    assert shard_bounds(1048576, 8)[-1] == 1048576
