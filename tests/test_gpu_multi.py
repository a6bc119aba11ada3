"""fks_multi_*: the batched forward simulation over several GPUs behind the C ABI.  The records must not
depend on the number of devices: N-GPU results == 1-GPU results byte for byte (Philox noise keyed by the global
particle id; tapes sharded with the particles in injection mode)."""
import numpy as np
import pytest
import torch

from fast_kinematic_simulator_b200 import capi, simulator as S, workloads as W

import parity
from oracle import oracle_binding as OB

pytestmark = pytest.mark.gpu


def _ndev():
    return torch.cuda.device_count()


def test_multi_with_one_device_equals_the_single_device_simulator():
    w = W.arm_table(1000)  # ragged against any device count
    single = w.make_simulator().forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    m = S.MultiGpuParticleContactSimulator(w.environment(), w.robot, 1, prng_seed=W.PRNG_SEED)
    multi = m.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    assert np.array_equal(single.records, multi.records)
    st = m.get_statistics()
    assert st["total_microsteps"] == int(multi.n_microsteps.sum())


@pytest.mark.skipif(_ndev() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name,n", [("arm_table", 4099), ("se3_narrow_passage", 2048), ("se2_arena", 130)])
def test_n_gpu_records_equal_one_gpu_records(name, n):
    w = W.make(name, n_particles=n)
    single = w.make_simulator().forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    for g in sorted({2, _ndev()}):
        m = S.MultiGpuParticleContactSimulator(w.environment(), w.robot, g, prng_seed=W.PRNG_SEED)
        multi = m.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
        assert np.array_equal(single.records, multi.records), (name, g)
        st = m.get_statistics()
        assert st["total_microsteps"] == int(single.n_microsteps.sum())
        assert st["total_resolver_iterations"] == int(single.n_resolver_iters.sum())
        m.close()


@pytest.mark.skipif(_ndev() < 2, reason="needs at least 2 GPUs")
def test_injection_mode_shards_the_tapes():
    n = 512
    w = W.arm_table(n)
    orc = parity.make_oracle(w)
    OB.lib().oracle_set_decision_cond_limit(orc._h, 100.0)
    ref, tape, sens = OB.run_with_tape(orc, w.starts, w.targets, True)
    m = S.MultiGpuParticleContactSimulator(w.environment(), w.robot, 2, prng_seed=W.PRNG_SEED)
    multi = m.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_INJECTED, tape)
    single = w.make_simulator().forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_INJECTED, tape)
    assert np.array_equal(single.records, multi.records)
    rep = parity.compare(multi, ref, sens, with_decisions=True)
    assert rep["discrete_ok"].all() and rep["n_match"] >= 0.99 * n


@pytest.mark.skipif(_ndev() < 2, reason="needs at least 2 GPUs")
def test_device_resident_results_are_all_gathered_on_every_device():
    g = _ndev()
    n = 1024 * g
    w = W.arm_table(n)
    single = w.make_simulator().forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    m = S.MultiGpuParticleContactSimulator(w.environment(), w.robot, g, prng_seed=W.PRNG_SEED)
    per = n // g
    ds, dt, dr = [], [], []
    for d in range(g):
        dev = torch.device("cuda", d)
        ds.append(torch.from_numpy(w.starts[d * per:(d + 1) * per].copy()).to(dev))
        dt.append(torch.from_numpy(w.targets.copy()).to(dev))
        dr.append(torch.zeros(n * m.result_stride, dtype=torch.uint8, device=dev))
    m.forward_simulate_device(ds, dt, n, 1, dr)
    for d in range(g):
        rec = dr[d].cpu().numpy().view(m.dtype)
        assert np.array_equal(rec, single.records), d


# ---- one process per GPU: fast_kinematic_simulator_b200/distributed.py over NCCL (the torchrun deployment) ---------------
def _nccl_worker(rank, world, port, n, q):
    import os
    import sys

    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from fast_kinematic_simulator_b200 import capi, workloads as W
    from fast_kinematic_simulator_b200.distributed import ShardedForwardSimulator, reduce_statistics

    w = W.arm_table(n)
    sim = w.make_simulator(device=rank)

    def shard(starts, targets, first_id):
        return sim.forward_simulate_robots(starts, targets, True, capi.NOISE_PHILOX, first_particle_id=first_id).records

    sh = ShardedForwardSimulator(shard, sim.result_stride, torch.device("cuda", rank))
    out = sh.forward_simulate_robots(w.starts, w.targets)
    s = sim.get_statistics()
    stats = reduce_statistics([s[k] for k in capi.STAT_NAMES], torch.device("cuda", rank))
    if rank == 0:
        q.put((out.cpu().numpy().tobytes(), stats.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_per_gpu_over_nccl_matches_one_gpu():
    """ShardedForwardSimulator (ragged contiguous shards, all_gather_into_tensor over NCCL, all-reduced counters) on two
    ranks == one GPU, byte for byte."""
    import os

    import torch.multiprocessing as mp

    n, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    raw, stats = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = W.arm_table(n)
    sim = w.make_simulator()
    ref = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    assert raw == ref.records.tobytes()
    s = sim.get_statistics()
    assert stats == [s[k] for k in capi.STAT_NAMES]
