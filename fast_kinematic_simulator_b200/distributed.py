"""Multi-GPU plumbing of the batched forward-simulate call: one process per GPU, particles sharded in
contiguous ranges, environment + robot replicated per GPU, ONE all-gather of end-state records per call
(SURVEY.md 8e).  There is no other exchange step: particles are independent
(simple_particle_contact_simulator.hpp:795-802), and Philox noise is keyed by GLOBAL particle id so the
result does not depend on the number of ranks.

The class is backend-agnostic (NCCL with CUDA tensors on the GPU box, gloo with CPU tensors in the tests);
`simulate_shard` is whatever computes one shard's records -- GpuParticleContactSimulator on a B200.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous ranges: the first (n % world) ranks get one extra particle."""
    base, extra = divmod(int(n), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


class ShardedForwardSimulator:
    def __init__(self, simulate_shard, record_bytes, device="cpu", group=None):
        """simulate_shard(starts, targets, first_particle_id) -> uint8 torch tensor [n_local * record_bytes] on `device`
        (or a numpy record array, which is wrapped)."""
        self.simulate_shard = simulate_shard
        self.record_bytes = int(record_bytes)
        self.device = torch.device(device)
        self.group = group

    def forward_simulate_robots(self, starts, targets):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = len(starts)
        if not (len(targets) == 1 or len(targets) == n):  # spcs.hpp:790-793
            raise ValueError("need 1 target or one per start")
        b = shard_bounds(n, world)
        lo, hi = b[rank], b[rank + 1]
        t = targets if len(targets) == 1 else targets[lo:hi]
        local = self.simulate_shard(starts[lo:hi], t, lo)
        if isinstance(local, np.ndarray):
            local = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1)).to(self.device)
        if world == 1:
            return local
        # equal-size all-gather over padded shards (ragged N is padded to the largest shard, then trimmed)
        widest = max(b[r + 1] - b[r] for r in range(world)) * self.record_bytes
        send = torch.zeros(widest, dtype=torch.uint8, device=self.device)
        send[: local.numel()] = local
        recv = torch.empty(world * widest, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        parts = [recv[r * widest: r * widest + (b[r + 1] - b[r]) * self.record_bytes] for r in range(world)]
        return torch.cat(parts)


def reduce_statistics(stats_vector, device="cpu", group=None):
    """Sum of the per-rank counters (GetStatistics keys + totals)."""
    t = torch.as_tensor(np.asarray(stats_vector, dtype=np.int64), device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
