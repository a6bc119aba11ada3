"""the first consumer of a batch on the device -- particles split by did_contact and pairwise configuration
distances (ComputeConfigurationDistanceTo, the call at spcs.hpp:898) -- against the oracle's restatement, on end states that
never leave the GPU.  The split is index work: bit-exact.  Distances: the linked and SE2 formulas are sums, products and one
square root (bit-exact up to FMA contraction: 4 ulp), SE3 goes through acos of a clamped trace, where the rotation angle
of nearly-equal orientations amplifies k ulp of the trace to sqrt(2 k eps) ~ 1e-7 rad: absolute tolerance 2e-7 there."""
import numpy as np
import pytest
import torch

from fast_kinematic_simulator_b200 import capi, workloads as W
from oracle import oracle_binding as OB

import parity

pytestmark = pytest.mark.gpu


def _run_on_device(w, n):
    sim = w.make_simulator()
    dev = torch.device("cuda")
    ds, dt = torch.from_numpy(w.starts).to(dev), torch.from_numpy(w.targets).to(dev)
    dr = torch.empty(n * sim.result_stride, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    sim.forward_simulate_device(ds, dt, n, w.targets.shape[0], dr, True, capi.NOISE_PHILOX, stream=st.cuda_stream)
    return sim, dr, st


@pytest.mark.parametrize("name,n,m,atol", [("arm_table", 3000, 700, 0.0), ("se2_arena", 1500, 600, 0.0), ("se3_narrow_passage", 4096, 500, 2e-7)])
def test_partition_and_distances_match_the_oracle(name, n, m, atol):
    w = W.make(name, n_particles=n)
    sim, dr, st = _run_on_device(w, n)
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    n_free, n_contact = sim.end_states_partition(dr, n, order, stream=st.cuda_stream)
    rec = dr.cpu().numpy().view(sim.dtype)
    ref_order, ref_free, ref_contact = OB.end_states_partition(rec["flags"])
    assert (n_free, n_contact) == (ref_free, ref_contact) and n_free + n_contact == n
    assert np.array_equal(order.cpu().numpy().view(np.uint32), ref_order)
    if name != "arm_table":
        assert 0 < n_contact  # the workloads are contact workloads (arm_table: every particle)
    # distances inside the bigger part, on a subset that is given by index
    part = ref_order[:n_free] if n_free >= n_contact else ref_order[n_free:]
    subset = np.ascontiguousarray(part[:: max(1, len(part) // m)][:m])
    d_subset = torch.from_numpy(subset.view(np.int32)).cuda()
    out = torch.empty(len(subset), len(subset), dtype=torch.float64, device="cuda")
    sim.end_states_pairwise_distance(dr, len(subset), out, d_subset, stream=st.cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    orc = parity.make_oracle(w)
    want = orc.pairwise_config_distance(rec["cfg"][subset])
    assert np.all(np.isfinite(got)) and np.all(got >= 0.0)
    assert np.allclose(got, want, rtol=1e-15 * 4, atol=atol), float(np.max(np.abs(got - want)))
    assert np.all(np.diag(got) <= (atol if atol else 0.0))
    # no subset: records 0 .. k-1
    k = 64
    out2 = torch.empty(k, k, dtype=torch.float64, device="cuda")
    sim.end_states_pairwise_distance(dr, k, out2, None, stream=st.cuda_stream)
    torch.cuda.synchronize()
    assert np.allclose(out2.cpu().numpy(), orc.pairwise_config_distance(rec["cfg"][:k]), rtol=4e-15, atol=atol)
    sim.close()


def test_partition_sizes_and_edge_cases():
    w = W.make("se3_narrow_passage", n_particles=2500)  # not a multiple of the 1024-record blocks
    sim, dr, st = _run_on_device(w, 2500)
    rec = dr.cpu().numpy().view(sim.dtype)
    for n in (1, 31, 1024, 1025, 2500):
        order = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        a, b = sim.end_states_partition(dr, n, order, stream=st.cuda_stream)
        ref_order, ra, rb = OB.end_states_partition(rec["flags"][:n])
        assert (a, b) == (ra, rb) and np.array_equal(order.cpu().numpy().view(np.uint32), ref_order)
    assert sim.end_states_partition(dr, 0, None) == (0, 0)
    sim.end_states_pairwise_distance(dr, 0, None)
    sim.close()
