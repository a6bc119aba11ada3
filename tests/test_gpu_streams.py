"""Stream and lifetime semantics of the C ABI (include/fksgpu.h): batch calls of ONE simulator issued on different streams are
serialised by the library and give the results of serial calls; two live simulators of the same robot kind with robots of
different size do not disturb each other (the kernel's shared-memory attribute is per function and per device)."""
import numpy as np
import pytest
import torch

from fast_kinematic_simulator_b200 import capi, workloads as W

pytestmark = pytest.mark.gpu


def test_calls_on_different_streams_are_serialised():
    n = 3000
    w = W.arm_table(n)
    sim = w.make_simulator()
    serial_a = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX, first_particle_id=0).records.copy()
    serial_b = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX, first_particle_id=n).records.copy()
    dev = torch.device("cuda")
    ds, dt = torch.from_numpy(w.starts).to(dev), torch.from_numpy(w.targets).to(dev)
    out_a = torch.empty(n * sim.result_stride, dtype=torch.uint8, device=dev)
    out_b = torch.empty_like(out_a)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):  # back to back on two streams, then a host-buffer call on the simulator's own stream on top
        sim.forward_simulate_device(ds, dt, n, 1, out_a, True, capi.NOISE_PHILOX, first_particle_id=0, stream=sa.cuda_stream)
        sim.forward_simulate_device(ds, dt, n, 1, out_b, True, capi.NOISE_PHILOX, first_particle_id=n, stream=sb.cuda_stream)
        host = sim.forward_simulate_robots(w.starts[:500], w.targets, True, capi.NOISE_PHILOX, first_particle_id=0).records
        torch.cuda.synchronize()
        assert out_a.cpu().numpy().tobytes() == serial_a.tobytes()
        assert out_b.cpu().numpy().tobytes() == serial_b.tobytes()
        assert host.tobytes() == serial_a[:500].tobytes()
    sim.close()


def test_two_live_simulators_of_one_kind_with_different_robots():
    big, small = W.arm_table(600), W.gantry(600)  # both FKS_ROBOT_LINKED: 8 links x 48 points against a handful
    sim_big = big.make_simulator()
    ref_big = sim_big.forward_simulate_robots(big.starts, big.targets, True, capi.NOISE_PHILOX).records.copy()
    sim_small = small.make_simulator()             # created while the first one is alive
    ref_small = sim_small.forward_simulate_robots(small.starts, small.targets, True, capi.NOISE_PHILOX).records.copy()
    for _ in range(2):
        assert np.array_equal(sim_big.forward_simulate_robots(big.starts, big.targets, True, capi.NOISE_PHILOX).records, ref_big)
        assert np.array_equal(sim_small.forward_simulate_robots(small.starts, small.targets, True, capi.NOISE_PHILOX).records, ref_small)
    sim_small.close()
    assert np.array_equal(sim_big.forward_simulate_robots(big.starts, big.targets, True, capi.NOISE_PHILOX).records, ref_big)
    sim_big.close()
