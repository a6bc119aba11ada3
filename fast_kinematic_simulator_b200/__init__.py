"""B200-native batched particle forward-simulator.

Drop-in replacement for ONE path of calderpg/fast_kinematic_simulator: the batched
``SimpleParticleContactSimulator::ForwardSimulateRobots`` call
(include/fast_kinematic_simulator/simple_particle_contact_simulator.hpp:788-804) and everything it
runs per particle.  The product is ``libfksgpu.so`` (hand-written sm_100a kernels behind the C ABI of
``include/fksgpu.h``); this package is the thin Python host side used by the tests and ``bench.py``.
There is no CPU fallback: importing :mod:`fast_kinematic_simulator_b200.capi` raises if the library
is missing, and every compute call raises :class:`FksError` when no B200 is present.
"""
import importlib

_CAPI = ("FksError", "lib", "library_path", "default_solver_params", "SolverParams", "NOISE_PHILOX", "NOISE_INJECTED", "NOISE_NONE",
         "ROBOT_SE2", "ROBOT_SE3", "ROBOT_LINKED")
_SIMULATOR = ("BuiltEnvironment", "build_complete_environment", "RobotDescription", "GpuEnvironment", "GpuRobot",
              "GpuParticleContactSimulator", "MultiGpuParticleContactSimulator", "SimulationResults", "make_se2_simulator",
              "make_se3_simulator", "make_linked_simulator")
__all__ = list(_CAPI + _SIMULATOR)


def __getattr__(name):
    # resolved on first use, so that `fast_kinematic_simulator_b200.workloads` / `.abi` (pure descriptions) can be imported by
    # the CPU reference arm without loading libfksgpu.so; anything that computes still raises when the library is missing
    if name in _CAPI:
        return getattr(importlib.import_module(".capi", __name__), name)
    if name in _SIMULATOR:
        return getattr(importlib.import_module(".simulator", __name__), name)
    raise AttributeError(name)
