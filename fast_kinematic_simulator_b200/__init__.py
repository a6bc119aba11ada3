"""B200-native batched particle forward-simulator.

Drop-in replacement for ONE path of calderpg/fast_kinematic_simulator: the batched
``SimpleParticleContactSimulator::ForwardSimulateRobots`` call
(include/fast_kinematic_simulator/simple_particle_contact_simulator.hpp:788-804) and everything it
runs per particle.  The product is ``libfksgpu.so`` (hand-written sm_100a kernels behind the C ABI of
``include/fksgpu.h``); this package is the thin Python host side used by the tests and ``bench.py``.
There is no CPU fallback: importing :mod:`fast_kinematic_simulator_b200.capi` raises if the library
is missing, and every compute call raises :class:`FksError` when no B200 is present.
"""
from .capi import (  # noqa: F401
    FksError,
    lib,
    library_path,
    default_solver_params,
    SolverParams,
    NOISE_PHILOX,
    NOISE_INJECTED,
    NOISE_NONE,
    ROBOT_SE2,
    ROBOT_SE3,
    ROBOT_LINKED,
)
from .simulator import (  # noqa: F401
    BuiltEnvironment,
    build_complete_environment,
    RobotDescription,
    GpuEnvironment,
    GpuRobot,
    GpuParticleContactSimulator,
    SimulationResults,
    make_se2_simulator,
    make_se3_simulator,
    make_linked_simulator,
)
