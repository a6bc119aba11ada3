"""The C-ABI library loads and exports every symbol include/fksgpu.h declares (no compute calls)."""
import ctypes as C
import os
import re

from fast_kinematic_simulator_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fksgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fks_[a-z0-9_]+)\s*\(", text)))


def test_header_functions_are_exported():
    names = declared_functions()
    assert len(names) >= 25
    for nm in names:
        assert hasattr(capi.lib, nm), "libfksgpu.so does not export %s" % nm
    assert sorted(capi.EXPORTS) == names


def test_abi_version_and_defaults():
    assert capi.lib.fks_abi_version() == 2
    p = capi.default_solver_params()  # SimulatorSolverParameters() defaults, spcs.hpp:357-368
    assert (p.forward_simulation_time, p.simulation_shortcut_distance, p.environment_collision_check_tolerance) == (1.0, 0.0, 0.001)
    assert (p.resolve_correction_step_scaling_decay_rate, p.resolve_correction_initial_step_size, p.resolve_correction_min_step_scaling) == (0.5, 1.0, 0.03125)
    assert (p.max_resolver_iterations, p.resolve_correction_step_scaling_decay_iterations, p.failed_resolves_end_motion) == (25, 5, 1)


def test_struct_sizes_match_header_layout():
    assert C.sizeof(capi.SolverParams) == 64
    assert C.sizeof(capi.AxisParams) == 64
    assert C.sizeof(capi.JointDesc) == 16 + 12 * 8 + 3 * 8 + 3 * 8
    assert C.sizeof(capi.Obstacle) == 12 * 8 + 3 * 8 + 8
    assert C.sizeof(capi.NoiseTape) == 32


def test_pure_python_defaults_equal_the_librarys():
    from fast_kinematic_simulator_b200 import abi

    a, b = abi.default_solver_params(), capi.default_solver_params()
    for name, _ in abi.SolverParams._fields_:
        assert getattr(a, name) == getattr(b, name), name


def test_workloads_import_without_the_product_library():
    """bench.py --impl reference describes robots / environments with the package's pure modules: importing them must not
    load libfksgpu.so (checked in a fresh interpreter)."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); import fast_kinematic_simulator_b200.workloads as W; w = W.arm_table(4); "
            "w.robot.to_c(); import ctypes; "
            "maps = open('/proc/self/maps').read(); assert 'libfksgpu' not in maps, 'product library loaded'; print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_no_device_is_an_error_not_a_fallback():
    """Without a GPU every compute entry point must fail loudly (there is no CPU path in the product)."""
    import torch

    if torch.cuda.is_available():
        return
    from fast_kinematic_simulator_b200 import workloads as W
    import pytest

    w = W.se2_arena(4)
    with pytest.raises(capi.FksError) as ei:
        w.make_simulator()
    assert ei.value.code in (capi.ERR_NO_DEVICE, capi.ERR_CUDA)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "fast_kinematic_simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_binding" not in text and "libfks_oracle" not in text and "fks_oracle" not in text, f
