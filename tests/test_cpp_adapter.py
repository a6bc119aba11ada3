"""The C++ host adapter (include/fksgpu_simulator.hpp) compiles with a plain C++11 compiler against the C ABI
and behaves like the reference interface: CPU box -> loud failure (no fallback), GPU box -> particles stop at the wall."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = "/tmp/fksgpu_se2_example"


def build():
    lib = os.path.join(ROOT, "fast_kinematic_simulator_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "forward_simulate_se2.cpp"), "-L" + lib, "-lfksgpu",
                           "-Wl,-rpath," + lib, "-o", EXE])


def test_adapter_compiles_and_fails_loudly_without_a_gpu():
    import torch

    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_adapter_on_gpu():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "ok" in r.stdout
