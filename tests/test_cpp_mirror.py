"""The header-only C++ mirror of the reference interface (mimosa_b200/host/mimosa_b200.hpp) compiles and
links against the C-ABI library like a mimosa translation unit would; on a GPU box the program also runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_smoke.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_smoke")


ADAPTER_SRC = os.path.join(ROOT, "tests", "cpp", "adapter_check.cpp")
ADAPTER_EXE = os.path.join(ROOT, "tests", "cpp", "adapter_check")


def build():
    lib_dir = os.path.join(ROOT, "mimosa_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", SRC, "-I", os.path.join(ROOT, "mimosa_b200", "host"),
                    "-L", lib_dir, "-lmimosa_b200", f"-Wl,-rpath,{lib_dir}", "-o", EXE], check=True)


def build_adapter():
    """mimosa::lidar::ICPFactorB200 : gtsam::NonlinearFactor (mimosa_b200/host/adapters/geometric_factor_b200.hpp) against
    the minimal GTSAM / PCL stand-ins of tests/cpp/stubs/."""
    lib_dir = os.path.join(ROOT, "mimosa_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", ADAPTER_SRC, "-I", os.path.join(ROOT, "mimosa_b200", "host"),
                    "-I", os.path.join(ROOT, "tests", "cpp", "stubs"), "-L", lib_dir, "-lmimosa_b200", f"-Wl,-rpath,{lib_dir}",
                    "-o", ADAPTER_EXE], check=True)


def test_gtsam_adapter_compiles_links_and_fails_loudly_without_gpu():
    import torch

    build_adapter()
    r = subprocess.run([ADAPTER_EXE], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 2 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_gtsam_adapter_runs_on_gpu():
    build_adapter()
    r = subprocess.run([ADAPTER_EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "adapter ok" in r.stdout, r.stdout + r.stderr


def test_cpp_mirror_compiles_links_and_fails_loudly_without_gpu():
    import torch

    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 2 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_runs_on_gpu():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
