"""The header-only C++ mirror of the reference interface (mimosa_b200/host/mimosa_b200.hpp) compiles and
links against the C-ABI library like a mimosa translation unit would; on a GPU box the program also runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_smoke.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_smoke")


def build():
    lib_dir = os.path.join(ROOT, "mimosa_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", SRC, "-I", os.path.join(ROOT, "mimosa_b200", "host"),
                    "-L", lib_dir, "-lmimosa_b200", f"-Wl,-rpath,{lib_dir}", "-o", EXE], check=True)


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
