"""The C++ call surface (include/lio_facade.hpp): a ROS-free host program written against KD_TREE / esekf compiles with
plain g++ -std=c++14 against the C-ABI library, fails loudly without a GPU, and recovers a known pose with one."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _build(tmp_path, name="replay_main"):
    exe = tmp_path / name
    pkg = ROOT / "agi_lidar_slam_b200"
    cmd = ["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / f"{name}.cpp"),
           f"-L{pkg}", "-llio_b200", f"-Wl,-rpath,{pkg}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_facade_compiles_links_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_recovers_pose(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr


def test_reference_call_expressions_compile(tmp_path):
    """examples/reference_calls.cpp: laserMapping.cpp's own call expressions on the path (:361-364, :430-431, :683,
    :737-738, :747-756, :771-774) and esekfom.hpp:300's h_share_model call, verbatim, against the facade."""
    import torch

    exe = _build(tmp_path, "reference_calls")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_call_expressions_behave(tmp_path):
    exe = _build(tmp_path, "reference_calls")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
