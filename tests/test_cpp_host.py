"""Builds and runs the C++ host layer's tests (tests/cpp/test_deepq.cpp, the C++ twin of the
reference's memory_test.go / policy_test.go) against libgoldq.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_deepq")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "gold_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "gold_b200", "host"), os.path.join(ROOT, "tests", "cpp", "test_deepq.cpp"),
           "-L" + lib, "-lgoldq", "-Wl,-rpath," + lib, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_host_layer_cpu_parts():
    _build()
    r = subprocess.run([EXE, "--cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok TestMemory TestSchedule" in r.stdout


@pytest.mark.gpu
def test_cpp_host_layer_on_device():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok TestPolicy TestAgentLoop" in r.stdout
