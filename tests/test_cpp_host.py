"""The C++ host-side mirror (include/b200ring.hpp) over the C ABI: compile tests/cpp/ring_host_test.cpp
against libb200ring.so.  CPU: constructor validation only.  GPU: KAT round trip, plugin path, fused product."""
import os
import subprocess

import pytest

from conftest import ROOT
import b200ring


def _build(tmp_path, src="ring_host_test.cpp"):
    b200ring.load_library()
    exe = str(tmp_path / src.replace(".cpp", ""))
    lib = os.path.join(ROOT, "matrix-fhe-lattigo_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", src), "-o", exe,
                           "-L" + lib, "-lb200ring", "-Wl,-rpath," + lib])
    return exe


def test_cpp_host_mirror_builds_and_validates(tmp_path):
    r = subprocess.run([_build(tmp_path), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_all_kernel_families_driver(tmp_path):
    """tests/cpp/sanitize_driver.cpp: every kernel family once through the C ABI from C++ (round trips, fused product
    vs operator sequence, host-buffer pipeline, key switch).  Also the binary to put under compute-sanitizer where
    that tool is available (it is closed on the B200 pool of this build)."""
    r = subprocess.run([_build(tmp_path, "sanitize_driver.cpp")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_concurrent_callers_share_one_ring(tmp_path):
    """tests/cpp/threads_test.cpp: four host threads on ONE context (ring/ring.go:192-194 thread-safety contract);
    the per-context lock keeps each call's launches (shared scratch) together on the stream."""
    r = subprocess.run([_build(tmp_path, "threads_test.cpp")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
