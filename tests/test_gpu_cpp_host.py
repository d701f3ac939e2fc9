"""GPU: compiles and runs the C++ host-side mirror test (include/dockgpu.hpp over libdockgpu.so)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp_test():
    exe = os.path.join(ROOT, 'build', 'test_host_api')
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ['g++', '-std=c++17', '-O2', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'cpp', 'test_host_api.cpp'),
           '-L', os.path.join(ROOT, 'crypto_b200'), '-L', os.path.join(ROOT, 'oracle'), '-ldockgpu', '-lcpuref', '-lpthread',
           '-Wl,-rpath,' + os.path.join(ROOT, 'crypto_b200'), '-Wl,-rpath,' + os.path.join(ROOT, 'oracle'), '-o', exe]
    subprocess.check_call(cmd)
    return exe


def test_cpp_header_compiles(cref):
    """CPU: the header-only mirror and its test compile and link against both libraries."""
    build_cpp_test()


@pytest.mark.gpu
def test_cpp_host_api(cref, dg):
    exe = build_cpp_test()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'cpp host api ok' in out.stdout
