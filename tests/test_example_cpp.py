"""The C++ host example (examples/host_cpp/trace_example.cpp) builds against include/tracer_rq.h and drives the
library through the C-ABI only. CPU: it must refuse to run without a device (exit 3). GPU: it must trace."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "host_cpp", "trace_example.cpp")
EXE = os.path.join(ROOT, "examples", "host_cpp", "trace_example")


def _build():
    subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT}/include", SRC, f"-L{ROOT}/tracer_b200", "-ltracer_rq",
                    f"-Wl,-rpath,{ROOT}/tracer_b200", "-o", EXE], check=True)


def test_cpp_host_compiles_and_fails_loudly_without_gpu(built):
    from tracer_b200._lib import lib
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    if lib.trq_device_count() == 0:
        assert out.returncode == 3 and "no CPU fallback" in out.stderr
    else:
        assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_cpp_host_traces_on_gpu(built):
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "rays hit" in out.stdout and "A" in out.stdout and "." in out.stdout
