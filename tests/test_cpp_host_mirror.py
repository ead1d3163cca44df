"""The C++ host-side mirror (include/aether_b200.hpp) re-states the reference crate's own unit tests
and doctests; CPU: it builds against the C ABI and fails loudly without a GPU; GPU: it passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "host_mirror_test")


def build():
    subprocess.check_call(["make", "-C", ROOT, "cpptest"], stdout=subprocess.DEVNULL)


def test_builds_and_fails_loudly_without_gpu():
    build()
    import aether_primitives_b200 as ae

    if ae.device_count() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode != 0
    assert "no CUDA device available" in (p.stderr + p.stdout)


@pytest.mark.gpu
def test_reference_tests_pass_through_the_cpp_mirror():
    build()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "host mirror ok" in p.stdout
