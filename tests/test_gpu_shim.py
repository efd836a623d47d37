"""Drop-in check of the C++ Evaluator shim (include/seal_b200/evaluator.hpp): tests/cpp/shim_test.cpp links the
reference's own libseal and compares seal::Evaluator with seal_b200::Evaluator word for word (and exception type for
exception type).  The binary is built in the container that has /root/reference and travels to the GPU box."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "cpp", "_bin", "shim_test")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="tests/cpp/_bin/shim_test not built (needs the reference headers)")
def test_cpp_evaluator_shim_matches_reference():
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout


MULTI = os.path.join(HERE, "cpp", "_bin", "multi_test")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(MULTI), reason="tests/cpp/_bin/multi_test not built")
def test_cpp_multi_device_dispatcher():
    """sb200_group_* (C++ host dispatcher, one context + one host thread per device) == the single-device entry points; on a box
    with one GPU the group is two contexts on that GPU"""
    import torch

    ndev = torch.cuda.device_count()
    r = subprocess.run([MULTI, str(ndev)], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout
