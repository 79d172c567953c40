"""The CUDA library inside the real decision chain -- see tests/chain_on_device.py.  Runs last and in a subprocess so that nothing it
does to the CUDA context can touch the other GPU tests."""
import os
import subprocess
import sys

import pytest

SCRIPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chain_on_device.py")


def _run(*args):
    return subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=1200)


def test_chain_plumbing_with_a_cpu_stand_in_for_the_device():
    """same driver, callbacks and picture handling as the GPU test, the device context replaced by oracle calls: pins the plumbing
    (and, with oracle/_ref present, ends in a byte-identical bitstream) on machines without a GPU"""
    r = _run("--stand-in")
    assert r.returncode == 0 and "CHAIN_ON_DEVICE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_cuda_operators_inside_the_decision_chain():
    """every CU analysis of every tree node (xb200_analyze_cu, xb200_analyze_intra), the winner's prediction (xb200_mc) and every
    picture's loop filter + border expansion (xb200_deblock, device-resident references) computed by the library while the oracle's tree
    bookkeeping drives: pictures, maps and coder states equal the reference's and, injected into the unmodified reference, the decisions
    give the byte-identical bitstream (first hardware run: profiles/r01s21_chain_on_device.txt)."""
    r = _run()
    assert r.returncode == 0 and "CHAIN_ON_DEVICE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    print(r.stdout)
