"""The CUDA library inside the real decision chain -- see tests/chain_on_device.py.  Runs last and in a subprocess so that nothing it
does to the CUDA context can touch the other GPU tests."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SCRIPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chain_on_device.py")


def _run(*args):
    return subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=1200)


@pytest.mark.parametrize("args", [(), ("--all-inputs", "--more")])
def test_chain_plumbing_with_a_cpu_stand_in_for_the_device(args):
    """same driver, callbacks and picture handling as the GPU tests, the device context replaced by oracle calls: pins the plumbing
    (and, with oracle/_ref present, ends in a byte-identical bitstream) on machines without a GPU"""
    r = _run("--stand-in", *args)
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


@pytest.mark.gpu
@pytest.mark.parametrize("args", [("--all-inputs",), ("--all-inputs", "--more")])
def test_cuda_operators_and_their_inputs_inside_the_decision_chain(args):
    """as above, with the inputs of the analyses from the device too -- xb200_mvp (MV predictor candidates, temporal direct MVs) and
    xb200_intra_nbr (availability, reference samples, MPM list) -- so every device row of SURVEY 8 takes part; --more: 10-bit medium,
    P slices, plain quantiser.
    First hardware run of the live parts: round 2 (GPUTEST_r01 / profiles/r02s01_gpu_tests.log, 46 passed); a hard assertion since."""
    r = _run(*args)
    assert r.returncode == 0 and "CHAIN_ON_DEVICE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    print(r.stdout)


def test_picture_dag_over_two_ranks_with_stand_in_device_contexts():
    """tests/chain_on_device.py --dag under torchrun (gloo, 2 ranks, the device contexts replaced by the CPU stand-in): the script the
    GPU box runs with NCCL and one Hotpath per rank -- waves of the picture DAG, reference pictures broadcast after each wave and
    adopted by the other rank's context, every rank's pictures checked against the reference's"""
    from oracle import refharness as rh
    if not rh.available():
        pytest.skip("oracle/_ref (compiled reference) not built here")
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), SCRIPT, "--dag", "--stand-in"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "CHAIN_DAG_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0/2: decided [0, 1, 2, 4, 5, 8, 9, 10, 13, 16]" in r.stdout and "rank 1/2: decided [3, 6, 7, 11, 12, 14, 15]" in r.stdout
