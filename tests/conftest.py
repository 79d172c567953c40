import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a B200 skips the gpu tests instead of erroring (the GPU box selects them with -m gpu)"""
    try:
        import torch
        have = torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] >= 10
    except Exception:  # noqa: BLE001
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs an sm_100 GPU")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def trace():
    """CIF work lists traced from the reference (live when oracle/_ref is present, else golden)."""
    import tracedata
    return tracedata.get_trace()


@pytest.fixture(scope="session")
def golden():
    import tracedata
    return tracedata.golden_trace()


@pytest.fixture(scope="session")
def gpu_ctx(trace):
    """Device context + the trace's pictures uploaded (active area; padding done on the device)."""
    import numpy as np
    from xeve_b200 import api
    hp = api.Hotpath(trace.seq)
    handles = []
    for i, p in enumerate(trace.pics):
        h = hp.pic_create(padded=int(p["kind"]) == 1)
        y, u, v = (np.ascontiguousarray(a) for a in trace.planes[i])
        hp.pic_upload_s16(h, y, u, v)
        handles.append(h)
    hp.handles = np.array(handles, np.int32)
    yield hp
    hp.close()
