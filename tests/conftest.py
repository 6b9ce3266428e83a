import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity bar from BASELINE.json's north_star: fp32 within 1e-5 relative / 1e-6 absolute
RTOL = 1e-5
ATOL = 1e-6


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_close(got, want, rtol=RTOL, atol=ATOL, what=""):
    got = got.detach().cpu().double()
    want = want.detach().cpu().double()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    bad = err > bound
    if bad.any():
        i = torch.argmax(err / bound)
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} outside {rtol:g} rel / {atol:g} abs; "
            f"worst {float((err / bound).flatten()[i]):.2f}x bound: got {float(got.flatten()[i]):.9g} "
            f"want {float(want.flatten()[i]):.9g}"
        )


@pytest.fixture(scope="session")
def dev():
    return torch.device("cuda:0")
