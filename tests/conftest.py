import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "constrastive-predictive-coding-audio_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    torch.backends.cudnn.allow_tf32 = False                     # every comparison in here is against fp32 references
    torch.backends.cuda.matmul.allow_tf32 = False
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libcpc_b200.so exists (compiles with nvcc on first use; no GPU needed to build)."""
    sys.path.insert(0, PKG)
    import build
    return build.build()


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".json"):
        with open(path) as fh:
            return json.load(fh)
    return dict(np.load(path, allow_pickle=False))


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| over the finite entries; non-finite entries (log(0) = -inf on
    digital silence) must coincide exactly."""
    import torch
    a = torch.as_tensor(a).detach().to(torch.float64).cpu()
    b = torch.as_tensor(b).detach().to(torch.float64).cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    fin = torch.isfinite(b)
    if not bool(fin.all()):
        assert bool((torch.isfinite(a) == fin).all()), "non-finite pattern differs"
        assert bool((a[~fin] == b[~fin]).all()), "non-finite values differ"
        a, b = a[fin], b[fin]
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_err(mine, ref, floor=1e-4):
    """Relative gradient error with an absolute floor: parameters whose true gradient is zero (a conv bias
    feeding a train-mode BatchNorm) carry only rounding noise in both implementations."""
    import torch
    a = torch.as_tensor(mine).detach().to(torch.float64).cpu()
    b = torch.as_tensor(ref).detach().to(torch.float64).cpu()
    return float((a - b).norm() / max(float(b.norm()), floor))


def bn_shadowed_biases(keys):
    """Conv biases that feed a train-mode BatchNorm directly: their true gradient is exactly zero, so both
    implementations only hold rounding noise there and a relative comparison is meaningless."""
    keys = set(keys)
    out = set()
    for k in keys:
        if not k.endswith(".bias"):
            continue
        head, idx = k[:-len(".bias")].rsplit(".", 1)
        if idx.isdigit() and ("%s.%d.running_mean" % (head, int(idx) + 1)) in keys and \
                ("%s.%s.running_mean" % (head, idx)) not in keys:
            out.add(k)
    return out
