import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def ref_backend_cls():
    from oracle.refmod import RefBackend, build
    build()
    return RefBackend


@pytest.fixture(scope="session")
def emu_lib():
    """g++ -DFHE_EMU build of the CUDA sources (functional simulator, tests only)."""
    from aes_fhe_b200 import build as b
    return b.build_emu()


@pytest.fixture(scope="session")
def cuda_lib():
    from aes_fhe_b200 import build as b
    return b.build_cuda()


def rand_poly(params, rng, npoly, nq, with_p, batch=1):
    """uniform residues [npoly, batch, limbs, N]"""
    ids = list(range(nq)) + ([params.n_q + k for k in range(params.n_p)] if with_p else [])
    a = np.empty((npoly, batch, len(ids), params.n), dtype=np.uint64)
    for p in range(npoly):
        for b in range(batch):
            for r, l in enumerate(ids):
                a[p, b, r] = rng.integers(0, params.moduli[l], size=params.n, dtype=np.uint64)
    return a


def make_engines(params, ref_backend_cls, gpu_backend, seed=3):
    """Two facades over the same parameter set and the same seed: identical keys."""
    from aes_fhe_b200.engine import Engine
    return (Engine(_params=params, _backend=gpu_backend, seed=seed),
            Engine(_params=params, _backend=ref_backend_cls(params), seed=seed))
