"""pytest configuration: `gpu` marker, repo root on sys.path, in-tree build of the C-ABI library."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def native_lib():
    """libcasq_b200.so, built on demand (nvcc cross-compiles sm_100a without a GPU)."""
    from casq_b200 import _build, _native

    if not _native.lib_path().exists():
        _build.build()
    return _native.load()


@pytest.fixture(scope="session")
def manila_q0():
    """(oracle model, static, operators, channels) for qubit 0 of ibmq_manila in the H0 frame."""
    import oracle as O

    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    return O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=st), st, ops, ch


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
