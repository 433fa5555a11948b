import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STATE_JSON = os.path.join(ROOT, "tests", "golden", "hepem_state.json")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def flat_tables():
    from g4hepem_b200 import tables

    return tables.load_state_json(STATE_JSON)


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (built here; travels to the GPU box)."""
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref/libg4hepem_ref.so not built (needs /root/reference)")
    return ref.Reference(STATE_JSON)


@pytest.fixture(scope="session")
def oracle(reference):
    return reference


@pytest.fixture(scope="session")
def engine(flat_tables):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from g4hepem_b200 import engine as eng

    return eng.Engine(flat_tables, device=0)
