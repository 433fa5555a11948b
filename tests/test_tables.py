"""Table hand-off: JSON schema conformity (the reference's own reader/writer), flattening parity."""
import ctypes as C
import json
import os

import numpy as np

from g4hepem_b200 import _capi, tables


def test_json_loads_in_reference_and_descriptors_agree(reference, flat_tables):
    rt = reference.flat_tables()
    for (n1, a1), (n2, a2) in zip(tables.descriptor_arrays(flat_tables.desc), tables.descriptor_arrays(rt)):
        assert n1 == n2
        assert a1.shape == a2.shape and np.array_equal(a1, a2), n1
    for f in tables.SCALAR_FIELDS:
        assert getattr(flat_tables.desc, f) == getattr(rt, f), f
    for part in ("electron", "positron"):
        for f in tables.ELECTRON_SCALAR_FIELDS:
            assert getattr(getattr(flat_tables.desc, part), f) == getattr(getattr(rt, part), f), (part, f)


def test_json_round_trip_through_reference_writer(reference, tmp_path):
    """G4HepEmStateToJson -> file -> our loader == our loader on the original (testing/DataImportExport pattern)."""
    out = tmp_path / "rt.json"
    assert reference.save_state(str(out)) == 0
    a = tables.load_state_json(str(out))
    from tests.conftest import STATE_JSON

    b = tables.load_state_json(STATE_JSON)
    for (n1, x), (n2, y) in zip(tables.descriptor_arrays(a.desc), tables.descriptor_arrays(b.desc)):
        assert n1 == n2 and np.array_equal(x, y), n1


def test_fixture_shapes(flat_tables):
    t = flat_tables.desc
    assert t.num_matcut == 7 and t.num_mat == 5 and t.num_regions == 3
    assert t.electron.num_loss == 85 and t.gm_data_per_mat == 2 * 32 + 3 * 32 + 9 * 256
    nel = np.ctypeslib.as_array(t.mat_num_elem, shape=(t.num_mat,))
    assert list(nel) == [1, 1, 1, 3, 2]


def test_library_exports_every_declared_symbol():
    """The C-ABI shared library loads without a GPU and exports what include/g4hepem_b200.h declares."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "g4hepem_b200.h")).read()
    declared = set(re.findall(r"\b(g4hb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    lib = _capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        return
    from g4hepem_b200 import engine
    from tests.conftest import STATE_JSON
    import pytest

    ft = tables.load_state_json(STATE_JSON)
    with pytest.raises(RuntimeError):
        engine.Engine(ft)
    lib = _capi.load_library()
    h = C.c_void_p()
    assert lib.g4hb200_create(C.byref(ft.desc), 0, C.byref(h)) == -2  # G4HB200_ENODEVICE
