"""The checker the smoke test and the bench use (TEST INFRASTRUCTURE).

There is one oracle: oracle/_ref/libg4hepem_ref.so, the unmodified reference compiled from /root/reference by
oracle/Makefile (kind "reference").  It is built where /root/reference exists and travels to the GPU box prebuilt.
"""
from . import ref as _ref


def best_available(json_path):
    if not _ref.available():
        raise RuntimeError(f"no oracle library built ({_ref.REF_LIB} missing): run __graft_entry__.build() where "
                           "/root/reference exists")
    r = _ref.Reference(json_path)
    r.kind = "reference"
    return r
