"""Pick the strongest available checker (TEST INFRASTRUCTURE; see oracle/README in DESIGN.md).

  1. oracle/_ref/libg4hepem_ref.so -- the unmodified reference compiled from /root/reference (kind "reference")
  2. oracle/_build/libg4hepem_oracle.so -- the plain-C restatement (kind "port")
"""
import os

from . import ref as _ref


def best_available(json_path):
    if _ref.available():
        r = _ref.Reference(json_path)
        r.kind = "reference"
        return r
    from . import port as _port

    if _port.available():
        p = _port.Port(json_path)
        p.kind = "port"
        return p
    raise RuntimeError("no oracle library built: run __graft_entry__.build()")
