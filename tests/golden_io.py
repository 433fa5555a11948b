"""Helpers for the committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the
unmodified reference)."""
import os

import numpy as np

from g4hepem_b200 import batches

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def batch_from(z, prefix, cls):
    n = z[prefix + "winner"].shape[0]
    b = cls(n)
    for g in b.groups() + ("meta", "winner"):
        if prefix + g in z.files:  # `prestep` (state of the track-level calls only) is younger than the golden files
            getattr(b, g)[...] = z[prefix + g]
    return b


class GoldenSecondaries:
    """Quacks like a SecondaryHostQueue for tests/compare.compare_secondaries."""

    def __init__(self, z, prefix):
        self.rec = {k: z[prefix + k] for k in ("parent_index", "slot", "parent_id", "kind", "dir", "ekin")}

    def sorted_records(self):
        return self.rec


def electron_batch(z, prefix):
    return batch_from(z, prefix, batches.ElectronHostBatch)


def gamma_batch(z, prefix):
    return batch_from(z, prefix, batches.GammaHostBatch)
