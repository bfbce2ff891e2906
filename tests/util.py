"""Helpers shared by the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, f"case_{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def load_npz(fname):
    with np.load(os.path.join(GOLDEN, fname)) as z:
        return {k: z[k] for k in z.files}


TREE_KEYS = ("ids", "x", "y", "parent", "cost", "nchild")


def assert_tree_equal(got, want, what=""):
    """Bit-exact comparison of two canonical trees (ids ascending, parent ids, fp64 coordinates / costs)."""
    for k in TREE_KEYS:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, f"{what}: {k} has {a.shape} entries, expected {b.shape}"
        if a.dtype.kind == "f":
            # bit-exact: compare the raw 64-bit patterns so that -0.0 / NaN cannot hide
            bad = np.flatnonzero(a.view(np.uint64) != b.astype(np.float64).view(np.uint64))
        else:
            bad = np.flatnonzero(a != b)
        assert bad.size == 0, f"{what}: {k} differs at rows {bad[:8]} (got {a[bad[:4]]}, want {b[bad[:4]]})"


def assert_trace_equal(tr, g, n, mode, what=""):
    """Per-iteration records vs the golden records."""
    def eq(field, gold):
        a, b = np.asarray(tr[field][:n]), np.asarray(g[gold][:n])
        bad = np.flatnonzero(a != b)
        assert bad.size == 0, f"{what}: trace.{field} differs at iterations {bad[:8]} (got {a[bad[:4]]}, want {b[bad[:4]]})"
    eq("id_new", "rec_id_new")
    eq("id_near", "rec_id_near")
    if mode != "rrt":
        eq("n_near", "rec_n_near")
        eq("n_rewired", "rec_n_rewired")
        eq("removed_id", "rec_removed")
