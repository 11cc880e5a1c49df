"""Shared helpers for the parity tests."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: reference-matching output within 1e-3 relative.
REL_TOL = 1e-3


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def max_over_max(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def assert_close(got, ref, tol=REL_TOL, what=""):
    got_t = torch.as_tensor(got)
    assert torch.isfinite(got_t.float()).all(), f"{what}: non-finite values"
    r, m = rel_l2(got, ref), max_over_max(got, ref)
    assert r <= tol and m <= tol, f"{what}: rel_l2={r:.3e} max/max={m:.3e} (tol {tol:g})"
    return r, m


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "meta" in d:
        d["meta"] = ast.literal_eval(str(d["meta"]))
    return d
