"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_files(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def rel_err(got, ref, floor=1e-3):
    """SURVEY 8(c) metric: max |d| / max(|ref|, floor * scale), scale = max |ref| of the tensor."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    if ref.size == 0:
        return 0.0
    scale = np.abs(ref).max()
    if scale == 0:
        return float(np.abs(got).max())
    return float((np.abs(got - ref) / np.maximum(np.abs(ref), floor * scale)).max())


def scale_err(got, ref):
    """max |d| / max |ref|: error relative to the tensor's scale (used for
    gradient sums, whose element-wise conditioning in fp32 is ~1e-3; see
    DESIGN.md 'tolerances')."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    if ref.size == 0:
        return 0.0
    scale = np.abs(ref).max()
    return float(np.abs(got - ref).max() / (scale if scale > 0 else 1.0))


def quat_sign_align(q, ref):
    """Quaternions are compared up to sign (roma does not canonicalise it)."""
    s = np.sign((q * ref).sum(-1, keepdims=True))
    s[s == 0] = 1
    return q * s
