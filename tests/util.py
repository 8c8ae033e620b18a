"""Shared helpers of the test-suite (fixture loading, tolerances)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances stated by BASELINE.json north_star.
FP32_RTOL = 1e-5   # fp32 outputs and gradients: 1e-5 relative
BF16_RTOL = 1e-2   # bf16 outputs and gradients: 1e-2 relative


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def bf16_bits_to_f32(a: np.ndarray) -> np.ndarray:
    """int16/uint16 array of bf16 bit patterns -> float32."""
    return (a.view(np.uint16).astype(np.uint32) << 16).view(np.float32)


def rel_err(got, want) -> float:
    """max |got-want| over the finite entries of `want`, relative to max |want|.

    "Relative" is taken against the tensor's scale (max magnitude): elementwise relative
    error is meaningless for sums that cancel to ~0.  Non-finite patterns must agree.
    """
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), "non-finite pattern differs"
    if not fin.any():
        return 0.0
    scale = max(np.abs(want[fin]).max(), 1e-30)
    return float(np.abs(got[fin] - want[fin]).max() / scale)


def assert_close(got, want, rtol, what=""):
    e = rel_err(got, want)
    assert e <= rtol, f"{what}: relative error {e:.3e} > {rtol:.1e}"


def elementwise_err(got, want, rtol: float, atol_rms: float = None) -> float:
    """max over elements of |got-want| / (rtol*|want| + atol), atol = atol_rms (default rtol) x RMS(want).

    The elementwise companion of `rel_err`: <= 1 means every element satisfies
    |got-want| <= rtol*|want| + atol with the absolute term tied to the tensor's RMS (not its max), so
    small elements may not hide behind one large one."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), "non-finite pattern differs"
    if not fin.any():
        return 0.0
    w, g = want[fin], got[fin]
    rms = max(float(np.sqrt(np.mean(w * w))), 1e-30)
    atol = (rtol if atol_rms is None else atol_rms) * rms
    return float((np.abs(g - w) / (rtol * np.abs(w) + atol)).max())


def assert_close_elementwise(got, want, rtol, what="", atol_rms=None):
    e = elementwise_err(got, want, rtol, atol_rms)
    assert e <= 1.0, f"{what}: elementwise |a-b| exceeds rtol*|b| + atol(RMS) by a factor {e:.3g} (rtol {rtol:.1e})"
