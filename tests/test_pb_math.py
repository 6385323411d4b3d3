"""pb_expf (include/pb_math.h) against a correctly rounded exp: the shared exponential must
stay within 2 ulp so that float results sit far inside the 1e-4 relative parity tolerance."""
import numpy as np


def _ulp_err(got32, x32):
    ref = np.exp(x32.astype(np.float64))
    ulp = np.ldexp(1.0, np.floor(np.log2(ref)).astype(np.int64) - 23)
    return np.abs(got32.astype(np.float64) - ref) / ulp


def test_expf_accuracy_negative_range(orc):
    rng = np.random.default_rng(1)
    x = np.concatenate([-rng.uniform(0, 86, 400_000), -np.logspace(-8, 1.9, 50_000), [0.0, -0.0, -86.0]]).astype(np.float32)
    y = orc.expf(x)
    assert _ulp_err(y, x).max() <= 2.0


def test_expf_accuracy_positive_range(orc):
    x = np.linspace(0, 88, 200_001, dtype=np.float32)
    assert _ulp_err(orc.expf(x), x).max() <= 2.0


def test_expf_edges(orc):
    y = orc.expf(np.array([0.0, -1e30, -87.0, -200.0, 1e9, np.nan], np.float32))
    assert y[0] == 1.0 and y[1] == 0.0 and y[2] == 0.0 and y[3] == 0.0
    assert np.isfinite(y[4]) and np.isnan(y[5])


def test_expf_monotone_on_grid(orc):
    x = np.linspace(-30, 0, 300_001, dtype=np.float32)
    y = orc.expf(x)
    # never decreases by more than 1 ulp on an increasing grid
    assert (np.diff(y.astype(np.float64)) >= -np.spacing(y[1:]).astype(np.float64)).all()
