"""Statistics of the DEVICE samplers on >= 10^7 draws: the precise Box-Muller stream (fp32 paths) and
the hardware-approximate one every bf16 path uses (lean sampling / finishing kernels, fused rank-1 epilogues):
moments (mean, variance, skewness, excess kurtosis), Kolmogorov-Smirnov against N(0,1) / U(0,1), sign balance, and
value agreement between the two normal streams."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 1 << 24  # 16.8 M draws


@pytest.mark.parametrize("kind", ["normal", "normal_fast"])
def test_device_normal_stream_statistics(cuda, kind):
    from scipy import stats

    from edward2_b200 import ops

    z = ops.philox(kind, N, 2026, 11).double().cpu().numpy()
    se = 1.0 / np.sqrt(N)
    assert abs(z.mean()) < 5 * se
    assert abs(z.var() - 1.0) < 5 * np.sqrt(2.0) * se
    assert abs(stats.skew(z)) < 5 * np.sqrt(6.0) * se
    assert abs(stats.kurtosis(z)) < 5 * np.sqrt(24.0) * se
    ks = stats.kstest(z, "norm")
    assert ks.pvalue > 1e-3, ks
    # tails: the 24-bit uniform bounds |z| by sqrt(-2 ln 2^-25) = 5.89; both tails are populated
    assert 4.5 < z.max() < 5.9 and -5.9 < z.min() < -4.5


def test_device_fast_stream_matches_precise(cuda):
    from edward2_b200 import ops

    a = ops.philox("normal", N, 7, 3)
    b = ops.philox("normal_fast", N, 7, 3)
    err = (a - b).abs()
    assert float(err.max()) < 2e-3 and float(err.mean()) < 5e-6  # sin/cos.approx near multiples of pi dominate the max


def test_device_uniform_and_sign_statistics(cuda):
    from scipy import stats

    from edward2_b200 import ops

    u = ops.philox("uniform", N, 2026, 12).double().cpu().numpy()
    assert 0.0 < u.min() and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    s = ops.philox("sign", N, 2026, 13).cpu().numpy()
    assert set(np.unique(s)) == {-1.0, 1.0}
    assert stats.binomtest(int((s > 0).sum()), N, 0.5).pvalue > 1e-3
