"""GPU: bit-exact emulation of the reference's np.linalg.norm (BLAS dnrm2 on the
x87 FPU) on the device."""
import numpy as np
import pytest

from distance3d_b200 import _lib
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _vectors():
    rs = np.random.RandomState(3)
    v = rs.randn(400000, 3) * 10.0 ** rs.uniform(-6, 6, size=(400000, 1))
    special = np.array([[1.0, 0, 0], [0, 0, 0], [0, 0, 1.0], [3.0, 4.0, 0.0], [1e-160, 1e-160, 0],
                        [1e150, 1e150, 1e150], [1.0, 1e-30, 0.0], [0.5, 0.5, 0.5],
                        [5e-324, 0, 0], [1e-310, 2e-310, 0]])
    unit = rs.randn(100000, 3)
    unit /= np.linalg.norm(unit, axis=1)[:, None]  # results close to 1.0 (a power of two)
    two_d = rs.randn(100000, 3)
    two_d[:, 2] = 0.0
    # components of very different magnitude (exactness hazards of the double-double replay)
    skew = rs.randn(100000, 3) * 10.0 ** rs.uniform(-12, 2, size=(100000, 3))
    # products that land exactly on / next to powers of two and 64-bit ties
    pow2 = np.zeros((4096, 3))
    pow2[:, 0] = 2.0 ** rs.randint(-30, 30, size=4096)
    pow2[:, 1] = pow2[:, 0] * 2.0 ** -rs.randint(0, 40, size=4096) * rs.choice([0.0, 1.0, 3.0], size=4096)
    ints = rs.randint(1, 2 ** 20, size=(100000, 3)).astype(float)   # exact integer squares
    return np.concatenate([v, special, unit, two_d, skew, pow2, ints])


def test_production_norm_is_bit_exact():
    v = _vectors()
    np.testing.assert_array_equal(_lib.debug_norm(v, 0), O.norm(v))


def test_exact_emulation_path_alone_is_bit_exact():
    v = _vectors()[:200000]
    np.testing.assert_array_equal(_lib.debug_norm(v, 1), O.norm(v))


def test_exact_path_seeded_by_the_double_double_estimate():
    """The hazard path as production runs it (integer root seeded by the double-double estimate),
    forced for every vector: unit vectors whose norm rounds to 1.0 from below take it every
    other time (EPA hands unit face normals to the sphere / ellipsoid supports)."""
    v = _vectors()
    np.testing.assert_array_equal(_lib.debug_norm(v, 2), O.norm(v))
    rs = np.random.RandomState(5)
    n = rs.randn(200000, 3)
    n /= np.sqrt((n * n).sum(axis=1))[:, None]
    n *= 2.0 ** rs.randint(-3, 4, size=(200000, 1))      # norms next to other powers of two as well
    np.testing.assert_array_equal(_lib.debug_norm(n, 2), O.norm(n))
    np.testing.assert_array_equal(_lib.debug_norm(n, 0), O.norm(n))


def test_vector_division_equals_ieee_division():
    """v / s on the device (three quotients sharing one reciprocal refinement, d3d_math.cuh)
    is bit for bit the correctly rounded quotient, including signed zeros, subnormals,
    overflow / underflow ranges, infinities and divisors with all-ones mantissas."""
    rs = np.random.RandomState(11)
    n = 4_000_000
    v = rs.randn(n, 3) * 10.0 ** rs.uniform(-8, 8, size=(n, 1))
    s = rs.randn(n) * 10.0 ** rs.uniform(-8, 8, size=n)
    # random bit patterns over the whole exponent range
    bits = rs.randint(0, 2 ** 63, size=(200000, 4), dtype=np.int64) * rs.choice([-1, 1], size=(200000, 4))
    wide = bits.view(np.float64)
    wide = wide[np.isfinite(wide).all(axis=1)]
    # hard cases for Newton-Raphson division: divisors with all-ones / almost-one mantissas,
    # powers of two, quotients next to rounding ties
    ones = ((1 << 52) - 1 - rs.randint(0, 4, size=200000)) | (np.int64(1023) << 52)
    hard_s = np.concatenate([ones.view(np.float64), 2.0 ** rs.randint(-40, 40, size=1000).astype(float),
                             1.0 + rs.randint(0, 4, size=1000) * 2.0 ** -52])
    m = rs.randint(0, 2 ** 52, size=(len(hard_s), 3), dtype=np.int64)
    hard_v = (m | (np.int64(1023) << 52)).view(np.float64)
    special_v = np.array([[0.0, -0.0, 1.0], [5e-324, -5e-324, 1e-310], [1e308, -1e308, 1.0],
                          [1e-300, 1e300, 3.0], [np.inf, -np.inf, 1.0], [1.0, 2.0, 3.0],
                          [1.0, 2.0, 3.0], [1.0, -1.0, 0.0], [1e-200, 1e-250, 1e-320]])
    special_s = np.array([3.0, 7.0, 1e-10, 1e-20, 2.0, 0.0, np.inf, 5e-324, 1e200])
    V = np.concatenate([v, wide[:, :3], hard_v, special_v])
    S = np.concatenate([s, wide[:, 3], hard_s, special_s])
    with np.errstate(all="ignore"):
        expect = V / S[:, None]
    got = _lib.debug_vdiv(V, S)
    # bitwise comparison (distinguishes -0.0 from 0.0; NaNs compare by payload class only)
    nan = np.isnan(expect)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got.view(np.int64)[~nan], expect.view(np.int64)[~nan])
