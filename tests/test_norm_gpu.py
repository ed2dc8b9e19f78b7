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
