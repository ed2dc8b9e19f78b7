"""The x87 norm of the device code, compiled for the host (tests/host_shim/norm_host.cpp) and
compared with the oracle's long-double dnrm2: the same vectors as tests/test_norm_gpu.py, so a
change to distance3d_b200/csrc/d3d_math.cuh is checked here before it is checked on a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import cpu_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "distance3d_b200", "csrc", "d3d_math.cuh")
BEGIN, END = "// ---- exact emulation", "D3D_DEV real norm_dd("


@pytest.fixture(scope="module", params=[0.0, 3.0, -3.0])
def host_norm(tmp_path_factory, request):
    """params: bias of the stand-in for the device's reciprocal estimate (units of 2^-36); the
    sqrt correction of the fast path tolerates a quotient good to 2^-38, so every bias must give
    the same bits."""
    text = open(HEADER).read()
    b, e = text.index(BEGIN), text.index(END)
    d = tmp_path_factory.mktemp("norm_host")
    section = d / "norm_section.inc"
    section.write_text(text[b:e])
    so = d / "norm_host.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-mfma", "-ffp-contract=off", "-w",
                           "-DRCP_BIAS=%r" % request.param,
                           '-DNORM_SECTION="%s"' % section, os.path.join(HERE, "host_shim", "norm_host.cpp"),
                           "-o", str(so)])
    lib = ctypes.CDLL(str(so))

    def run(v, mode):
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(v))
        lib.host_norm(v.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(v)), ctypes.c_int(mode),
                      out.ctypes.data_as(ctypes.c_void_p))
        return out
    return run


def vectors():
    rs = np.random.RandomState(11)
    v = rs.randn(300000, 3)
    scaled = rs.randn(100000, 3) * 10.0 ** rs.uniform(-30, 30, size=(100000, 1))
    spread = rs.randn(100000, 3) * 10.0 ** rs.uniform(-12, 12, size=(100000, 3))
    unit = rs.randn(300000, 3)
    unit /= np.sqrt((unit * unit).sum(axis=1))[:, None]
    unit *= 2.0 ** rs.randint(-3, 4, size=(len(unit), 1))
    ints = rs.randint(-20, 21, size=(100000, 3)).astype(np.float64) * 2.0 ** rs.randint(-4, 5, size=(100000, 1))
    axis = np.zeros((30000, 3))
    axis[np.arange(30000), rs.randint(0, 3, 30000)] = rs.randn(30000)
    pow2 = 2.0 ** rs.randint(-40, 40, size=(30000, 3)) * rs.choice([-1.0, 0.0, 1.0], size=(30000, 3))
    tiny = rs.randn(20000, 3) * 10.0 ** rs.uniform(-300, -140, size=(20000, 1))   # norms in the subnormal
    # range (< 2.2e-308) are rounded twice by the rescaling path: out of scope, as on the GPU test
    huge = rs.randn(20000, 3) * 10.0 ** rs.uniform(140, 300, size=(20000, 1))
    return np.concatenate([v, scaled, spread, unit, ints, axis, pow2, tiny, huge, np.zeros((3, 3))])


def test_production_path_matches_the_long_double_norm(host_norm):
    v = vectors()
    np.testing.assert_array_equal(host_norm(v, 0), O.norm(v))


def test_integer_emulation_seeded_by_the_estimate_matches(host_norm):
    v = vectors()
    np.testing.assert_array_equal(host_norm(v, 2), O.norm(v))


def test_integer_emulation_with_a_coarse_seed_matches(host_norm):
    v = vectors()[::4]
    np.testing.assert_array_equal(host_norm(v, 1), O.norm(v))
