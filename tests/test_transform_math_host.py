"""The arithmetic the transform kernel ships (aspire_b200/csrc/transform_math.cuh) compiled for the HOST and run
against the reference-generated golden vectors and the numpy oracle: the GPU test then only has to establish that
the kernel's indexing hands every element to this code.  Needs nvcc (host compilation only, no GPU)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import transforms as ot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CASES = ["logit4", "probit3", "mixed32", "affine5"]


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    out = tmp_path_factory.mktemp("native") / "libtransform_host.so"
    cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
           os.path.join(ROOT, "tests", "native", "transform_host.cu"), "-o", str(out)]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    lib = C.CDLL(str(out))
    P = C.c_void_p
    lib.transform_host.restype = C.c_int
    lib.transform_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, P, P, P, P, P, P, C.c_int64, P, P]
    return lib


def run(lib, t, x, inverse):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, d = x.shape
    out, lj = np.empty_like(x), np.empty(n)
    kind = np.ascontiguousarray(t["kind"], dtype=np.int32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    mean = None if t["mean"] is None else np.ascontiguousarray(t["mean"])
    std = None if t["std"] is None else np.ascontiguousarray(t["std"])
    lower, upper = np.ascontiguousarray(t["lower"]), np.ascontiguousarray(t["upper"])
    bounded = {None: 0, "logit": 1, "probit": 2}[t["bounded"]]
    rc = lib.transform_host(int(inverse), d, bounded, int(mean is not None), t["eps"], ptr(kind), ptr(lower), ptr(upper),
                            ptr(mean), ptr(std), ptr(x), n, ptr(out), ptr(lj))
    assert rc == 0
    return out, lj


@pytest.mark.parametrize("name", CASES)
def test_shipped_arithmetic_matches_reference_golden(host_lib, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"transforms_{name}.npz"))
    aff = bool(g["affine"])
    t = ot.make(g["kind"], g["lower"], g["upper"], str(g["bounded"]) or None, float(g["eps"]),
                g["mean"] if aff else None, g["std"] if aff else None)
    y, lj = run(host_lib, t, g["x"], inverse=False)
    np.testing.assert_allclose(y, g["y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lj, g["log_j"], rtol=0, atol=1e-10)
    x, lji = run(host_lib, t, g["z_in"], inverse=True)
    np.testing.assert_allclose(x, g["x_inv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lji, g["log_j_inv"], rtol=0, atol=1e-10)


def test_shipped_arithmetic_periodic_is_numpy_mod_bit_for_bit(host_lib):
    rng = np.random.default_rng(0)
    lower, upper = np.array([0.0, -1.5, 3.0]), np.array([2 * np.pi, 2.5, 3.5])
    t = ot.make([1, 1, 1], lower, upper)
    x = rng.normal(size=(5000, 3)) * 20
    x[:5] = [[0.0, -1.5, 3.0], [2 * np.pi, 2.5, 3.5], [-1e-20, -1.5 - 1e-17, 3.0 - 1e-16], [1e300, -1e300, 7.0],
             [-0.0, 2.5 - 1e-16, 3.25]]
    y, lj = run(host_lib, t, x, inverse=False)
    y_ref, _ = ot.forward(t, x)
    np.testing.assert_array_equal(y, y_ref)
    assert not lj.any()
    np.testing.assert_array_equal(run(host_lib, t, x, inverse=True)[0], ot.inverse(t, x)[0])


@pytest.mark.parametrize("bounded", ["logit", "probit"])
def test_shipped_arithmetic_matches_oracle_random(host_lib, bounded):
    rng = np.random.default_rng(9)
    d, n = 40, 700
    kind = rng.integers(0, 3, d).astype(np.int32)
    lower = rng.uniform(-5, 0, d)
    upper = lower + rng.uniform(0.5, 7, d)
    t = ot.make(kind, lower, upper, bounded, 1e-6)
    x = rng.normal(size=(n, d)) * 3
    for j in np.nonzero(kind == ot.BOUNDED)[0]:
        x[:, j] = rng.uniform(lower[j] - 0.1, upper[j] + 0.1, n)  # a few outside: clamped
    t, _ = ot.fit(t, x, affine=True)
    y, lj = run(host_lib, t, x, inverse=False)
    y_ref, lj_ref = ot.forward(t, x)
    np.testing.assert_allclose(y, y_ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lj, lj_ref, rtol=1e-13, atol=1e-10)
    z = rng.normal(size=(n, d)) * 2
    xi, lji = run(host_lib, t, z, inverse=True)
    x_ref, lji_ref = ot.inverse(t, z)
    np.testing.assert_allclose(xi, x_ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(lji, lji_ref, rtol=1e-13, atol=1e-10)
