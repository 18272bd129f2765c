"""The integer-pair arithmetic the exact-scan kernels ship (aspire_b200/csrc/scan_math.cuh) compiled for the HOST:
a cumsum rebuilt from it - pairs per element, composed by a balanced tree, applied to the running mantissa, one
genuine add per binade crossing - must equal ``numpy.cumsum`` (strictly sequential fp64 adds, what
``numpy.random.Generator.choice`` builds at src/aspire/samples.py:1278) BIT FOR BIT on adversarial inputs (the domain of the scan: finite, non-negative weights - NaN / inf are caught
upstream by the reweighting statistics and flagged by the kernels' ERR_BAD_WEIGHT).
Needs nvcc (host compilation only, no GPU)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    out = tmp_path_factory.mktemp("native") / "libscan_host.so"
    cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
           os.path.join(ROOT, "tests", "native", "scan_host.cu"), "-o", str(out)]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    lib = C.CDLL(str(out))
    lib.scan_host.restype = C.c_int
    lib.scan_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    return lib


def scan(lib, w, tile=2048):
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty_like(w)
    phases = C.c_int64(0)
    assert lib.scan_host(w.ctypes.data, len(w), tile, out.ctypes.data, C.byref(phases)) == 0
    return out, phases.value


def cases():
    rng = np.random.default_rng(2024)
    n = 50_000
    lw = rng.normal(size=n) * 5
    yield "softmax weights sigma=5", np.exp(lw - lw.max()) / np.exp(lw - lw.max()).sum()
    yield "uniform weights", np.full(n, 1.0 / n)
    yield "one-hot", np.eye(1, n, 1234).ravel()
    z = rng.random(n)
    z[rng.random(n) < 0.3] = 0.0
    yield "30 % zeros", z / z.sum()
    yield "ties: halves of an ulp", np.concatenate([[1.0], np.full(n - 1, 2.0 ** -53)])
    yield "ties: odd mantissa", np.concatenate([[1.0 + 2.0 ** -52], np.full(n - 1, 2.0 ** -53)])
    yield "powers of two across 80 binades", 2.0 ** rng.integers(-60, 20, n).astype(np.float64)
    yield "denormals", rng.random(n) * 1e-310
    yield "huge dynamic range", np.exp(rng.normal(size=n) * 40)
    yield "increasing", np.sort(rng.random(n))
    yield "decreasing", np.sort(rng.random(n))[::-1].copy()
    yield "many crossings", np.concatenate([[2.0 ** -1000], 2.0 ** np.arange(-1000, -1000 + 1500, dtype=np.float64)])
    yield "single element", np.array([0.3])
    yield "ragged tail", rng.random(2048 * 3 + 17)


@pytest.mark.parametrize("name,w", list(cases()), ids=[c[0] for c in cases()])
@pytest.mark.parametrize("tile", [2048, 7])
def test_shipped_pair_arithmetic_rebuilds_numpy_cumsum(host_lib, name, w, tile):
    out, phases = scan(host_lib, w, tile)
    ref = np.cumsum(w)
    assert out.view(np.int64).tolist() == ref.view(np.int64).tolist() if len(w) < 10 else np.array_equal(
        out.view(np.int64), ref.view(np.int64)), name
    assert phases >= -(-len(w) // tile)
