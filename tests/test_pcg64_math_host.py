"""The PCG64 arithmetic asp_pcg64_uniform ships (aspire_b200/csrc/pcg64_math.cuh) compiled for the HOST and replayed
with the kernel's jump-ahead / grid-stride scheme: the stream must equal ``numpy.random.Generator.random`` - the
uniforms behind ``Generator.choice`` at src/aspire/samples.py:1278 - BIT FOR BIT, for any seed, offset and grid
size.  Needs nvcc (host compilation only, no GPU)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    out = tmp_path_factory.mktemp("native") / "libpcg64_host.so"
    cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           os.path.join(ROOT, "tests", "native", "pcg64_host.cu"), "-o", str(out)]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    lib = C.CDLL(str(out))
    lib.pcg64_host.restype = C.c_int
    lib.pcg64_host.argtypes = [C.c_uint64] * 4 + [C.c_int64] * 3 + [C.c_void_p]
    return lib


def draw(lib, rng, n, offset=0, grid_threads=1024):
    st = rng.bit_generator.state["state"]
    state, inc = int(st["state"]), int(st["inc"])
    out = np.empty(n)
    assert lib.pcg64_host((state >> 64) & M64, state & M64, (inc >> 64) & M64, inc & M64, offset, n, grid_threads,
                          out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("seed", [0, 7, 20251104, 2**63 + 5])
@pytest.mark.parametrize("grid_threads", [1, 96, 1024, 148 * 8 * 256])
def test_shipped_pcg64_equals_numpy_random(host_lib, seed, grid_threads):
    n = 20_000
    got = draw(host_lib, np.random.default_rng(seed), n, grid_threads=grid_threads)
    ref = np.random.default_rng(seed).random(n)
    assert np.array_equal(got.view(np.int64), ref.view(np.int64))


def test_offset_continues_the_stream_and_matches_choice(host_lib):
    rng = np.random.default_rng(123)
    ref = np.random.default_rng(123).random(5000)
    a = draw(host_lib, rng, 1234)
    b = draw(host_lib, rng, 5000 - 1234, offset=1234)
    assert np.array_equal(np.concatenate([a, b]), ref)
    # after the host generator has been used, its current state is the starting point
    rng = np.random.default_rng(9)
    rng.random(777)
    assert np.array_equal(draw(host_lib, rng, 100), rng.random(100))
    # the ancestors Generator.choice would pick: searchsorted(cdf, u, 'right') with these uniforms
    p = np.random.default_rng(1).random(1000)
    p /= p.sum()
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    u = draw(host_lib, np.random.default_rng(55), 1000)
    assert np.array_equal(cdf.searchsorted(u, side="right"), np.random.default_rng(55).choice(1000, 1000, p=p))
