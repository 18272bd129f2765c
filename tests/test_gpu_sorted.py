"""Sorted-uniform multinomial resampling on the GPU (asp_sorted_spacings / asp_sorted_tile_scan / asp_search_sorted)
against the numpy spec oracle/sorted_resample.py: bit-exact integers and ancestors, through the C-ABI."""
import numpy as np
import pytest
import torch

from oracle import sorted_resample as sr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda")


def _cdf(kind, n, rng):
    if kind == "equal":
        w = np.full(n, 1.0 / n)
    elif kind == "one_hot":
        w = np.zeros(n)
        w[n // 3] = 1.0
    elif kind == "two_ends":
        w = np.zeros(n)
        w[0] = 0.25
        w[-1] = 0.75
    elif kind == "zeros30":
        w = np.exp(rng.normal(0, 2.0, n))
        w[rng.random(n) < 0.3] = 0.0
        w[-1] = max(w[-1], 1e-3)  # never all zero
    else:
        w = np.exp(rng.normal(0, float(kind), n))
    c = np.cumsum(w)
    return c / c[-1]


@pytest.mark.parametrize("m", [1, 5, 2047, 2048, 2049, 100_003])
def test_spacings_and_scan_match_the_spec(cuda, m):
    from aspire_b200 import ops

    sp = ops.SortedSpacings(20251104, 7, m, cuda).scan()
    q = sr.spacings(20251104, 7, m)
    got = sp.q.cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(got[: m + 1], q)
    assert not got[m + 1:].any()
    s, total = sr.partial_sums(q)
    assert int(sp.total.item()) == total
    tile_first = np.concatenate([[0], np.cumsum(q.astype(np.uint64))])[:: ops.SPACING_TILE][: sp.tiles_total]
    np.testing.assert_array_equal(sp.tile_prefix.cpu().numpy().view(np.uint64), tile_first)


@pytest.mark.parametrize("kind", ["0.5", "2", "5", "equal", "one_hot", "two_ends", "zeros30"])
@pytest.mark.parametrize("n,m", [(1000, 1000), (6000, 2500), (3000, 40_000), (70_001, 70_001), (1, 300)])
def test_ancestors_match_the_spec(cuda, kind, n, m):
    from aspire_b200 import ops

    rng = np.random.default_rng(n + m)
    cdf = _cdf(kind, n, rng)
    got = ops.search_multinomial_sorted(torch.from_numpy(cdf).to(cuda), 99, 3, m).cpu().numpy()
    ref = sr.resample(cdf, 99, 3, m)
    np.testing.assert_array_equal(got, ref)
    assert np.all(np.diff(got) >= 0)


def test_slot_ranges_windows_and_lower_bound(cuda):
    from aspire_b200 import ops

    rng = np.random.default_rng(8)
    n, m = 50_000, 60_000
    cdf = _cdf("2", n, rng)
    ref = sr.resample(cdf, 5, 11, m)
    full = ops.SortedSpacings(5, 11, m, cuda).scan()
    d_cdf = torch.from_numpy(cdf).to(cuda)
    for j_lo, j_hi in [(0, 1), (17, 4100), (2048, 4096), (33_333, 60_000), (59_999, 60_000)]:
        np.testing.assert_array_equal(full.search(d_cdf, j_lo, j_hi).cpu().numpy(), ref[j_lo:j_hi])
    # a window of the spacings (what a rank regenerates for the slots it owns) + the global tile table
    j0, tiles = 8 * ops.SPACING_TILE, 9
    win = ops.SortedSpacings(5, 11, m, cuda, j0=j0, n_tiles=tiles)
    np.testing.assert_array_equal(win.q.cpu().numpy(), full.q[j0: j0 + tiles * ops.SPACING_TILE].cpu().numpy())
    win.scan(full.tile_sums)
    lo, hi = j0 + 5, j0 + tiles * ops.SPACING_TILE - 7
    np.testing.assert_array_equal(win.search(d_cdf, lo, hi).cpu().numpy(), ref[lo:hi])
    # a CDF slice (global values): ancestors are local indices, clamped outside the slice
    k0, k1 = 12_000, 31_000
    s, total = sr.partial_sums(sr.spacings(5, 11, m))
    t = sr.thresholds(cdf, total)
    j_a, j_b = np.searchsorted(s, t[k0 - 1], side="left"), np.searchsorted(s, t[k1 - 1], side="left")
    got = full.search(d_cdf[k0:k1].contiguous(), int(j_a), int(j_b)).cpu().numpy()
    np.testing.assert_array_equal(got + k0, ref[j_a:j_b])
    thr = torch.tensor([int(t[k0 - 1]), int(t[k1 - 1]), 0, 1, total, total + 5], dtype=torch.int64, device=cuda)
    lb = full.lower_bound(thr).cpu().numpy()
    np.testing.assert_array_equal(lb, np.minimum(np.searchsorted(s, thr.cpu().numpy().astype(np.uint64), side="left"), m))


def test_at_scale_against_torch_searchsorted(cuda):
    """N = M = 3e7: ancestors equal searchsorted-right of the integer thresholds at the integer partial sums."""
    from aspire_b200 import ops

    n = m = 30_000_000
    lw = ops.philox_normal(3, 1, n, 1, cuda).reshape(-1) * 2.0
    p = torch.exp(lw - lw.max())
    p /= p.sum()
    cdf = ops.cdf_exact(p)
    sp = ops.SortedSpacings(17, 2, m, cuda).scan()
    idx = sp.search(cdf)
    s = torch.cumsum(sp.q[:m].to(torch.int64) & 0xFFFFFFFF, 0)
    total = float(sp.total.item())
    t = torch.round(cdf * total).to(torch.int64)  # round-half-even on exact products, as rint
    ref = torch.searchsorted(t, s, right=True).clamp_max(n - 1)
    assert torch.equal(idx, ref)
    assert bool((idx[1:] >= idx[:-1]).all())
    counts = torch.bincount(idx, minlength=n).to(torch.float64)
    # offspring counts follow N w: the weighted chi-square over 1e3 coarse cells
    cells = 1000
    edges = torch.linspace(0, n, cells + 1, device=cuda).long()
    got = torch.stack([counts[edges[i]:edges[i + 1]].sum() for i in range(cells)])
    exp = torch.stack([p[edges[i]:edges[i + 1]].sum() for i in range(cells)]) * m
    chi2 = float(((got - exp) ** 2 / exp).sum())
    assert abs(chi2 - cells) < 6 * np.sqrt(2 * cells)
