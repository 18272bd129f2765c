"""The numpy restatement of the preconditioning transforms (oracle/transforms.py, scope row 8f-4) against the
golden vectors produced by the reference's own ``CompositeTransform`` (oracle/make_golden.py) and - in the dev
container - against the reference itself, live.  Tolerance: 1e-13 relative on y / x, 1e-12 absolute on log|det J|
(the stages' Jacobians are added in a different order)."""
import glob
import logging
import os

import numpy as np
import pytest

from oracle import ref_shim
from oracle import transforms as ot

CASES = ["logit4", "probit3", "mixed32", "affine5"]


def load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"transforms_{name}.npz"))
    t = ot.make(g["kind"], g["lower"], g["upper"], str(g["bounded"]) or None, float(g["eps"]))
    return g, t


def test_golden_files_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "transforms_*.npz"))) == len(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(golden_dir, name):
    g, t = load(golden_dir, name)
    fitted, z = ot.fit(t, g["x"], affine=bool(g["affine"]))
    np.testing.assert_allclose(z, g["z_fit"], rtol=1e-13, atol=1e-13)
    if bool(g["affine"]):
        np.testing.assert_allclose(fitted["mean"], g["mean"], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(fitted["std"], g["std"], rtol=1e-14)
    y, lj = ot.forward(fitted, g["x"])
    np.testing.assert_allclose(y, g["y"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(lj, g["log_j"], rtol=0, atol=1e-12)
    x, lji = ot.inverse(fitted, g["z_in"])
    np.testing.assert_allclose(x, g["x_inv"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(lji, g["log_j_inv"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("name", CASES)
def test_round_trip_and_jacobians_cancel(golden_dir, name):
    g, t = load(golden_dir, name)
    fitted, _ = ot.fit(t, g["x"], affine=bool(g["affine"]))
    rng = np.random.default_rng(5)
    z = 0.3 * rng.normal(size=(200, len(t["kind"])))  # away from the eps clamp of the unit-interval value
    x, lji = ot.inverse(fitted, z)
    z2, ljf = ot.forward(fitted, x)
    per = t["kind"] == ot.PERIODIC
    np.testing.assert_allclose(z2[:, ~per], z[:, ~per], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(lji + ljf, 0.0, atol=1e-9)
    # periodic dimensions come back inside [lower, upper)
    assert np.all(x[:, per] >= t["lower"][per]) and np.all(x[:, per] < t["upper"][per])


def test_argument_errors():
    with pytest.raises(ValueError, match="prior bounds"):
        ot.from_composite_args(["a"], periodic_parameters=["a"], prior_bounds=None)
    with pytest.raises(ValueError, match="Unknown bounded transform"):
        ot.make([2], [0.0], [1.0], bounded="tanh")
    with pytest.raises(ValueError, match="too small"):
        ot.make([2], [1.0], [1.0], bounded="logit")
    t = ot.from_composite_args(["a", "b"], prior_bounds={"a": [0, 1], "b": [-np.inf, 3]}, bounded_transform="logit")
    assert t["kind"].tolist() == [ot.BOUNDED, ot.UNTOUCHED]  # half-open ranges are left alone (transforms.py:185-191)
    t = ot.from_composite_args(["a", "b"], prior_bounds={"a": [0, 1], "b": [0, 3]}, bounded_to_unbounded=False)
    assert t["kind"].tolist() == [0, 0] and t["bounded"] is None


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("bounded_transform", ["logit", "probit"])
@pytest.mark.parametrize("affine", [True, False])
def test_oracle_matches_reference_live(bounded_transform, affine):
    ref_shim.import_reference()
    logging.disable(logging.CRITICAL)
    try:
        import array_api_compat.numpy as xp
        from aspire.transforms import CompositeTransform

        params = ["a", "b", "c", "d", "e"]
        bounds = {"a": [0, 1], "b": [0, 2 * np.pi], "c": [-np.inf, np.inf], "d": [-2, 3], "e": [0, np.inf]}
        ref = CompositeTransform(parameters=params, periodic_parameters=["b"], prior_bounds=bounds,
                                 bounded_to_unbounded=True, bounded_transform=bounded_transform,
                                 affine_transform=affine, xp=xp, dtype="float64")
        t = ot.from_composite_args(params, ["b"], bounds, True, bounded_transform, affine)
        rng = np.random.default_rng(3)
        x = np.stack([rng.uniform(0, 1, 300), rng.uniform(-9, 9, 300), rng.normal(size=300), rng.uniform(-2, 3, 300),
                      rng.exponential(size=300)], axis=1)
        z_ref = np.asarray(ref.fit(x.copy()))
        fitted, z = ot.fit(t, x, affine=affine)
        np.testing.assert_allclose(z, z_ref, rtol=1e-13, atol=1e-13)
        y_ref, lj_ref = ref.forward(x.copy())
        y, lj = ot.forward(fitted, x)
        np.testing.assert_allclose(y, np.asarray(y_ref), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(lj, np.asarray(lj_ref), atol=1e-12)
        zz = rng.normal(size=(300, 5)) * 2
        x_ref, lji_ref = ref.inverse(zz.copy())
        xi, lji = ot.inverse(fitted, zz)
        np.testing.assert_allclose(xi, np.asarray(x_ref), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(lji, np.asarray(lji_ref), atol=1e-12)
    finally:
        logging.disable(logging.NOTSET)


@pytest.mark.parametrize("bounded", ["logit", "probit"])
def test_log_det_jacobian_is_the_log_determinant_of_the_map(bounded):
    """Independent of the reference: log|det J| reported by forward / inverse equals the log-determinant of a
    central-difference Jacobian of the map itself (the map is diagonal, so the determinant is the product of the
    per-dimension slopes)."""
    rng = np.random.default_rng(17)
    kind = np.array([2, 1, 0, 2, 0], dtype=np.int32)
    lower = np.array([0.0, -1.0, 0.0, -3.0, 0.0])
    upper = np.array([1.0, 2.0, 1.0, 5.0, 1.0])
    t = ot.make(kind, lower, upper, bounded, 1e-9)
    x = np.stack([rng.uniform(0.05, 0.95, 40), rng.uniform(-0.9, 1.9, 40), rng.normal(size=40),
                  rng.uniform(-2.5, 4.5, 40), rng.normal(size=40)], axis=1)
    t, _ = ot.fit(t, x, affine=True)
    h = 1e-6
    y, lj = ot.forward(t, x)
    slopes = np.empty_like(x)
    for j in range(x.shape[1]):
        e = np.zeros(x.shape[1])
        e[j] = h
        slopes[:, j] = (ot.forward(t, x + e)[0][:, j] - ot.forward(t, x - e)[0][:, j]) / (2 * h)
    np.testing.assert_allclose(lj, np.log(np.abs(slopes)).sum(1), rtol=0, atol=2e-6)
    z = rng.normal(size=(40, 5)) * 0.7
    xi, lji = ot.inverse(t, z)
    for j in range(x.shape[1]):
        e = np.zeros(x.shape[1])
        e[j] = h
        slopes[:, j] = (ot.inverse(t, z + e)[0][:, j] - ot.inverse(t, z - e)[0][:, j]) / (2 * h)
    np.testing.assert_allclose(lji, np.log(np.abs(slopes)).sum(1), rtol=0, atol=2e-6)
