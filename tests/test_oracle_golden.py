"""The numpy oracle against the golden vectors produced by the unmodified
reference (oracle/make_golden.py).  Runs everywhere (no GPU, no /root/reference)."""
import glob
import os

import numpy as np
import pytest

from oracle import smc_numpy as sn

RTOL = 1e-12


def _cases(prefix, golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    return sorted(glob.glob(os.path.join(golden_dir, f"{prefix}_*.npz")))


@pytest.mark.parametrize("path", _cases("reweight"), ids=os.path.basename)
def test_reweight_golden(path):
    g = np.load(path)
    ll, lp, lq, b0 = g["log_l"], g["log_p"], g["log_q"], float(g["beta0"])
    beta, _ = sn.determine_beta(ll, lp, lq, b0, target_efficiency=float(g["target_eff"]),
                                beta_tolerance=float(g["tol"]), min_beta_step=float(g["min_step"]),
                                max_beta_step=float(g["max_step"]))
    assert beta == float(g["beta"])  # identical decisions -> identical fp64 beta
    np.testing.assert_array_equal(sn.unnormalized_log_weights(ll, lp, lq, b0, beta), g["uw"])
    assert float(sn.effective_sample_size(sn.log_weights(ll, lp, lq, b0, beta))) == pytest.approx(float(g["ess"]), rel=RTOL)
    assert float(sn.effective_sample_size(sn.log_weights(ll, lp, lq, b0, 1.0))) == pytest.approx(float(g["ess_target"]), rel=RTOL)
    assert sn.log_evidence_ratio(ll, lp, lq, b0, beta) == pytest.approx(float(g["log_evidence_ratio"]), rel=RTOL)
    assert sn.log_evidence_ratio_variance(ll, lp, lq, b0, beta) == pytest.approx(float(g["log_evidence_ratio_var"]), rel=RTOL)
    w = sn.resample_weights(ll, lp, lq, b0, beta)
    np.testing.assert_array_equal(w, g["w"])
    # indices: bit-exact against what the reference's rng.choice returned
    np.testing.assert_array_equal(sn.choice_indices(g["w"], g["u"]), g["idx"])


@pytest.mark.parametrize("path", _cases("importance"), ids=os.path.basename)
def test_importance_golden(path):
    g = np.load(path)
    r = sn.compute_weights(g["log_l"], g["log_p"], g["log_q"])
    np.testing.assert_array_equal(r["log_w"], g["log_w"])
    np.testing.assert_array_equal(r["weights"], g["weights"])
    for k in ("log_evidence", "evidence", "evidence_error", "log_evidence_error", "effective_sample_size"):
        assert r[k] == pytest.approx(float(g[k]), rel=RTOL), k
