"""The container contract (scope row a15) on the CPU suite: the reference's own tests of ``BaseSamples`` /
``Samples`` / ``SMCSamples`` (tests/test_samples.py - the ones that concern this path; HDF5, dataframe and
chain containers are out of scope) run against ``aspire_b200.samples`` with the ctypes layer replaced by the test
doubles.  Carrier difference, by design: fields are torch fp64 tensors, not numpy arrays."""
import math
import os
import pickle
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))

import ops_double  # noqa: E402


@pytest.fixture(autouse=True)
def cpu(monkeypatch):
    ops_double.install(monkeypatch)
    return "cpu"


def make_simple_samples(n=5, d=2, a=0.0):
    """tests/test_samples.py:13-30: log_likelihood = a, log_prior = log_q = 0 -> evidence e^a, ESS n."""
    return np.arange(n * d, dtype=float).reshape(n, d), np.full(n, a), np.zeros(n), np.zeros(n)


def test_basesamples_defaults_and_dims():  # :40-57
    from aspire_b200.samples import BaseSamples

    bs = BaseSamples(x=np.zeros((3, 4)))
    assert bs.xp is torch and bs.dtype == torch.float64 and bs.x.dtype == torch.float64
    assert bs.parameters == ["x_0", "x_1", "x_2", "x_3"]
    assert bs.dims == 4 and len(bs) == 3


def test_basesamples_to_dict_flat_and_nested():  # :60-70
    from aspire_b200.samples import BaseSamples

    ll = np.array([0.1, 0.2, 0.3])
    bs = BaseSamples(x=np.arange(6.0).reshape(3, 2), log_likelihood=ll, parameters=["a", "b"])
    flat = bs.to_dict(flat=True)
    assert "a" in flat and "b" in flat and np.allclose(flat["log_likelihood"], ll)
    assert np.allclose(flat["b"], [1.0, 3.0, 5.0])
    nested = bs.to_dict(flat=False)
    assert isinstance(nested["samples"], dict) and set(nested["samples"]) == {"a", "b"}
    assert nested["parameters"] == ["a", "b"]


def test_basesamples_getitem_returns_same_type_and_fields():  # :90-109
    from aspire_b200.samples import BaseSamples

    ll, lp, lq = np.linspace(0, 1, 6), np.linspace(1, 2, 6), np.linspace(-1, 0, 6)
    bs = BaseSamples(x=np.arange(12.0).reshape(6, 2), log_likelihood=ll, log_prior=lp, log_q=lq, parameters=["a", "b"])
    part = bs[1:4]
    assert isinstance(part, BaseSamples) and part.x.shape == (3, 2) and part.parameters == ["a", "b"]
    assert np.allclose(part.log_likelihood.numpy(), ll[1:4]) and np.allclose(part.log_q.numpy(), lq[1:4])
    mask = torch.tensor([True, False, True, False, False, True])
    assert bs[mask].x.shape == (3, 2) and np.allclose(bs[mask].log_prior.numpy(), lp[[0, 2, 5]])
    idx = torch.tensor([5, 5, 0])
    assert np.allclose(bs[idx].x.numpy(), bs.x.numpy()[[5, 5, 0]])


def test_basesamples_pickle_restores_namespace_and_fields():  # :121-130
    from aspire_b200.samples import BaseSamples

    x = np.arange(6.0).reshape(3, 2)
    bs2 = pickle.loads(pickle.dumps(BaseSamples(x=x)))
    assert bs2.xp is torch and bs2.dims == 2 and np.allclose(bs2.x.numpy(), x)
    from aspire_b200.samples import SMCSamples

    smc = SMCSamples(x=x, log_likelihood=np.zeros(3), log_prior=np.zeros(3), log_q=np.zeros(3), beta=0.25)
    smc.log_weights(0.5)  # fills the device scratch, which must not be pickled
    back = pickle.loads(pickle.dumps(smc))
    assert back.beta == 0.25 and "_slope_cache" not in back.__dict__ and back.device == smc.device
    assert "Log evidence" not in str(back) and "No. samples: 3" in str(back)


def test_samples_compute_weights_constant_case():  # :162-185
    from aspire_b200.samples import Samples

    n, a = 10, 0.7
    x, ll, lp, lq = make_simple_samples(n=n, d=3, a=a)
    s = Samples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, parameters=[f"p{i}" for i in range(3)])
    assert np.allclose(s.log_w.numpy(), a)
    assert s.log_evidence == pytest.approx(a, rel=1e-12) and s.evidence == pytest.approx(math.exp(a), rel=1e-12)
    assert s.evidence_error == pytest.approx(0.0, abs=1e-12)
    assert float(s.effective_sample_size) == pytest.approx(n, rel=1e-12) and s.efficiency == pytest.approx(1.0)
    assert np.allclose(s.scaled_weights.numpy(), 1.0)


def test_samples_to_dict_includes_weight_fields():  # :188-206
    from aspire_b200.samples import Samples

    x, ll, lp, lq = make_simple_samples(n=4, d=2, a=0.3)
    d = Samples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, parameters=["a", "b"]).to_dict(flat=True)
    for k in ["log_w", "weights", "evidence", "log_evidence", "evidence_error", "log_evidence_error",
              "effective_sample_size", "a", "b"]:
        assert k in d, k


def test_samples_getitem_preserves_evidence_fields():  # :209-225
    from aspire_b200.samples import Samples

    x, ll, lp, lq = make_simple_samples(n=6, d=2, a=0.2)
    s = Samples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, parameters=["a", "b"])
    s2 = s[1:4]
    assert isinstance(s2, Samples) and s2.x.shape[0] == 3
    assert math.isclose(float(s2.log_evidence), float(s.log_evidence), rel_tol=1e-12)
    assert math.isclose(float(s2.log_evidence_error), float(s.log_evidence_error), rel_tol=1e-12, abs_tol=1e-300)


def test_samples_concatenate_success_and_mismatch_errors():  # :228-286
    from aspire_b200.samples import Samples

    x1, ll1, lp1, lq1 = make_simple_samples(n=3, d=2, a=0.0)
    x2, ll2, lp2, lq2 = make_simple_samples(n=4, d=2, a=0.1)
    s1 = Samples(x=x1, log_likelihood=ll1, log_prior=lp1, log_q=lq1, parameters=["a", "b"])
    s2 = Samples(x=x2, log_likelihood=ll2, log_prior=lp2, log_q=lq2, parameters=["a", "b"])
    sc = Samples.concatenate([s1, s2])
    assert isinstance(sc, Samples) and sc.x.shape == (7, 2) and sc.parameters == ["a", "b"]
    assert np.allclose(sc.log_likelihood.numpy(), np.concatenate([ll1, ll2]))
    bad = Samples(x=x2, log_likelihood=ll2, log_prior=lp2, log_q=lq2, parameters=["a", "c"])
    with pytest.raises(ValueError):
        Samples.concatenate([s1, bad])


def test_samples_rejection_sample_accepts_proportionally():  # :289-306
    from aspire_b200.samples import Samples

    n = 100
    ll = np.zeros(n)
    ll[0] = 10.0
    s = Samples(x=np.arange(n, dtype=float).reshape(n, 1), log_likelihood=ll, log_prior=np.zeros(n), log_q=np.zeros(n),
                parameters=["p0"])
    rs = s.rejection_sample(rng=np.random.default_rng(1234))
    assert isinstance(rs, Samples) and len(rs) <= len(s)
    if len(rs) > 0:
        assert 0 in rs.x.numpy().squeeze().astype(int).reshape(-1)


def test_smc_to_standard_samples_and_getitem_preserve_fields():  # :734-759
    from aspire_b200.samples import Samples, SMCSamples

    x, ll, lp, lq = make_simple_samples(n=5, d=2, a=0.4)
    smc = SMCSamples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=0.3, parameters=["a", "b"], log_evidence=1.23,
                     log_evidence_error=0.1)
    std = smc.to_standard_samples()
    assert isinstance(std, Samples) and not isinstance(std, SMCSamples)
    assert std.log_q is None and std.log_w is None and std.log_evidence == 1.23 and std.log_evidence_error == 0.1
    part = smc[1:3]
    assert isinstance(part, SMCSamples) and part.beta == 0.3 and part.log_evidence == 1.23


def test_smcsamples_to_dict_includes_smc_fields():  # :775-791
    from aspire_b200.samples import SMCSamples

    x, ll, lp, lq = make_simple_samples(n=4, d=2, a=0.5)
    d = SMCSamples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=0.3, parameters=["a", "b"], log_evidence=0.8,
                   log_evidence_error=0.1).to_dict(flat=True)
    assert d["beta"] == 0.3 and d["log_evidence"] == 0.8 and d["log_evidence_error"] == 0.1


def test_str_contains_counts_and_metrics():  # :762-772
    from aspire_b200.samples import Samples

    x, ll, lp, lq = make_simple_samples(n=4, d=2, a=0.7)
    text = str(Samples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, parameters=["a", "b"]))
    for needle in ("No. samples: 4", "No. parameters: 2", "Log evidence:", "Effective sample size:", "Efficiency:"):
        assert needle in text


@pytest.mark.parametrize("flat", [True, False])
def test_basesamples_from_dict(flat):  # :144-152
    from aspire_b200.samples import BaseSamples, Samples

    ll = np.array([0.1, 0.2, 0.3])
    bs = BaseSamples(x=np.arange(6.0).reshape(3, 2), log_likelihood=ll, parameters=["a", "b"])
    bs2 = BaseSamples.from_dict(bs.to_dict(flat=flat))
    assert isinstance(bs2, BaseSamples) and bs2.parameters == bs.parameters
    assert np.allclose(bs2.log_likelihood.numpy(), ll) and np.allclose(bs2.x.numpy(), bs.x.numpy())
    x, l2, lp, lq = make_simple_samples(n=4, d=2, a=0.3)
    s = Samples(x=x, log_likelihood=l2, log_prior=lp, log_q=lq, parameters=["a", "b"])
    s2 = Samples.from_dict(s.to_dict(flat=flat))  # weights and evidence are recomputed, not read back
    assert s2.log_evidence == pytest.approx(s.log_evidence) and np.allclose(s2.weights.numpy(), s.weights.numpy())
    with pytest.raises(ValueError, match="Parameters must be provided"):
        BaseSamples.from_dict({"a": np.zeros(3)})


def test_to_dataframe():  # :155-160, :309-319, :794-806
    from aspire_b200.samples import BaseSamples, Samples, SMCSamples

    bs = BaseSamples(x=np.arange(6.0).reshape(3, 2), log_likelihood=np.array([0.1, 0.2, 0.3]), parameters=["a", "b"])
    df = bs.to_dataframe(include=["log_likelihood"])
    assert "log_likelihood" in df.columns and "log_prior" not in df.columns
    assert "a" in df.columns and "b" in df.columns and len(df) == len(bs)
    assert bs.to_dataframe()["log_prior"].isna().all()
    x, ll, lp, lq = make_simple_samples(n=4, d=2, a=0.5)
    df = Samples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, parameters=["a", "b"]).to_dataframe()
    assert {"a", "b", "log_w", "log_likelihood"} <= set(df.columns) and len(df) == 4
    df = SMCSamples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=0.2, parameters=["a", "b"]).to_dataframe()
    assert {"a", "b", "log_likelihood"} <= set(df.columns) and len(df) == 4
