"""Host-side logic of aspire_b200.transforms and of the preconditioning plumbing in Aspire.init_sampler
(no kernel calls: the transforms are built on device="cpu", where any compute would raise).  The configuration
tests follow the reference's tests/test_transforms.py:284-331."""
import math

import pytest
import torch

from aspire_b200 import transforms as T
from aspire_b200._lib import AspireB200Error


def test_composite_transform_requires_prior_bounds_for_periodic():
    with pytest.raises(ValueError):
        T.CompositeTransform(parameters=[0], periodic_parameters=[0], prior_bounds=None, device="cpu")


def test_composite_transform_new_instance_copies_configuration():
    parameters = ["x0", "x1"]
    t = T.CompositeTransform(parameters=parameters, periodic_parameters=[], prior_bounds={p: (-1.0, 1.0) for p in parameters},
                             affine_transform=False, device="cpu")
    new = t.new_instance()
    assert new.xp is t.xp and new.dtype == t.dtype
    assert new.parameters == t.parameters and new.periodic_parameters == t.periodic_parameters
    assert new._kind_host == t._kind_host == [T.BOUNDED, T.BOUNDED]


def test_flow_transform_config_drops_periodic_parameters():
    parameters = ["a", "b"]
    t = T.FlowTransform(parameters=parameters, prior_bounds={p: (-2.0, 2.0) for p in parameters}, device="cpu")
    cfg = t.config_dict()
    assert "periodic_parameters" not in cfg and cfg["parameters"] == parameters


def test_logit_transform_interval_check_raises():
    with pytest.raises(ValueError):
        T.LogitTransform(lower=0.0, upper=0.0, device="cpu")


def test_kind_selection_follows_the_reference():
    # transforms.py:185-191: finite two-sided bounds and not periodic -> bounded; half-open ranges are left alone
    bounds = {"a": [0, 1], "b": [0, 2 * math.pi], "c": [-math.inf, math.inf], "d": [0, math.inf]}
    t = T.CompositeTransform(list(bounds), periodic_parameters=["b"], prior_bounds=bounds, bounded_transform="logit",
                             device="cpu")
    assert t._kind_host == [T.BOUNDED, T.PERIODIC, T.UNTOUCHED, T.UNTOUCHED]
    assert t.bounded_parameters == ["a"] and not t.is_identity
    t = T.CompositeTransform(list(bounds), prior_bounds=bounds, bounded_to_unbounded=False, affine_transform=False,
                             device="cpu")
    assert t._kind_host == [0, 0, 0, 0] and t.is_identity and t.bounded_parameters is None
    with pytest.raises(ValueError, match="Unknown bounded transform"):
        T.CompositeTransform(["a"], prior_bounds={"a": [0, 1]}, bounded_transform="tanh", device="cpu")


def test_unfitted_affine_and_cpu_tensors_fail_loudly():
    t = T.CompositeTransform(["a"], prior_bounds={"a": [0, 1]}, device="cpu")
    with pytest.raises(RuntimeError, match="fitted"):
        t.forward(torch.zeros(3, 1, dtype=torch.float64))
    t = T.CompositeTransform(["a"], prior_bounds={"a": [0, 1]}, affine_transform=False, device="cpu")
    with pytest.raises(AspireB200Error, match="CUDA tensor"):  # there is no CPU path
        t.forward(torch.full((3, 1), 0.5, dtype=torch.float64))
    ident = T.IdentityTransform(device="cpu")
    x = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    y, lj = ident.forward(x)
    assert torch.equal(y, x) and torch.equal(lj, torch.zeros(3, dtype=torch.float64)) and ident.is_identity


class _Flow:
    device = "cpu"

    def log_prob(self, x):
        return torch.zeros(len(x), dtype=torch.float64)


def _aspire(**kw):
    from aspire_b200 import Aspire

    return Aspire(log_likelihood=lambda s: 0, log_prior=lambda s: 0, dims=2, parameters=["a", "b"], device="cpu",
                  flow=_Flow(), **kw)


def test_init_sampler_preconditioning_selection():
    # aspire.py:330-368
    a = _aspire()
    assert a.init_sampler("smc").preconditioning_transform is None            # no bounds: nothing to apply
    assert a.init_sampler("importance").preconditioning_transform is None
    a = _aspire(prior_bounds={"a": [0, 1], "b": [0, 6.0]}, periodic_parameters=["b"])
    t = a.init_sampler("smc").preconditioning_transform
    assert t._kind_host == [T.UNTOUCHED, T.PERIODIC] and not t.affine_transform  # defaults of :338-341
    t = a.init_sampler("smc", preconditioning="standard",
                       preconditioning_kwargs={"bounded_to_unbounded": True, "affine_transform": True}
                       ).preconditioning_transform
    assert t._kind_host == [T.BOUNDED, T.PERIODIC] and t.affine_transform and t.bounded_transform == "logit"
    assert a.init_sampler("smc", preconditioning="none").preconditioning_transform is None
    t = a.init_sampler("smc", preconditioning="flow", preconditioning_kwargs={"fit_kwargs": {"n_epochs": 3}}
                       ).preconditioning_transform                            # aspire.py:351-366
    assert isinstance(t, T.FlowPreconditioningTransform) and not t.affine_transform and t.fit_kwargs == {"n_epochs": 3}
    assert t._data_transform._kind_host == [T.BOUNDED, T.PERIODIC] and t.flow is None
    with pytest.raises(RuntimeError, match="fitted"):
        t.forward(torch.zeros(3, 2, dtype=torch.float64))
    with pytest.raises(ValueError, match="Unknown preconditioning"):
        a.init_sampler("smc", preconditioning="whitening")


def test_bounds_need_parameter_names():
    from aspire_b200 import Aspire

    with pytest.raises(ValueError, match="parameters"):
        Aspire(log_likelihood=lambda s: 0, log_prior=lambda s: 0, dims=1, prior_bounds={"a": [0, 1]}, device="cpu")
