"""History containers of the drop-in boundary (reference ``src/aspire/history.py:71-82``)."""

from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class History:
    pass


@dataclass
class FlowHistory(History):
    training_loss: list = field(default_factory=list)
    validation_loss: list = field(default_factory=list)


@dataclass
class SMCHistory(History):
    """Same fields as the reference's ``SMCHistory`` (plotting / HDF5 are out of scope)."""

    log_norm_ratio: list = field(default_factory=list)
    log_norm_ratio_var: list = field(default_factory=list)
    beta: list = field(default_factory=list)
    ess: list = field(default_factory=list)
    ess_target: list = field(default_factory=list)
    eff_target: list = field(default_factory=list)
    mcmc_autocorr: list = field(default_factory=list)
    mcmc_acceptance: list = field(default_factory=list)
    sample_history: list = field(default_factory=list)
    # extras of the B200 sampler (not in the reference)
    rho: list = field(default_factory=list)
    n_beta_evals: list = field(default_factory=list)
