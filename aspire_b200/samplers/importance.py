"""``sampler="importance"`` (reference ``src/aspire/samplers/importance.py:6-22``)."""

from __future__ import annotations

from ..samples import Samples
from .base import Sampler, track_calls


class ImportanceSampler(Sampler):
    @track_calls
    def sample(self, n_samples: int) -> Samples:
        x, log_q = self.prior_flow.sample_and_log_prob(n_samples)
        samples = Samples(x, log_q=log_q, parameters=self.parameters, device=self.device)
        samples.log_prior = self.log_prior(samples)          # prior first, then likelihood (importance.py:17-20)
        samples.log_likelihood = self.log_likelihood(samples)
        samples.compute_weights()                            # asp_importance_weights
        self._check_flow()                                   # a kernel time-out must not reach the weights unnoticed
        return samples
