"""aspire_b200 - B200-native SMC particle engine behind aspire's sampler API.

One hot path of mj-will/aspire, rebuilt as hand-written sm_100a CUDA behind a
C-ABI (``include/aspire_b200.h``): importance reweighting, log-sum-exp / ESS,
adaptive-beta bisection, exact multinomial / systematic resampling, the pCN
mutation sweep and built-in benchmark log-densities.  See DESIGN.md.
"""

from ._lib import AspireB200Error  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    """``from aspire_b200 import Aspire, Samples`` like ``from aspire import Aspire, Samples`` (lazy: torch
    is only imported when the classes are asked for)."""
    if name == "Aspire":
        from .aspire import Aspire

        return Aspire
    if name in ("Samples", "SMCSamples"):
        from . import samples

        return getattr(samples, name)
    raise AttributeError(f"module 'aspire_b200' has no attribute {name!r}")
