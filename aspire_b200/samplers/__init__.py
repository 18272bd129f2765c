from .base import Sampler  # noqa: F401
from .importance import ImportanceSampler  # noqa: F401
from .smc import BetaScheduleError, MiniPCNSMC, SMCSampler  # noqa: F401
