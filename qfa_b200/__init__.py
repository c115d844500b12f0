"""qfa_b200 -- B200-native (sm_100a) implementation of the QFA hot path.

Drop-in surface (same names as the reference package `QFA`):
    from qfa_b200.model import QFA
    from qfa_b200.optimizer import Adam, step_scheduler
    from qfa_b200.utils import tau
"""
from ._lib import QfaError, build, lib  # noqa: F401
from .model import QFA  # noqa: F401
from .optimizer import Adam, step_scheduler  # noqa: F401
from .utils import tau, default_tau  # noqa: F401
from .dataloader import DeviceDataloader  # noqa: F401

__version__ = "0.1.0"
