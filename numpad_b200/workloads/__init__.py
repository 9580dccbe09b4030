"""Residual functions of the benchmark configurations (BASELINE.json `configs`).

Each workload takes the array namespace ``xp`` explicitly (``numpad_b200``, the
CPU oracle, or the unmodified reference) so that the very same operation
sequence -- hence the same tape -- is recorded on all three.
"""
from .cavity import CavityProblem  # noqa: F401
from .kec import KecProblem  # noqa: F401
