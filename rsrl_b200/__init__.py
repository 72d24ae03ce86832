"""rsrl_b200 — B200-native vectorised RL step engine behind rsrl's trait surface.

The product is the CUDA library rsrl_b200/csrc/librsrl_b200.so (C ABI: include/rsrl_b200.h).
This package is the Python host side over that ABI; it has no CPU fallback.
"""
from . import abi  # noqa: F401
from .abi import Config, default_config, RsrlError  # noqa: F401
