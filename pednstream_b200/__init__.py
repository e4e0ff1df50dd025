"""pednstream_b200: B200-native Link-Transmission-Model timestep behind PedNStream's Python API.

Host side (this package) mirrors the reference's objects -- `NetworkEnvGenerator`, `Network`,
`Link`/`Separator`/`Node` views -- and drives hand-written sm_100a CUDA kernels
(`csrc/`) through a thin C-ABI (`include/pns_b200.h`).
"""
from .config import load_config, validate_config  # noqa: F401
from .env_loader import NetworkEnvGenerator  # noqa: F401
from .network import Network  # noqa: F401
from .link import BaseLink, Link, Separator  # noqa: F401
from .node import Node, OneToOneNode, RegularNode  # noqa: F401

__version__ = "0.1.0"
