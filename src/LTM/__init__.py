"""Reference import path `src.LTM` -> pednstream_b200."""
from pednstream_b200.network import Network  # noqa: F401
from pednstream_b200.link import BaseLink, Link, Separator  # noqa: F401
from pednstream_b200.node import Node, OneToOneNode, RegularNode  # noqa: F401
