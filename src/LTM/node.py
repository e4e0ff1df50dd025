from pednstream_b200.node import Node, OneToOneNode, RegularNode  # noqa: F401
