from pednstream_b200.network import Network  # noqa: F401
