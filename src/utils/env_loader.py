from pednstream_b200.env_loader import NetworkEnvGenerator  # noqa: F401
