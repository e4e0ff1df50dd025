from pednstream_b200.config import load_config, validate_config  # noqa: F401
