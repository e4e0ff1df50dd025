from pednstream_b200.od_manager import DemandConfig, DemandGenerator, ODManager  # noqa: F401
