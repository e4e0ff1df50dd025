from pednstream_b200.path_finder import PathFinder, enumerate_shortest_simple_paths  # noqa: F401
