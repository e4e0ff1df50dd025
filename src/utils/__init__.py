"""Reference import path `src.utils` -> pednstream_b200."""
