"""Import-path shim: `src.*` of the reference layout resolves to pednstream_b200."""
