from pednstream_b200.link import BaseLink, Link, Separator  # noqa: F401
