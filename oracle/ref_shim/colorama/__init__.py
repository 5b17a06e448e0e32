"""Stub of the `colorama` package (absent in this image); see colorlog stub."""


def init(*args, **kwargs):
    pass
