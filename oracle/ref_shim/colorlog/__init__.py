"""Stub of the `colorlog` package (absent in this image) so that the read-only
reference at /root/reference can be imported by oracle/gen_golden.py.
TEST INFRASTRUCTURE ONLY -- never imported by the product package."""
import logging
import re


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, log_colors=None, **kw):
        fmt = re.sub(r"%\(log_color\)s", "", fmt or "")
        super().__init__(fmt, datefmt)
