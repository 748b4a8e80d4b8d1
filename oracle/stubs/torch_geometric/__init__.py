"""Stub of the PyG surface /root/reference/src/hashing.py imports (lines 13-15). TEST INFRASTRUCTURE ONLY."""
from . import nn, utils, loader  # noqa: F401
