"""Stub of the `datasketch` surface that /root/reference/src/hashing.py touches (lines 12, 69-80).

TEST INFRASTRUCTURE ONLY.  `datasketch` is not installed in this image; this stub lets the unmodified
reference module import.  Constants follow the HyperLogLog++ definition: alpha_m, max_rank = 64 - p,
2^p registers.  The empirical tables come from the packaged Monte-Carlo file (see hyperloglog_const).
"""
import hashlib
import struct

import numpy as np

from . import hyperloglog_const  # noqa: F401


def sha1_hash64(data):
    return struct.unpack('<Q', hashlib.sha1(data).digest()[:8])[0]


class HyperLogLogPlusPlus(object):
    def __init__(self, p=8, reg=None, hashfunc=sha1_hash64, hashobj=None):
        if not 4 <= p <= 18:
            raise ValueError('p must be in [4, 18]')
        self.p = p
        self.m = 1 << p
        self.reg = np.zeros((self.m,), dtype=np.int8)
        self.hashfunc = hashfunc
        self.max_rank = 64 - p
        if p == 4:
            self.alpha = 0.673
        elif p == 5:
            self.alpha = 0.697
        elif p == 6:
            self.alpha = 0.709
        else:
            self.alpha = 0.7213 / (1.0 + 1.079 / self.m)


class MinHash(object):  # imported by the reference tests only; not used by the oracle
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')
