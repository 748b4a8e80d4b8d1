"""Stub of `datasketch.hyperloglog_const`: `_thresholds`, `_bias`, `_raw_estimate`, indexable by p - 4.

The real module carries the Heule et al. (2013) tables; they cannot be reproduced offline, so the stub
serves the packaged Monte-Carlo tables -- the SAME arrays the engine loads when datasketch is absent.
"""
import os

import numpy as np

_path = os.environ.get('SS_B200_HLLPP_TABLES') or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), '..', '..', '..', 'subgraph_sketching_b200', 'data',
    'hllpp_tables.npz')
_blob = np.load(_path)
_thresholds = [int(t) for t in _blob['thresholds']]
_bias = []
_raw_estimate = []
for _p in range(4, 19):
    if f'bias_p{_p}' in _blob:
        _bias.append([float(v) for v in _blob[f'bias_p{_p}']])
        _raw_estimate.append([float(v) for v in _blob[f'raw_estimate_p{_p}']])
    else:  # not generated yet
        _bias.append([0.0] * 6)
        _raw_estimate.append([float(i) for i in range(6)])
