"""subgraph_sketching_b200 -- B200-native subgraph-sketch engine behind the ElphHashes API of
melifluos/subgraph-sketching (src/hashing.py).  Importing it loads libss_b200.so and fails loudly when the
library is missing; there is no CPU fallback."""
from . import _lib  # noqa: F401  (loads the CUDA library or raises)
from .hashing import (LABEL_LOOKUP, ElphHashes, HllPropagation, HopSketch, MinhashPropagation, SketchTables,
                      build_csr, hll_alpha, hllpp_tables, log2_window_table)

__all__ = ['LABEL_LOOKUP', 'ElphHashes', 'HllPropagation', 'MinhashPropagation', 'SketchTables', 'HopSketch',
           'build_csr', 'hll_alpha', 'hllpp_tables', 'log2_window_table']
