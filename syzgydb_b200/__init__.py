"""syzgydb_b200 -- B200-native (sm_100a) search hot path of SyzgyDB.

The product is the CUDA library libsyzgy_b200.so behind the C ABI of include/syzgy_b200.h;
this package holds its sources (csrc/), the ctypes binding tests and benchmarks use
(_capi), and the host-side mirror of the reference's Collection.Search surface
(collection).  There is no CPU implementation: importing works without a GPU, every call
that computes needs a B200.
"""
from . import _capi
from ._capi import COSINE, EUCLIDEAN, Index, SpanFile, SzgError

__all__ = ["_capi", "Index", "SpanFile", "SzgError", "EUCLIDEAN", "COSINE"]
