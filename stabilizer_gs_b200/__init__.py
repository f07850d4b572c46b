"""stabilizer_gs_b200 — B200-native exact stabilizer-ground-state search (drop-in for the sparse /
klocal_sparse / klocal_periodic path of SUSYUSTC/stabilizer_gs).

Everything that computes runs in hand-written sm_100a CUDA kernels inside libsgs.so (C ABI in
include/sgs.h); this package is the ctypes binding plus thin mirrors of the reference's Python
classes.  There is no CPU fallback: importing works without a GPU, searching does not.
"""
from . import _capi
from ._capi import (HamHandle, SgsError, SparsePlan, klocal_periodic, klocal_search, measure_int_peaks,
                    nccl_unique_id, sparse_search)

__all__ = ['HamHandle', 'SgsError', 'SparsePlan', 'klocal_periodic', 'klocal_search', 'measure_int_peaks',
           'nccl_unique_id', 'sparse_search', '_capi']
