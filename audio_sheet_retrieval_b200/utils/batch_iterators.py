"""batch_compute1 / batch_compute2 with the reference's semantics
(audio_sheet_retrieval/utils/batch_iterators.py:17-62, 65-111): fixed-size batches, the last one
zero-padded, `prepare` applied per batch, valid rows copied out.

When `compute` is a function compiled by `network.compile_function` and `prepare` is the model's
own `prepare`, the per-batch Python loop is replaced by one call into the library, which does the
same batching on the device side (chunks of the encoder's max_batch, prepare fused into layer 0).
Rows are independent in deterministic mode, so the result is identical to the padded loop.
"""
from __future__ import print_function

import sys

import numpy as np


def _fused(compute, prepare):
    return getattr(compute, "asr_fused", None) is not None and (
        prepare is None or getattr(prepare, "asr_prepare_mode", None) is not None)


def batch_compute1(X, compute, batch_size, verbose=False, prepare=None):
    """Batch compute data"""
    if _fused(compute, prepare):
        return compute.asr_fused(X, prepare)
    R = None
    n_samples = X.shape[0]
    in_shape = list(X.shape)[1:]
    n_batches = int(np.ceil(float(n_samples) / batch_size))
    for i_batch in range(n_batches):
        if verbose:
            print("Processing batch %d / %d" % (i_batch + 1, n_batches), end='\r')
            sys.stdout.flush()
        start_idx = i_batch * batch_size
        E = X[start_idx:start_idx + batch_size]
        n_missing = batch_size - E.shape[0]
        if n_missing > 0:
            E = np.vstack((E, np.zeros([n_missing] + in_shape, dtype=X.dtype)))
        if prepare is not None:
            E = prepare(E)
        r = compute(E)
        if R is None:
            R = np.zeros([n_samples] + list(r.shape[1:]), dtype=r.dtype)
        R[start_idx:start_idx + r.shape[0]] = r[0:batch_size - n_missing]
    return R


def batch_compute2(X1, X2, compute, batch_size, prepare1=None, prepare2=None):
    """Batch compute data.  As in the reference (:98-99), a non-None `prepare2` makes `prepare1`
    run on the second input; every caller passes prepare2=None."""
    if prepare2 is None and _fused(compute, prepare1):
        return compute.asr_fused2(X1, X2, prepare1)
    R = None
    n_samples = X1.shape[0]
    in_shape1 = list(X1.shape)[1:]
    in_shape2 = list(X2.shape)[1:]
    n_batches = int(np.ceil(float(n_samples) / batch_size))
    for i_batch in range(n_batches):
        start_idx = i_batch * batch_size
        E1, E2 = X1[start_idx:start_idx + batch_size], X2[start_idx:start_idx + batch_size]
        n_missing = batch_size - E1.shape[0]
        if n_missing > 0:
            E1 = np.vstack((E1, np.zeros([n_missing] + in_shape1, dtype=X1.dtype)))
            E2 = np.vstack((E2, np.zeros([n_missing] + in_shape2, dtype=X2.dtype)))
        if prepare1 is not None:
            E1 = prepare1(E1)
        if prepare2 is not None:
            E2 = prepare1(E2)
        r = compute(E1, E2)
        if R is None:
            R = np.zeros([n_samples] + list(r.shape[1:]), dtype=r.dtype)
        R[start_idx:start_idx + r.shape[0]] = r[0:batch_size - n_missing]
    return R
