"""Lasagne parameter pickles: the reference's on-disk ABI.

Format (reference: audio_sheet_retrieval/refine_cca.py:110-111, utils/train_dcca_pool.py:399-401,
retrieval_wrapper.py:27-29): Python-2 pickle, protocol 2, a flat list of float32 ndarrays in
`lasagne.layers.get_all_param_values` order.  For the two shipped model families that is 97 arrays:
view 1 then view 2, 9 groups of (W, beta, gamma, mean, inv_std), then the CCALayer's
U, V, mean1, mean2, S12, S11, S22 (layers/cca.py:69-77).
"""
import pickle

import numpy as np

N_CONV = 9
GROUP = 5
N_PARAMS = 2 * N_CONV * GROUP + 7
IDX_U, IDX_V, IDX_MEAN1, IDX_MEAN2, IDX_S12, IDX_S11, IDX_S22 = 90, 91, 92, 93, 94, 95, 96


def load_params(param_file):
    """Read a parameter pickle written by Python 2 (or by `save_params`)."""
    with open(param_file, "rb") as fp:
        params = pickle.load(fp, encoding="latin1")
    if len(params) and isinstance(params[0], list):
        # "old redundant dump": one full list per output layer (run_eval.py:76-80); the list
        # for the last layer (l_v2latent) covers the whole graph
        params = params[-1]
    return [np.ascontiguousarray(p, dtype=np.float32) for p in params]


class _Py2NumpyPickler(pickle._Pickler):
    """Writes `numpy.core.*` globals (NumPy >= 2 would write `numpy._core.*`, which the
    reference's NumPy 1.13 cannot import)."""

    def save_global(self, obj, name=None):
        module = getattr(obj, "__module__", None) or ""
        if module.startswith("numpy._core"):
            name = name or getattr(obj, "__qualname__", obj.__name__)
            self.write(pickle.GLOBAL + module.replace("numpy._core", "numpy.core").encode("ascii") + b"\n" +
                       name.encode("ascii") + b"\n")
            self.memoize(obj)
            return
        super().save_global(obj, name)


def save_params(param_file, params):
    """Write the flat list the way the reference does (`pickle.dump(..., protocol=-1)` under
    Python 2 = protocol 2)."""
    params = [np.ascontiguousarray(p, dtype=np.float32) for p in params]
    with open(param_file, "wb") as fp:
        _Py2NumpyPickler(fp, protocol=2).dump(params)


def split_params(params):
    """-> (views, cca): views[v][l] = dict(W, beta, gamma, mean, inv_std); cca = dict(U, V, ...)."""
    if len(params) != N_PARAMS:
        raise ValueError("expected %d parameter arrays, got %d" % (N_PARAMS, len(params)))
    views = []
    for v in range(2):
        layers = []
        for l in range(N_CONV):
            o = (v * N_CONV + l) * GROUP
            W, beta, gamma, mean, inv_std = params[o:o + GROUP]
            layers.append(dict(W=W, beta=beta, gamma=gamma, mean=mean, inv_std=inv_std))
        views.append(layers)
    cca = dict(U=params[IDX_U], V=params[IDX_V], mean1=params[IDX_MEAN1], mean2=params[IDX_MEAN2],
               S12=params[IDX_S12], S11=params[IDX_S11], S22=params[IDX_S22])
    return views, cca


def expected_shapes(filters, dim_latent=32):
    """Array shapes of a 97-array list for the given per-layer filter counts."""
    shapes = []
    for _ in range(2):
        cin = 1
        for l in range(N_CONV):
            cout = filters[l] if l < 8 else dim_latent
            k = 3 if l < 8 else 1
            shapes += [(cout, cin, k, k)] + [(cout,)] * 4
            cin = cout
    shapes += [(dim_latent, dim_latent)] * 2 + [(dim_latent,)] * 2 + [(dim_latent, dim_latent)] * 3
    return shapes
