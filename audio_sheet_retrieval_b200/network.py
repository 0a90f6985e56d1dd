"""Graph-free stand-in for the handful of Lasagne/Theano entry points the reference scripts use.

The reference builds a Lasagne graph (`model.build_model`) and then calls
`lasagne.layers.set_all_param_values / get_all_param_values / get_output` and
`theano.function` on the returned layers (audio_sheet_retrieval/run_eval.py:59-95,
refine_cca.py:41-111, retrieval_wrapper.py:23-45).  Here `build_model` returns four light
`Layer` handles onto one `RetrievalNet`; "compiling a function" means creating an encoder handle
in libasr_b200.so (BN folded, bf16 weights packed, activation arena allocated).
"""
import ctypes

import numpy as np

from . import _lib
from . import params as P

MAX_BATCH_DEFAULT = 256


class Layer(object):
    """What the scripts touch on a Lasagne layer: `.input_var`, `.output_shape`, `.net`."""

    def __init__(self, net, kind, view, output_shape):
        self.net, self.kind, self.view = net, kind, view
        self.output_shape = tuple(output_shape)
        self.input_var = (net, view)          # opaque token, as a Theano variable is to callers
        self.name = "%s_view%d" % (kind, view)

    def __repr__(self):
        return "<Layer %s %s>" % (self.name, self.output_shape)


class CCALayer(object):
    """The CCALayer's refittable state: `.mean1/.mean2/.U/.V.set_value(...)` and `.input_layers`
    (audio_sheet_retrieval/refine_cca.py:78-84,104-107)."""

    class _Var(object):
        def __init__(self, net, index):
            self.net, self.index = net, index

        def set_value(self, value):
            value = np.ascontiguousarray(value, np.float32)
            if value.shape != self.net.params[self.index].shape:
                raise ValueError("shape mismatch for CCA parameter")
            self.net.params[self.index] = value
            self.net._cca_changed()

        def get_value(self):
            return self.net.params[self.index]

    def __init__(self, net):
        self.net = net
        self.U = CCALayer._Var(net, P.IDX_U)
        self.V = CCALayer._Var(net, P.IDX_V)
        self.mean1 = CCALayer._Var(net, P.IDX_MEAN1)
        self.mean2 = CCALayer._Var(net, P.IDX_MEAN2)
        self.input_layers = [Layer(net, "cca_in", 1, (None, 32)), Layer(net, "cca_in", 2, (None, 32))]


class Encoder(object):
    """One branch compiled into libasr_b200.so (asr_encoder_t)."""

    def __init__(self, net, view, prepare_mode, max_batch):
        self.net, self.view, self.prepare_mode, self.max_batch = net, view, prepare_mode, max_batch
        views, cca = P.split_params(net.params)
        layers = views[view - 1]
        d = _lib.EncoderDesc()
        in_shape = net.raw_shape(view) if prepare_mode == _lib.PREP_SCALE_HALF else net.input_shape(view)
        d.in_h, d.in_w = int(in_shape[1]), int(in_shape[2])
        d.prepare = prepare_mode
        d.flip_filters = int(net.flip_filters)
        self._keep = []
        fp = ctypes.POINTER(ctypes.c_float)
        for l in range(9):
            d.channels[l] = int(layers[l]["W"].shape[0])
            for key in ("W", "beta", "gamma", "mean", "inv_std"):
                a = np.ascontiguousarray(layers[l][key], np.float32)
                self._keep.append(a)
                getattr(d, key)[l] = a.ctypes.data_as(fp)
        mean = np.ascontiguousarray(cca["mean1" if view == 1 else "mean2"], np.float32)
        proj = np.ascontiguousarray(cca["U" if view == 1 else "V"], np.float32)
        self._keep += [mean, proj]
        d.cca_mean = mean.ctypes.data_as(fp)
        d.cca_proj = proj.ctypes.data_as(fp)
        self.in_h, self.in_w = d.in_h, d.in_w
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.asr_encoder_create(ctypes.byref(h), ctypes.byref(d), int(max_batch)))
        self.handle = h
        self.flops_per_sample = float(_lib.lib.asr_encoder_flops_per_sample(h))

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib.asr_encoder_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cca(self, mean, proj):
        fp = ctypes.POINTER(ctypes.c_float)
        mean = np.ascontiguousarray(mean, np.float32)
        proj = np.ascontiguousarray(proj, np.float32)
        _lib.check(_lib.lib.asr_encoder_set_cca(self.handle, mean.ctypes.data_as(fp), proj.ctypes.data_as(fp)))

    def _check_input(self, shape):
        if len(shape) != 4 or shape[1] != 1 or shape[2] != self.in_h or shape[3] != self.in_w:
            raise ValueError("expected input (n,1,%d,%d), got %s" % (self.in_h, self.in_w, tuple(shape)))

    def embed_host(self, X, want="codes", path=_lib.PATH_TCGEN05):
        """NumPy (or pinned torch CPU tensor) in, NumPy out.  want: 'codes', 'latents' or 'both'."""
        if hasattr(X, "data_ptr"):
            self._check_input(X.shape)
            n = X.shape[0]
            import torch
            dt = _lib.IN_U8 if X.dtype == torch.uint8 else _lib.IN_F32
            if dt == _lib.IN_F32 and X.dtype != torch.float32:
                X = X.float()
            X = X.contiguous()
            xptr = ctypes.c_void_p(X.data_ptr())
        else:
            X = np.asarray(X)
            self._check_input(X.shape)
            n = X.shape[0]
            dt = _lib.IN_U8 if X.dtype == np.uint8 else _lib.IN_F32
            X = np.ascontiguousarray(X, np.uint8 if dt == _lib.IN_U8 else np.float32)
            xptr = ctypes.c_void_p(X.ctypes.data)
        codes = np.empty((n, 32), np.float32) if want in ("codes", "both") else None
        lats = np.empty((n, 32), np.float32) if want in ("latents", "both") else None
        _lib.check(_lib.lib.asr_encoder_embed_host(self.handle, xptr, dt, n, _lib.dptr(codes), _lib.dptr(lats), path))
        if want == "codes":
            return codes
        if want == "latents":
            return lats
        return codes, lats

    def embed_device(self, X, codes=None, latents=None, path=_lib.PATH_TCGEN05, stream=None):
        """torch CUDA tensors in/out on torch's current stream; n <= max_batch per call."""
        import torch
        self._check_input(X.shape)
        dt = _lib.IN_U8 if X.dtype == torch.uint8 else _lib.IN_F32
        assert X.is_cuda and X.is_contiguous() and X.dtype in (torch.uint8, torch.float32)
        _lib.check(_lib.lib.asr_encoder_embed(self.handle, _lib.dptr(X), dt, X.shape[0], _lib.dptr(codes),
                                              _lib.dptr(latents), path, _lib.stream_ptr(stream)))
        return codes, latents

    def set_fusion(self, mask):
        """Bit 0: layers 0 + 1 as one kernel, bit 1: layers 2 + 3 (asr_encoder_set_fusion).  Returns the mask in effect."""
        _lib.check(_lib.lib.asr_encoder_set_fusion(self.handle, int(mask)))
        return self.fusion

    @property
    def fusion(self):
        return int(_lib.lib.asr_encoder_get_fusion(self.handle))

    def set_timing(self, enable):
        _lib.check(_lib.lib.asr_encoder_set_timing(self.handle, int(bool(enable))))

    def get_timing(self):
        """-> dict(ms_layer0, ms_conv_tc, ms_head, calls) summed since the last call."""
        a, b, c, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        _lib.check(_lib.lib.asr_encoder_get_timing(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                                                   ctypes.byref(n)))
        return dict(ms_layer0=a.value, ms_conv_tc=b.value, ms_head=c.value, calls=n.value)

    def debug_activation(self, layer, n, path=_lib.PATH_TCGEN05):
        c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib.asr_encoder_debug_activation(self.handle, layer, path, n, None, ctypes.byref(c),
                                                         ctypes.byref(h), ctypes.byref(w)))
        out = np.empty((n, c.value, h.value, w.value), np.float32)
        _lib.check(_lib.lib.asr_encoder_debug_activation(self.handle, layer, path, n, _lib.dptr(out), ctypes.byref(c),
                                                         ctypes.byref(h), ctypes.byref(w)))
        return out


class RetrievalNet(object):
    """Both branches + the CCA layer's state for one model family."""

    def __init__(self, model_name, filters, raw_shape_1, input_shape_1, input_shape_2, model_prepare_mode,
                 dim_latent=32, flip_filters=False, max_batch=MAX_BATCH_DEFAULT):
        self.model_name = model_name
        self.filters = tuple(filters)
        self._raw1, self._in1, self._in2 = tuple(raw_shape_1), tuple(input_shape_1), tuple(input_shape_2)
        self.model_prepare_mode = model_prepare_mode
        self.dim_latent = dim_latent
        self.flip_filters = flip_filters
        self.max_batch = max_batch
        self.params = None
        self._encoders = {}
        self.cca_layer = CCALayer(self)
        self.l_view1 = Layer(self, "input", 1, (None,) + self._in1)
        self.l_view2 = Layer(self, "input", 2, (None,) + self._in2)
        self.l_v1latent = Layer(self, "latent", 1, (None, dim_latent))
        self.l_v2latent = Layer(self, "latent", 2, (None, dim_latent))

    def layers(self):
        return self.l_view1, self.l_view2, self.l_v1latent, self.l_v2latent

    def raw_shape(self, view):
        return self._raw1 if view == 1 else self._in2

    def input_shape(self, view):
        return self._in1 if view == 1 else self._in2

    def set_params(self, params):
        params = [np.ascontiguousarray(p, np.float32) for p in params]
        shapes = P.expected_shapes(self.filters, self.dim_latent)
        if len(params) != len(shapes):
            raise ValueError("mismatch: parameter list has %d arrays, network expects %d" % (len(params), len(shapes)))
        for i, (p, s) in enumerate(zip(params, shapes)):
            if tuple(p.shape) != tuple(s):
                raise ValueError("mismatch: parameter %d has shape %s, network expects %s" % (i, p.shape, s))
        self.params = params
        for e in self._encoders.values():
            e.close()
        self._encoders = {}

    def _cca_changed(self):
        _, cca = P.split_params(self.params)
        for (view, _mode), e in self._encoders.items():
            e.set_cca(cca["mean1" if view == 1 else "mean2"], cca["U" if view == 1 else "V"])

    def encoder(self, view, prepare_mode=_lib.PREP_NONE):
        if self.params is None:
            raise RuntimeError("parameters not set (call set_all_param_values first)")
        key = (view, prepare_mode)
        if key not in self._encoders:
            self._encoders[key] = Encoder(self, view, prepare_mode, self.max_batch)
        return self._encoders[key]


# ---- the Lasagne helper names the scripts call -------------------------------------------
def _net_of(layers):
    layer = layers[0] if isinstance(layers, (list, tuple)) else layers
    return layer.net


def set_all_param_values(layers, values):
    """lasagne.layers.set_all_param_values (run_eval.py:82, refine_cca.py:58, retrieval_wrapper.py:29)."""
    _net_of(layers).set_params(values)


def get_all_param_values(layers):
    """lasagne.layers.get_all_param_values (refine_cca.py:111)."""
    return [p.copy() for p in _net_of(layers).params]


def get_all_layers(layer):
    """lasagne.layers.helper.get_all_layers, reduced to what refine_cca.py:78-84 needs: the
    traversal contains the CCALayer."""
    net = _net_of(layer)
    return [net.l_view1, net.l_view2, net.cca_layer, layer]


def compile_function(inputs, output_layer, prepare_mode=_lib.PREP_NONE, path=_lib.PATH_TCGEN05):
    """Stand-in for `theano.function(inputs, lasagne.layers.get_output(layer, deterministic=True))`.

    Two-input functions (run_eval.py:91-95, retrieval_wrapper.py:33-38) take (X1, X2) and use only
    the view the output layer belongs to: in deterministic mode the CCALayer's two halves are
    independent (layers/cca.py:188-199), so the other input is never evaluated."""
    net, view = output_layer.net, output_layer.view
    want = "codes" if output_layer.kind == "latent" else "latents"
    n_inputs = len(inputs)

    def fn(*args):
        if len(args) != n_inputs:
            raise TypeError("expected %d inputs, got %d" % (n_inputs, len(args)))
        X = args[0] if n_inputs == 1 else args[view - 1]
        return net.encoder(view, prepare_mode).embed_host(X, want=want, path=path)

    def _mode(prepare):
        # raw input + the model's own prepare -> fused on the device; view 2 is never prepared
        mode = getattr(prepare, "asr_prepare_mode", None) if prepare is not None else None
        return mode if (mode is not None and view == 1) else prepare_mode

    def fused(X, prepare=None):
        return net.encoder(view, _mode(prepare)).embed_host(X, want=want, path=path)

    def fused2(X1, X2, prepare1=None):
        return fused(X1 if view == 1 else X2, prepare1)

    fn.asr_fused, fn.asr_fused2 = fused, fused2
    return fn
