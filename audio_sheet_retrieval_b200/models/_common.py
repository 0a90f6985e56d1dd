"""Shared pieces of the two model modules (the reference duplicates them per file)."""
import numpy as np

from .. import _lib
from ..network import RetrievalNet

SPEC_CONTEXT = 42      # audio_sheet_retrieval/utils/mutopia_data.py (SPEC_CONTEXT)
SPEC_BINS = 92


def make_build_model(module, num_filters_1, half_res):
    """Counterpart of `get_build_model(...)` (models/mutopia_ccal_cont_rsz.py:61-147)."""
    f = num_filters_1
    filters = (f, f, 2 * f, 2 * f, 4 * f, 4 * f, 4 * f, 4 * f)
    raw1 = tuple(module.INPUT_SHAPE_1)
    in1 = (raw1[0], raw1[1] // 2, raw1[2] // 2) if half_res else raw1
    in2 = tuple(module.INPUT_SHAPE_2)

    def model(show_model):
        net = RetrievalNet(module.__name__.split(".")[-1], filters, raw1, in1, in2,
                           _lib.PREP_SCALE_HALF if half_res else _lib.PREP_SCALE,
                           dim_latent=module.DIM_LATENT)
        if show_model:
            print_architecture(net)
        return net.layers()

    return model


def print_architecture(net):
    """utils/monitoring.py:9-32 prints the Lasagne layer table; this prints the same facts."""
    for view in (1, 2):
        c, h, w = net.input_shape(view)
        print("view %d: input (%d, %d, %d)" % (view, c, h, w))
        cin = c
        for l, cout in enumerate(net.filters):
            print("  conv3x3 %3d -> %3d + BN + ELU @ %dx%d%s" % (cin, cout, h, w, "  + maxpool2" if l % 2 else ""))
            if l % 2:
                h, w = h // 2, w // 2
            cin = cout
        print("  conv1x1 %3d -> %3d + BN, global mean, CCA projection, length norm" % (cin, net.dim_latent))


def prepare_scale(x, y=None):
    """models/mutopia_ccal_cont.py:170-190."""
    x = x.astype(np.float32)
    x /= 255
    return x if y is None else (x, y)


def prepare_scale_half(x, y=None):
    """models/mutopia_ccal_cont_rsz.py:170-190.  cv2.resize(INTER_LINEAR) to exactly half size is a
    2x2 box mean (max deviation 1 ulp); the same arithmetic runs fused on the device when this
    function is passed as `prepare_view_1`."""
    x = x.astype(np.float32)
    x /= 255
    h, w = x.shape[2] // 2, x.shape[3] // 2
    x = x[:, :, :2 * h, :2 * w].reshape(x.shape[0], x.shape[1], h, 2, w, 2)
    x = ((x[:, :, :, 0, :, 0] + x[:, :, :, 0, :, 1]) + (x[:, :, :, 1, :, 0] + x[:, :, :, 1, :, 1])) * np.float32(0.25)
    x = np.ascontiguousarray(x, np.float32)
    return x if y is None else (x, y)
