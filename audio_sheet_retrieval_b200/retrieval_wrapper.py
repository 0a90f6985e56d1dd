"""RetrievalWrapper: the embed API of the reference (audio_sheet_retrieval/retrieval_wrapper.py:12-77)
on top of libasr_b200.so.

Same constructor, methods, attributes and return types.  Differences that are not observable from
the results: each branch is evaluated once (the reference evaluates both branches and feeds zeros
to the unused one, :58-61,74-77), `prepare_view_1 = model.prepare` is fused into layer 0 on the
device, and batching (<= 100 rows per Theano call in the reference) happens inside the library.
"""
from __future__ import print_function

import numpy as np

from . import network
from .params import load_params
from .utils.batch_iterators import batch_compute2


class RetrievalWrapper(object):
    """ Wrapper for cross modality retrieval networks """

    def __init__(self, model, param_file, prepare_view_1=None, prepare_view_2=None):
        """ Constructor """
        self.prepare_view_1 = prepare_view_1
        self.prepare_view_2 = prepare_view_2

        self.code_dim = model.DIM_LATENT

        print("Building network ...")
        layers = model.build_model(show_model=False)

        print("Loading model parameters from:", param_file)
        params = load_params(param_file)
        network.set_all_param_values(layers, params)

        print("Compiling prediction functions ...")
        l_view1, l_view2, l_v1latent, l_v2latent = layers
        self.compute_v1_latent = network.compile_function([l_view1.input_var, l_view2.input_var], l_v1latent)
        self.compute_v2_latent = network.compile_function([l_view1.input_var, l_view2.input_var], l_v2latent)

        # dummy inputs for the respective second view (kept for interface compatibility)
        self.dummy_in_v1 = np.zeros(([1] + list(l_view1.output_shape[1:])), dtype=np.float32)
        self.dummy_in_v2 = np.zeros(([1] + list(l_view2.output_shape[1:])), dtype=np.float32)

        self.shape_view1 = l_view1.output_shape[1:]
        self.shape_view2 = l_view2.output_shape[1:]
        self.net = l_view1.net

    def compute_view_1(self, X):
        """ compute network output of view 1 """
        X = np.asarray(X)
        dummy_in_v2 = np.broadcast_to(self.dummy_in_v2, (X.shape[0],) + self.dummy_in_v2.shape[1:])
        return batch_compute2(X, dummy_in_v2, self.compute_v1_latent,
                              batch_size=min(100, X.shape[0]),
                              prepare1=self.prepare_view_1)

    def compute_view_2(self, Z):
        """ compute network output of view 2 """
        Z = np.asarray(Z)
        dummy_in_v1 = np.broadcast_to(self.dummy_in_v1, (Z.shape[0],) + self.dummy_in_v1.shape[1:])
        return batch_compute2(dummy_in_v1, Z, self.compute_v2_latent,
                              batch_size=min(100, Z.shape[0]),
                              prepare2=self.prepare_view_2)
