#!/usr/bin/env python
"""CCA refit with the reference's CLI and outputs (audio_sheet_retrieval/refine_cca.py:24-111):
embed n_train training pairs up to the CCA layer's inputs, CCA('svd').fit, overwrite
mean1/mean2/U/V, dump `<model>_est_UV/params_<tag>.pkl`.

With torchrun (WORLD_SIZE > 1) the n_train rows are sharded over the ranks, the covariance sums
are all-reduced with NCCL and rank 0 writes the pickle.
"""
from __future__ import print_function

import argparse
import os

import numpy as np

from . import network
from .config.settings import EXP_ROOT
from .network import CCALayer as CCALayer1
from .params import load_params, save_params
from .run_train import SYNTHETIC_BANNER, compile_tag, is_synthetic, select_data, select_model
from .utils.batch_iterators import batch_compute1
from .utils.cca import CCA


def refit(layers, X1, X2, prepare, group=None, verbose=True):
    """The refit proper (refine_cca.py:78-107) on this rank's rows; returns the fitted CCA."""
    l_view1, l_view2, l_v1latent, l_v2latent = layers
    input_1, input_2 = [l_view1.input_var], [l_view2.input_var]
    cca_layer = None
    for l in network.get_all_layers(l_v1latent):
        if isinstance(l, CCALayer1):
            cca_layer = l
            l_v1_cca = cca_layer.input_layers[0]
            l_v2_cca = cca_layer.input_layers[1]
            break
    compute_v1_latent = network.compile_function(input_1, l_v1_cca)
    compute_v2_latent = network.compile_function(input_2, l_v2_cca)
    n_train = X1.shape[0]
    lv1_tr = batch_compute1(X1, compute_v1_latent, np.min([10, n_train]), prepare=prepare)
    lv2_tr = batch_compute1(X2, compute_v2_latent, np.min([10, n_train]))
    cca = CCA(method='svd')
    cca.fit(lv1_tr, lv2_tr, verbose=verbose, group=group)
    cca_layer.mean1.set_value(cca.m1.astype(np.float32))
    cca_layer.mean2.set_value(cca.m2.astype(np.float32))
    cca_layer.U.set_value(cca.U.astype(np.float32))
    cca_layer.V.set_value(cca.V.astype(np.float32))
    return cca


def main(argv=None):
    parser = argparse.ArgumentParser(description='Train model.')
    parser.add_argument('--model', help='model parameters for evaluation.', default="mutopia_ccal_cont_rsz")
    parser.add_argument('--data', help="select evaluation data ('mutopia' = MSMD, 'synthetic').", type=str, default="mutopia")
    parser.add_argument('--n_train', help='number of train samples used for projection.', type=int, default=1000)
    parser.add_argument('--seed', help='query direction.', type=int, default=23)
    parser.add_argument('--train_split', help='path to train split file.', type=str, default=None)
    parser.add_argument('--config', help='path to experiment config file.', type=str, default=None)
    parser.add_argument('--param_file', help='explicit parameter pickle (overrides EXP_ROOT lookup).', default=None)
    parser.add_argument('--out_file', help='explicit output pickle.', default=None)
    args = parser.parse_args(argv)

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        group = dist.group.WORLD

    model, _ = select_model(args.model)
    if not hasattr(model, 'prepare'):
        model.prepare = None
    print("Building network %s ..." % model.EXP_NAME)
    layers = model.build_model(show_model=False)
    tag = compile_tag(args.train_split, args.config) if args.train_split and args.config else None
    print("Experimental Tag:", tag)
    out_path = os.path.join(os.path.join(EXP_ROOT), model.EXP_NAME)
    dump_file_name = 'params.pkl' if tag is None else 'params_%s.pkl' % tag
    dump_file = args.param_file or os.path.join(out_path, dump_file_name)
    print("\n")
    print("Loading model parameters from:", dump_file)
    network.set_all_param_values(layers, load_params(dump_file))

    model.EXP_NAME += "_est_UV"
    out_path = os.path.join(os.path.join(EXP_ROOT), model.EXP_NAME)
    if is_synthetic(args.data):      # never overwrite a real refit with one made on synthetic pairs
        dump_file_name = dump_file_name.replace(".pkl", "_synthetic.pkl")
    dump_file = args.out_file or os.path.join(out_path, dump_file_name)
    if rank == 0 and not os.path.exists(os.path.dirname(dump_file)):
        os.makedirs(os.path.dirname(dump_file))

    print("\nLoading data...")
    data = select_data(args.data, args.train_split, args.config, args.seed)
    if is_synthetic(args.data) and rank == 0:
        print(SYNTHETIC_BANNER)
    lo, hi = args.n_train * rank // world, args.n_train * (rank + 1) // world
    print("Computing train output...")
    X1, X2 = data['train'][lo:hi]
    print("Fitting CCA model...")
    refit(layers, X1, X2, model.prepare, group=group, verbose=(rank == 0))
    if rank == 0:
        print("Dumping refined model...")
        save_params(dump_file, network.get_all_param_values(layers))
    return dump_file


if __name__ == '__main__':
    main()
