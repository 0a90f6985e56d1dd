#!/usr/bin/env python
"""Snippet-retrieval evaluation with the reference's CLI and outputs
(audio_sheet_retrieval/run_eval.py:34-212): load params, embed n_test pairs, optional direction
flip / dimension clip, eval_retrieval, print R@k / MRR / MR, dump YAML.

    python -m audio_sheet_retrieval_b200.run_eval --model mutopia_ccal_cont_rsz --data mutopia \
        --train_split all_split.yaml --config mutopia_full_aug.yaml --n_test 2000 [--estimate_UV] \
        [--V2_to_V1] [--max_dim D] [--dump_results] [--param_file path.pkl]
"""
from __future__ import print_function

import argparse
import os

import numpy as np
import yaml

from . import network
from .config.settings import EXP_ROOT
from .params import load_params
from .run_train import SYNTHETIC_BANNER, compile_tag, is_synthetic, select_data, select_model
from .utils.batch_iterators import batch_compute2
from .utils.train_dcca_pool import eval_retrieval, retrieval_ranks


def flip_variables(v1, v2):
    """ flip variables """
    tmp = v1.copy()
    v1 = v2
    v2 = tmp
    return v1, v2


def main(argv=None):
    parser = argparse.ArgumentParser(description='Evaluate cross-modality retrieval model.')
    parser.add_argument('--model', help='select model to evaluate.')
    parser.add_argument('--data', help="select evaluation data ('mutopia' = MSMD, 'synthetic').", type=str)
    parser.add_argument('--show', help='show evaluation plots.', action='store_true')
    parser.add_argument('--n_test', help='number of test samples used.', type=int, default=None)
    parser.add_argument('--V2_to_V1', help='query direction.', action='store_true')
    parser.add_argument('--estimate_UV', help='load re-estimated U and V.', action='store_true')
    parser.add_argument('--max_dim', help='maximum dimension of retrieval space.', type=int, default=None)
    parser.add_argument('--seed', help='query direction.', type=int, default=23)
    parser.add_argument('--train_split', help='path to train split file.', type=str, default=None)
    parser.add_argument('--config', help='path to experiment config file.', type=str, default=None)
    parser.add_argument('--dump_results', help='dump results of current run to file.', action='store_true')
    parser.add_argument('--param_file', help='explicit parameter pickle (overrides EXP_ROOT lookup).', default=None)
    args = parser.parse_args(argv)

    model, _ = select_model(args.model)
    if not hasattr(model, 'prepare'):
        model.prepare = None

    print("Building network %s ..." % model.EXP_NAME)
    layers = model.build_model(show_model=False)

    tag = compile_tag(args.train_split, args.config) if args.train_split and args.config else None
    print("Experimental Tag:", tag)

    if args.estimate_UV:
        model.EXP_NAME += "_est_UV"
    out_path = os.path.join(os.path.join(EXP_ROOT), model.EXP_NAME)
    dump_file = 'params.pkl' if tag is None else 'params_%s.pkl' % tag
    dump_file = os.path.join(out_path, dump_file)
    if args.param_file:
        dump_file = args.param_file

    print("\n")
    print("Loading model parameters from:", dump_file)
    params = load_params(dump_file)
    network.set_all_param_values(layers, params)

    print("\nLoading data...")
    data = select_data(args.data, args.train_split, args.config, args.seed, test_only=True)
    if is_synthetic(args.data):
        print(SYNTHETIC_BANNER)

    print("\nCompiling prediction functions...")
    l_view1, l_view2, l_v1latent, l_v2latent = layers
    input_1 = input_2 = [l_view1.input_var, l_view2.input_var]
    compute_v1_latent = network.compile_function(input_1, l_v1latent)
    compute_v2_latent = network.compile_function(input_2, l_v2latent)

    print("Evaluating on test set...")
    eval_set = 'test'
    n_test = args.n_test if args.n_test is not None else data[eval_set].shape[0]
    indices = np.linspace(0, data[eval_set].shape[0] - 1, n_test).astype(int)
    X1, X2 = data[eval_set][indices]

    print("Computing embedding space...")
    lv1 = batch_compute2(X1, X2, compute_v1_latent, np.min([100, n_test]), prepare1=model.prepare)
    lv2 = batch_compute2(X1, X2, compute_v2_latent, np.min([100, n_test]), prepare1=model.prepare)
    lv1_cca, lv2_cca = lv1, lv2

    if args.V2_to_V1:
        lv1_cca, lv2_cca = flip_variables(lv1_cca, lv2_cca)
    n_test = lv1_cca.shape[0]

    max_dim = args.max_dim if args.max_dim is not None else lv1_cca.shape[1]
    lv1_cca = lv1_cca[:, 0:max_dim]
    lv2_cca = lv2_cca[:, 0:max_dim]

    print("V1.shape:", lv1_cca.shape)
    print("V2.shape:", lv2_cca.shape)

    print("Computing performance measures...")
    mean_rank_te, med_rank_te, dist_te, hit_rates, map_ = eval_retrieval(lv1_cca, lv2_cca)

    recall_at_k = dict()
    print("\nHit Rates:")
    for key in np.sort(list(hit_rates.keys())):
        recall_at_k[key] = float(100 * hit_rates[key]) / n_test
        pk = recall_at_k[key] / key
        print("Top %02d: %.3f (%d) %.3f" % (key, recall_at_k[key], hit_rates[key], pk))

    print("\n")
    print("Median Rank: %.2f (%d)" % (med_rank_te, lv2_cca.shape[0]))
    print("Mean Rank  : %.2f (%d)" % (mean_rank_te, lv2_cca.shape[0]))
    print("Mean Dist  : %.5f " % dist_te)
    print("MAP        : %.3f " % map_)

    _, ts = retrieval_ranks(lv1_cca, lv2_cca)
    dists = 1.0 - ts.astype(np.float64)
    print("Min Dist   : %.5f " % np.min(dists))
    print("Max Dist   : %.5f " % np.max(dists))
    print("Med Dist   : %.5f " % np.median(dists))

    results = {"map": float(map_), 'med_rank': float(med_rank_te),
               'recall_at_k': dict(("%d" % k, v) for k, v in recall_at_k.items())}
    if is_synthetic(args.data):
        results["data"] = "synthetic"
    if args.dump_results:
        ret_dir = "A2S" if args.V2_to_V1 else "S2A"
        if is_synthetic(args.data):
            ret_dir += "_synthetic"
        res_file = dump_file.replace("params_", "eval_").replace(".pkl", "_%s.yaml")
        res_file = res_file % ret_dir
        with open(res_file, 'w') as fp:
            yaml.dump(results, fp, default_flow_style=False)
    return results


if __name__ == '__main__':
    main()
