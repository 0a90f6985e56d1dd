#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE's own NumPy/SciPy code.

Run in the build container only (needs /root/reference; the GPU box has none):
    python tests/golden/make_golden.py

The reference is Python 2 and imports Theano/Lasagne at module level, so modules
cannot be imported.  Instead the source text of the pure NumPy functions on the
hot path is sliced out of the reference files *where they lie* and executed
under Python 3 with these shims only:
    xrange -> range, np.int/np.float -> int/float, `print "x"` -> print("x"),
    the two Python-2 integer divisions in eval_retrieval (`/` -> `//`).
No reference source is copied into this repository; only the numeric outputs
are committed (tests/golden/*.npz) and checked by tests/test_oracle_golden.py.
"""
import os
import re
import sys
import textwrap
import types

import numpy as np

REF = os.environ.get("ASR_REFERENCE", "/root/reference")
ASR = os.path.join(REF, "audio_sheet_retrieval")
OUT = os.path.dirname(os.path.abspath(__file__))


class _NP(object):
    """numpy with the aliases NumPy 1.13 still had."""
    int = int
    float = float

    def __getattr__(self, name):
        return getattr(np, name)


NP = _NP()


def _src(path):
    with open(os.path.join(ASR, path)) as fp:
        return fp.read()


def _slice_def(src, name, indent=""):
    """Text of `def name` up to the next def/class at the same indent."""
    stop = r"^%s(?:def |class |if __name__)" % indent
    if indent:
        stop += r"|^\S"                      # a method also ends where the class body ends
    pat = re.compile(r"^%sdef %s\(.*?(?=%s|\Z)" % (indent, name, stop), re.S | re.M)
    m = pat.search(src)
    assert m, name
    return textwrap.dedent(m.group(0))


def _py3(src):
    src = re.sub(r'^(\s*)print (".*)$', r"\1print(\2)", src, flags=re.M)
    return src


def load_eval_retrieval():
    code = _slice_def(_src("utils/train_dcca_pool.py"), "eval_retrieval")
    code = code.replace("k = n_v2 / n_v1", "k = n_v2 // n_v1").replace("h = n_v1 / n_v2", "h = n_v1 // n_v2")
    g = {"np": NP, "xrange": range}
    exec(compile(code, "ref:eval_retrieval", "exec"), g)
    return g["eval_retrieval"]


def load_cca_class():
    code = _py3(_src("utils/cca.py"))
    g = {"__name__": "ref_cca"}
    exec(compile(code, "ref:utils/cca.py", "exec"), g)
    return g["CCA"]


def load_batch_compute():
    src = _src("utils/batch_iterators.py")
    g = {"np": NP, "xrange": range, "sys": sys, "print_function": None}
    for name in ("batch_compute1", "batch_compute2"):
        exec(compile(_slice_def(src, name), "ref:" + name, "exec"), g)
    return g["batch_compute1"], g["batch_compute2"]


def load_server_methods():
    from scipy.spatial.distance import cdist
    src = _src("audio_sheet_server.py")
    g = {"np": NP, "cdist": cdist, "col": types.SimpleNamespace(print_colored=lambda s, color=None: s, UNDERLINE=0)}
    for name in ("_retrieve_sheet_snippet_ids", "_retrieve_perform_excerpt_ids", "detect_score", "detect_performance"):
        exec(compile(_slice_def(src, name, indent="    "), "ref:" + name, "exec"), g)
    return g


def main():
    rng = np.random.RandomState(1234)
    out = {}

    # ---- eval_retrieval (asr/utils/train_dcca_pool.py:28-82) ----
    eval_retrieval = load_eval_retrieval()

    def unit(x):
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)

    base = rng.normal(size=(300, 32))
    lv1 = unit(base + 0.9 * rng.normal(size=base.shape))
    lv2 = unit(base + 0.9 * rng.normal(size=base.shape))
    mr, med, md, hr, mrr = eval_retrieval(lv1, lv2)
    out["er_lv1"], out["er_lv2"] = lv1, lv2
    out["er_res"] = np.array([mr, med, md, hr[1], hr[5], hr[10], hr[25], mrr], np.float64)
    # grouped case: 3 view-2 items per view-1 item
    lv2g = unit(np.repeat(base[:100], 3, axis=0) + 1.0 * rng.normal(size=(300, 32)))
    mr, med, md, hr, mrr = eval_retrieval(lv1[:100], lv2g)
    out["er_g_lv1"], out["er_g_lv2"] = lv1[:100], lv2g
    out["er_g_res"] = np.array([mr, med, 0.0, hr[1], hr[5], hr[10], hr[25], mrr], np.float64)
    # clipped dims (run_eval.py:160-162 --max_dim)
    mr, med, md, hr, mrr = eval_retrieval(lv1[:, :8], lv2[:, :8])
    out["er_c_res"] = np.array([mr, med, md, hr[1], hr[5], hr[10], hr[25], mrr], np.float64)

    # ---- CCA('svd') fit / transform (asr/utils/cca.py:25-53,199-211,432-444) ----
    CCA = load_cca_class()
    sys.path.insert(0, os.path.join(OUT, "..", ".."))
    from oracle.cca import synth_latents
    H1, H2 = synth_latents(2000, seed=23)
    cca = CCA(method="svd")
    coeffs = cca.fit(H1, H2, verbose=False)
    out["cca_H1"], out["cca_H2"] = H1, H2
    out["cca_m1"], out["cca_m2"] = cca.m1, cca.m2
    out["cca_U"], out["cca_V"], out["cca_sigma"] = cca.U, cca.V, coeffs
    out["cca_T1"] = cca.transform_V1(H1[:16])
    out["cca_T2"] = cca.transform_V2(H2[:16])

    # ---- batch_compute1 / batch_compute2 (asr/utils/batch_iterators.py:17-111) ----
    bc1, bc2 = load_batch_compute()
    Xa = rng.normal(size=(23, 1, 4, 5)).astype(np.float32)
    Xb = rng.normal(size=(23, 1, 3, 2)).astype(np.float32)
    calls = []

    def f1(E):
        calls.append(E.shape[0])
        return E.reshape(E.shape[0], -1)[:, :3] * 2.0

    def f2(E1, E2):
        return np.hstack((E1.reshape(E1.shape[0], -1)[:, :2], E2.reshape(E2.shape[0], -1)[:, :2]))

    out["bc_Xa"], out["bc_Xb"] = Xa, Xb
    out["bc1"] = bc1(Xa, f1, 10, prepare=lambda e: e + 1.0)
    out["bc1_calls"] = np.array(calls)
    out["bc2"] = bc2(Xa, Xb, f2, 10, prepare1=lambda e: e * 0.5)

    # ---- DB search + vote (asr/audio_sheet_server.py:213-300,530-563) ----
    g = load_server_methods()
    n_pieces, per = 12, 40
    db = unit(rng.normal(size=(n_pieces * per, 32)))
    ids = np.repeat(np.arange(n_pieces), per)
    id_to_piece = dict((i, "piece_%02d" % i) for i in range(n_pieces))
    T = 400
    spec = np.abs(rng.normal(size=(92, T))).astype(np.float32)
    # embedding stub: excerpt -> code by a fixed random projection, so the search path is exercised
    P = rng.normal(size=(92 * 42, 32)).astype(np.float32)
    true_rows = 5 * per + rng.randint(0, per, 100)

    class Embed(object):
        def compute_view_2(self, X):
            c = X.reshape(X.shape[0], -1).dot(P)
            c = 0.02 * unit(c) + db[true_rows[:X.shape[0]]]
            return unit(c)

        compute_view_1 = compute_view_2

    srv = types.SimpleNamespace(sheet_snippet_codes=db, sheet_snippet_ids=ids, id_to_piece=id_to_piece,
                                perform_excerpt_codes=db, perform_excerpt_ids=ids, id_to_perform=id_to_piece,
                                spec_shape=(92, 42), sheet_shape=(92, 42), embed_network=Embed())
    srv._retrieve_sheet_snippet_ids = lambda code, n_candidates=1: g["_retrieve_sheet_snippet_ids"](srv, code, n_candidates)
    srv._retrieve_perform_excerpt_ids = lambda code, n_candidates=1: g["_retrieve_perform_excerpt_ids"](srv, code, n_candidates)
    q = Embed().compute_view_2(np.stack([spec[None, :, s:s + 42] for s in
                                          np.linspace(0, T - 42, 100).astype(int)]))
    pid, sidx = g["_retrieve_sheet_snippet_ids"](srv, q[3:4], 25)
    out["srv_db"], out["srv_ids"], out["srv_spec"], out["srv_q"] = db, ids, spec, q
    out["srv_pid"], out["srv_sidx"] = pid, sidx
    names, votes = g["detect_score"](srv, spec, top_k=5, n_candidates=25)
    out["srv_names"] = np.array([int(n.split("_")[1]) for n in names])
    out["srv_votes"] = votes
    names, votes = g["detect_performance"](srv, spec, top_k=5, n_candidates=10)
    out["srv_p_names"] = np.array([int(n.split("_")[1]) for n in names])
    out["srv_p_votes"] = votes

    # ---- DTW alignment (asr/utils/dtw_by_dist.py:6-34,69-83 ; asr/utils/alignment.py:113-174) ----
    gd = {"__name__": "ref_dtw"}
    exec(compile(_src("utils/dtw_by_dist.py"), "ref:utils/dtw_by_dist.py", "exec"), gd)
    ref_dtw = gd["dtw_by_dist"]
    for tag, shape in (("a", (60, 45)), ("b", (40, 70))):            # "b" exercises the transposition branch
        d = np.abs(rng.normal(size=shape)) + 0.3 * np.abs(np.subtract.outer(np.arange(shape[0]) / shape[0],
                                                                             np.arange(shape[1]) / shape[1]))
        md, C, D1, path = ref_dtw(d.copy())
        out["dtw_%s_dist" % tag] = d
        out["dtw_%s_min" % tag] = np.array([md])
        out["dtw_%s_acc" % tag] = D1
        out["dtw_%s_p0" % tag], out["dtw_%s_p1" % tag] = np.asarray(path[0]), np.asarray(path[1])
    from scipy.spatial.distance import cdist as _cdist
    from scipy.interpolate import interp1d as _interp1d
    sys.modules["dtw_by_dist"] = types.SimpleNamespace(dtw_by_dist=ref_dtw)     # `from dtw_by_dist import ...`
    ga = {"np": NP, "xrange": range, "cdist": _cdist, "interp1d": _interp1d}
    asrc = _src("utils/alignment.py")
    for name in ("align_baseline", "align_pydtw", "compute_alignment", "estimate_alignment_error"):
        exec(compile(_slice_def(asrc, name), "ref:" + name, "exec"), ga)
    from oracle.align import synth_alignment_problem
    img, spec, true_idx = synth_alignment_problem(120, 90, seed=3)
    sheet_idxs = np.arange(120) * 5 + 100
    spec_idxs = np.arange(90) * 2 + 10
    mapping, res = ga["compute_alignment"](img, spec, sheet_idxs, spec_idxs, "pydtw")
    out["al_img"], out["al_spec"], out["al_true"] = img, spec, true_idx
    out["al_dists"] = res["dists"]
    out["al_aligned"] = res["aligned_sheet_idxs"]
    out["al_map_k"] = np.array(sorted(mapping.keys()))
    out["al_map_v"] = np.array([mapping[k] for k in sorted(mapping.keys())])
    err = ga["estimate_alignment_error"](sheet_idxs[true_idx].astype(float), spec_idxs, mapping)
    out["al_err"] = err
    mapping_b, res_b = ga["compute_alignment"](img, spec, sheet_idxs, spec_idxs, "baseline")
    out["al_b_aligned"] = res_b["aligned_sheet_idxs"]

    np.savez_compressed(os.path.join(OUT, "reference_numpy_paths.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_numpy_paths.npz"), sorted(out))


if __name__ == "__main__":
    main()
