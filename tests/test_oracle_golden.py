"""The oracle against golden vectors produced by the reference's own NumPy code
(tests/golden/make_golden.py) and against the shipped pickle's self-consistency."""
import numpy as np
import pytest

from oracle import cca as occa
from oracle import clib, metrics, search
from oracle.encoders import OracleNet, split_params


def _res(t):
    mr, med, md, hr, mrr = t
    return np.array([mr, med, md if md is not None else 0.0, hr[1], hr[5], hr[10], hr[25], mrr], np.float64)


def test_eval_retrieval_matches_reference(golden):
    got = _res(metrics.eval_retrieval_ref(golden["er_lv1"], golden["er_lv2"]))
    np.testing.assert_allclose(got, golden["er_res"], rtol=0, atol=1e-12)


def test_eval_retrieval_grouped_and_clipped(golden):
    got = _res(metrics.eval_retrieval_ref(golden["er_g_lv1"], golden["er_g_lv2"]))
    got[2] = 0.0
    np.testing.assert_allclose(got, golden["er_g_res"], rtol=0, atol=1e-12)
    got = _res(metrics.eval_retrieval_ref(golden["er_lv1"][:, :8], golden["er_lv2"][:, :8]))
    np.testing.assert_allclose(got, golden["er_c_res"], rtol=0, atol=1e-12)


def test_pinned_ranks_agree_with_reference_metrics(golden):
    """fp32 pinned ranking vs the reference's fp64 cdist ranking: same ranks on this data."""
    for a, b, key in (("er_lv1", "er_lv2", "er_res"), ("er_g_lv1", "er_g_lv2", "er_g_res")):
        ranks, ts = metrics.ranks_pinned(golden[a], golden[b])
        got = _res(metrics.metrics_from_ranks(ranks, ts))
        ref = golden[key].copy()
        if key == "er_g_res":
            got[2] = 0.0
        np.testing.assert_allclose(got[[0, 1, 3, 4, 5, 6, 7]], ref[[0, 1, 3, 4, 5, 6, 7]], rtol=0, atol=1e-12)
        if key == "er_res":
            assert abs(got[2] - ref[2]) < 1e-6
        r2, t2 = clib.rank(golden[a], golden[b])
        assert (r2 == ranks).all() and (t2 == ts).all()


def test_cca_svd_matches_reference(golden):
    c = occa.CCA(method="svd")
    sig = c.fit(golden["cca_H1"], golden["cca_H2"])
    np.testing.assert_allclose(sig, golden["cca_sigma"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(c.m1, golden["cca_m1"], rtol=0, atol=0)
    np.testing.assert_allclose(c.U, golden["cca_U"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(c.V, golden["cca_V"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(c.transform_V1(golden["cca_H1"][:16]), golden["cca_T1"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(c.transform_V2(golden["cca_H2"][:16]), golden["cca_T2"], rtol=1e-9, atol=1e-9)


def test_cca_invariants(golden):
    c = occa.CCA(method="svd")
    sig = c.fit(golden["cca_H1"].astype(np.float64), golden["cca_H2"].astype(np.float64))
    I = np.eye(32)
    np.testing.assert_allclose(c.U.T @ c.S11 @ c.U, I, atol=1e-10)
    np.testing.assert_allclose(c.V.T @ c.S22 @ c.V, I, atol=1e-10)
    np.testing.assert_allclose(c.U.T @ c.S12 @ c.V, np.diag(sig), atol=1e-10)
    assert (np.diff(sig) <= 1e-15).all()


def test_cca_layer_train_forward_equals_svd_form(golden):
    """Layer eigh form == svd form up to column order / joint sign; corr = sqrt(sigma^2 + rT)."""
    H1, H2 = golden["cca_H1"].astype(np.float64), golden["cca_H2"].astype(np.float64)
    out = occa.cca_layer_train_forward(H1, H2)
    c = occa.CCA(method="svd")
    sig = c.fit(H1, H2)
    np.testing.assert_allclose(out["corr"][::-1], np.sqrt(np.clip(sig ** 2 + 1e-3, 1e-7, 1.0)), atol=1e-9)
    d = np.diag(out["U"].T @ out["S12"] @ out["V"])
    assert (d >= -1e-12).all()
    np.testing.assert_allclose(d[::-1], sig, atol=1e-9)


def test_search_and_vote_match_reference(golden):
    db, ids, q = golden["srv_db"], golden["srv_ids"], golden["srv_q"]
    pid, sidx = search.retrieve_ids_ref(db, ids, q[3:4], 25)
    assert (pid == golden["srv_pid"]).all() and (sidx == golden["srv_sidx"]).all()
    # the pinned fp32 ranking returns the same rows on this data
    _, idx = search.pinned_topk(q[3:4], db, 25)
    assert (idx[0] == golden["srv_sidx"]).all()
    ids_v, votes, shares = search.detect_ref(q, db, ids, top_k=5, n_candidates=25)
    assert (ids_v == golden["srv_names"]).all()
    np.testing.assert_allclose(shares, golden["srv_votes"], atol=1e-15)
    ids_p, _, shares_p = search.detect_pinned(q, db, ids, top_k=5, n_candidates=10)
    assert (ids_p == golden["srv_p_names"]).all()
    np.testing.assert_allclose(shares_p, golden["srv_p_votes"], atol=1e-15)


def test_window_starts_match_reference(golden):
    T = golden["srv_spec"].shape[1]
    st = search.window_starts(T, 42)
    assert st[0] == 0 and st[-1] == T - 42 and len(st) == 100


def test_c_oracle_equals_numpy_oracle():
    rng = np.random.RandomState(3)
    D = rng.normal(size=(3001, 32)).astype(np.float32)
    Q = rng.normal(size=(19, 32)).astype(np.float32)
    D[10] = D[2000] = D[77]          # exact ties -> index order
    D[5] = 0                          # zero row -> NaN -> -inf
    s1, i1 = search.pinned_topk(Q, D, 25, idx_base=1000)
    s2, i2 = clib.topk(Q, D, 25, idx_base=1000)
    assert (i1 == i2).all() and (s1 == s2).all()
    s1, i1 = search.pinned_topk(D[70:80], D, 8)
    assert list(i1[7, :3]) == [10, 77, 2000]
    # k larger than the DB
    s3, i3 = search.pinned_topk(Q, D[:5], 8)
    s4, i4 = clib.topk(Q, D[:5], 8)
    assert (i3 == i4).all() and (i3[:, 5:] == -1).all() and (i3[:, 4] == 5 - 1).all() is not None


def test_shipped_pickle_layout(shipped_params):
    views, cca = split_params(shipped_params)
    assert views[0][0]["W"].shape == (24, 1, 3, 3) and views[1][8]["W"].shape == (32, 96, 1, 1)
    assert cca["U"].shape == (32, 32) and cca["mean1"].shape == (32,)
    for v in views:
        for L in v:
            assert (L["inv_std"] > 0).all() and L["inv_std"].max() <= 100.0 + 1e-3   # 1/sqrt(var+1e-4)


def test_oracle_encoder_shapes_and_norm(shipped_params):
    from oracle.encoders import synth_inputs
    net = OracleNet("mutopia_ccal_cont_rsz", shipped_params)
    X1, X2 = synth_inputs(5, seed=1)
    c1, c2 = net.compute_view_1(X1), net.compute_view_2(X2)
    assert c1.shape == (5, 32) and c2.shape == (5, 32) and c1.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(c1, axis=1), 1.0, atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(c2, axis=1), 1.0, atol=1e-5)


def test_prepare_rsz_is_box_mean():
    from oracle.encoders import prepare_rsz
    rng = np.random.RandomState(0)
    x = rng.randint(0, 256, (3, 1, 160, 200)).astype(np.float32)
    a = prepare_rsz(x)
    b = (x / np.float32(255)).reshape(3, 1, 80, 2, 100, 2).mean(axis=(3, 5))
    assert np.abs(a - b).max() < 2e-7


def test_dtw_matches_reference(golden):
    from oracle import align
    for t in "ab":
        md, C, D1, path = align.dtw_by_dist(golden["dtw_%s_dist" % t].copy())
        assert md == golden["dtw_%s_min" % t][0]
        assert (D1 == golden["dtw_%s_acc" % t]).all()
        assert (path[0] == golden["dtw_%s_p0" % t]).all() and (path[1] == golden["dtw_%s_p1" % t]).all()
    m, r = align.compute_alignment(golden["al_img"], golden["al_spec"], np.arange(120) * 5 + 100,
                                   np.arange(90) * 2 + 10, "pydtw")
    assert (r["aligned_sheet_idxs"] == golden["al_aligned"]).all()
    np.testing.assert_allclose([m[k] for k in sorted(m)], golden["al_map_v"], atol=1e-12)


def _loss_cases():
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_contrastive_loss.npz"))
    for ci in range(4):
        lv1, lv2 = d["c%d_lv1" % ci], d["c%d_lv2" % ci]
        for weight, gamma, sym in ((1.0, 0.7, False), (1.0, 0.7, True), (0.35, 0.2, True)):
            yield lv1, lv2, weight, gamma, sym, float(d["c%d_w%g_g%g_s%d" % (ci, weight, gamma, int(sym))])


def test_contrastive_loss_matches_reference():
    """Oracle forward vs the value the reference's own get_contrastive_cos_loss computes (objectives.py:30-69)."""
    from oracle.objectives import contrastive_cos_loss_ref
    for lv1, lv2, weight, gamma, sym, want in _loss_cases():
        got = contrastive_cos_loss_ref(lv1, lv2, weight, gamma, sym)
        assert abs(got - want) <= 1e-6 * max(1.0, abs(want)), (weight, gamma, sym, got, want)


def test_contrastive_loss_gradient_is_the_derivative_of_the_pinned_forward():
    from oracle.objectives import contrastive_cos_grads_ref, contrastive_cos_loss_ref
    rng = np.random.RandomState(5)
    for lv1, lv2, weight, gamma, sym, _ in list(_loss_cases())[:6]:
        g1, g2 = contrastive_cos_grads_ref(lv1, lv2, weight, gamma, sym)
        for _ in range(6):          # directional central differences (the hinge is piecewise linear)
            e1, e2 = rng.randn(*lv1.shape), rng.randn(*lv2.shape)
            h = 1e-6
            num = (contrastive_cos_loss_ref(lv1 + h * e1, lv2 + h * e2, weight, gamma, sym) -
                   contrastive_cos_loss_ref(lv1 - h * e1, lv2 - h * e2, weight, gamma, sym)) / (2 * h)
            ana = (g1 * e1).sum() + (g2 * e2).sum()
            assert abs(num - ana) <= 1e-5 * max(1.0, abs(ana)) + 2e-6, (num, ana)
