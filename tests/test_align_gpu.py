"""DTW alignment through the C ABI vs the reference's golden outputs / the oracle."""
import numpy as np
import pytest

from oracle import align as oalign

pytestmark = pytest.mark.gpu


def test_dtw_bit_exact_vs_reference_golden(golden):
    from audio_sheet_retrieval_b200.utils.dtw_by_dist import dtw_by_dist
    for t in "ab":                                           # "b" takes the transposition branch
        md, C, D1, path = dtw_by_dist(golden["dtw_%s_dist" % t].copy())
        assert md == golden["dtw_%s_min" % t][0]
        assert (D1 == golden["dtw_%s_acc" % t]).all()
        assert (path[0] == golden["dtw_%s_p0" % t]).all() and (path[1] == golden["dtw_%s_p1" % t]).all()


def test_compute_alignment_matches_reference_golden(golden):
    from audio_sheet_retrieval_b200.utils.alignment import compute_alignment, cosine_distances, estimate_alignment_error
    sheet_idxs, spec_idxs = np.arange(120) * 5 + 100, np.arange(90) * 2 + 10
    d = cosine_distances(golden["al_img"], golden["al_spec"])
    np.testing.assert_allclose(d, golden["al_dists"], atol=1e-13)      # fp64, sequential sums
    mapping, res = compute_alignment(golden["al_img"], golden["al_spec"], sheet_idxs, spec_idxs, "pydtw")
    assert (res["aligned_sheet_idxs"] == golden["al_aligned"]).all()
    np.testing.assert_allclose([mapping[k] for k in sorted(mapping)], golden["al_map_v"], atol=1e-9)
    err = estimate_alignment_error(sheet_idxs[golden["al_true"]].astype(float), spec_idxs, mapping)
    np.testing.assert_allclose(err, golden["al_err"], atol=1e-9)
    _, res_b = compute_alignment(golden["al_img"], golden["al_spec"], sheet_idxs, spec_idxs, "baseline")
    assert (res_b["aligned_sheet_idxs"] == golden["al_b_aligned"]).all()


@pytest.mark.parametrize("r,c", [(1, 1), (1, 7), (9, 1), (300, 211), (150, 400), (1500, 1100)])
def test_dtw_vs_oracle_sizes(r, c):
    from audio_sheet_retrieval_b200.utils.dtw_by_dist import dtw_by_dist
    rng = np.random.RandomState(r * 7 + c)
    d = np.round(np.abs(rng.normal(size=(r, c))), 2)         # coarse values -> many exact ties on the path
    md, C, D1, path = dtw_by_dist(d.copy())
    if r * c <= 100000:
        md_o, C_o, D1_o, path_o = oalign.dtw_by_dist(d.copy())
        assert md == md_o and (D1 == D1_o).all()
        assert (path[0] == path_o[0]).all() and (path[1] == path_o[1]).all()
    # size-independent properties: monotone unit steps from corner to corner, cost consistent with the path
    p, q = (path[1], path[0]) if c <= r else (path[0], path[1])      # undo the reference's swap: p = rows of D1
    rr, cc = D1.shape
    assert p[0] == 0 and q[0] == 0 and p[-1] == rr - 1 and q[-1] == cc - 1
    dp, dq = np.diff(p), np.diff(q)
    assert ((dp >= 0) & (dq >= 0) & (dp <= 1) & (dq <= 1) & (dp + dq >= 1)).all()
    np.testing.assert_allclose(C[p, q].sum(), D1[-1, -1], rtol=1e-12)
