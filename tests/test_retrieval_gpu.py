"""Parity of the fused retrieval kernels (through the C ABI) with the pinned-order oracle.
Bit-exact: indices AND scores."""
import numpy as np
import pytest

from oracle import clib, metrics, search

pytestmark = pytest.mark.gpu


def _db(n, seed=0, unit=True):
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(n, 32)).astype(np.float32)
    if unit:
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d


@pytest.mark.parametrize("n,nq,k", [(1, 1, 1), (5, 3, 8), (255, 2, 25), (256, 17, 25), (257, 16, 1), (3001, 33, 25),
                                    (20000, 100, 25), (70000, 7, 128), (300000, 1, 25)])
def test_topk_bit_exact(n, nq, k):
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    D, Q = _db(n, 1), _db(nq, 2, unit=False)
    s_ref, i_ref = clib.topk(Q, D, k)
    db = EmbeddingDB(D)
    s, i = db.topk(Q, k)
    assert (i == i_ref).all(), "indices differ"
    assert (s.view(np.uint32) == s_ref.view(np.uint32)).all() or (s == s_ref).all()


def test_topk_ties_zero_rows_and_base():
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    D = _db(5000, 3)
    D[10] = D[2000] = D[4999] = D[77]          # exact duplicates -> ties resolve by index
    D[5] = 0                                   # zero row -> NaN -> -inf, never selected
    Q = np.concatenate([D[70:80], _db(6, 4)])
    s_ref, i_ref = search.pinned_topk(Q, D, 25, idx_base=12345)
    db = EmbeddingDB(D, idx_base=12345)
    s, i = db.topk(Q, 25)
    assert (i == i_ref).all() and (s == s_ref).all()
    assert list(i[7, :4] - 12345) == [10, 77, 2000, 4999]
    assert not (i - 12345 == 5).any()


def test_topk_unnormalised_inputs_and_no_normalise():
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    D, Q = _db(4000, 5, unit=False) * 3.7, _db(9, 6, unit=False) * 0.01
    db = EmbeddingDB(D)
    for norm in (True, False):
        s_ref, i_ref = clib.topk(Q, D, 10, normalise=norm)
        s, i = db.topk(Q, 10, normalise=norm)
        assert (i == i_ref).all() and (s == s_ref).all()


def test_topk_clipped_dims_match_oracle():
    """--max_dim clipping (run_eval.py:160-162): zero padding adds exactly nothing."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    D, Q = _db(3000, 7), _db(11, 8)
    s_ref, i_ref = clib.topk(Q[:, :8], D[:, :8], 5)
    s, i = EmbeddingDB(D[:, :8]).topk(Q[:, :8], 5)
    assert (i == i_ref).all() and (s == s_ref).all()


def test_topk_matches_reference_rows(golden):
    """Same rows as the reference's own cdist+argsort on the golden server fixture."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    db = EmbeddingDB(golden["srv_db"])
    _, i = db.topk(golden["srv_q"][3:4], 25)
    assert (i[0] == golden["srv_sidx"]).all()


def test_sharded_topk_merge_equals_global():
    import torch
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB, merge_topk_device
    D, Q = _db(10007, 9), _db(40, 10)
    D[9000] = D[100]
    k, world = 25, 4
    s_ref, i_ref = clib.topk(Q, D, k)
    ss, ii = [], []
    q = torch.as_tensor(Q).cuda()
    for r in range(world):
        lo, hi = len(D) * r // world, len(D) * (r + 1) // world
        s, i = EmbeddingDB(D[lo:hi], idx_base=lo).topk_device(q, k)
        ss.append(s); ii.append(i)
    S = torch.stack(ss, 1).reshape(len(Q), world * k)
    I = torch.stack(ii, 1).reshape(len(Q), world * k)
    s, i = merge_topk_device(S, I, world, k)
    assert (i.cpu().numpy() == i_ref).all() and (s.cpu().numpy() == s_ref).all()


@pytest.mark.parametrize("n1,n2", [(300, 300), (100, 300), (300, 100), (2000, 2000), (1, 1)])
def test_ranks_bit_exact(n1, n2):
    from audio_sheet_retrieval_b200.utils.train_dcca_pool import retrieval_ranks
    base = np.random.RandomState(11).normal(size=(max(n1, n2), 32))
    a = (base[:n1] if n1 <= n2 else base) + 0.8 * np.random.RandomState(12).normal(size=(n1, 32))
    b = (np.repeat(base[:n1], n2 // n1, axis=0) if n2 > n1 else base[:n2]) + 0.8 * np.random.RandomState(13).normal(size=(n2, 32))
    a, b = a.astype(np.float32), b.astype(np.float32)
    r_ref, t_ref = clib.rank(a, b)
    r, t = retrieval_ranks(a, b)
    assert (r == r_ref).all() and (t == t_ref).all()


def test_eval_retrieval_matches_reference_golden(golden):
    """R@k / MRR / mean+median rank equal the reference's own eval_retrieval on its golden data."""
    from audio_sheet_retrieval_b200.utils.train_dcca_pool import eval_retrieval
    for a, b, key in (("er_lv1", "er_lv2", "er_res"), ("er_g_lv1", "er_g_lv2", "er_g_res")):
        mr, med, md, hr, mrr = eval_retrieval(golden[a], golden[b])
        ref = golden[key]
        assert mr == ref[0] and med == ref[1] and mrr == pytest.approx(ref[7], abs=1e-12)
        assert [hr[1], hr[5], hr[10], hr[25]] == list(ref[3:7])
        if key == "er_res":
            assert md == pytest.approx(ref[2], abs=1e-6)
    mr, med, md, hr, mrr = eval_retrieval(golden["er_lv1"][:, :8], golden["er_lv2"][:, :8])
    ref = golden["er_c_res"]
    assert mr == ref[0] and med == ref[1] and [hr[1], hr[5], hr[10], hr[25]] == list(ref[3:7])


def test_vote_matches_oracle_and_reference(golden):
    import torch
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    srv = AudioSheetServer()
    n_pieces = int(golden["srv_ids"].max()) + 1
    names = dict((i, "piece_%02d" % i) for i in range(n_pieces))
    srv.set_sheet_db(golden["srv_db"], golden["srv_ids"], names)
    srv.set_audio_db(golden["srv_db"], golden["srv_ids"], names)
    pid, sidx = srv._retrieve_sheet_snippet_ids(golden["srv_q"][3:4], 25)
    assert (pid == golden["srv_pid"]).all() and (sidx == golden["srv_sidx"]).all()
    res, votes = srv._vote(srv._sheet_db, golden["srv_q"], names, 5, 25, False)
    assert [int(r.split("_")[1]) for r in res] == list(golden["srv_names"])
    np.testing.assert_allclose(votes, golden["srv_votes"], atol=1e-15)
    res, votes = srv._vote(srv._audio_db, golden["srv_q"], names, 5, 10, False)
    assert [int(r.split("_")[1]) for r in res] == list(golden["srv_p_names"])
    np.testing.assert_allclose(votes, golden["srv_p_votes"], atol=1e-15)


def test_piece_identification_batched_equals_oracle():
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    db, ids, q, true_piece = search.synth_piece_db(200, 50, 12, 100, sigma=0.25)
    srv = AudioSheetServer()
    srv.set_sheet_db(db, ids, dict((i, str(i)) for i in range(200)))
    got_ids, got_cnt = srv.identify_from_codes(q, 12, top_k=5, n_candidates=25)
    for r in range(12):
        ref_ids, ref_votes, _ = search.detect_pinned(q[r * 100:(r + 1) * 100], db, ids, top_k=5, n_candidates=25)
        assert list(got_ids[r][:len(ref_ids)]) == list(ref_ids)
        assert list(got_cnt[r][:len(ref_ids)]) == list(ref_votes)
    assert (got_ids[:, 0] == true_piece).mean() >= 0.9


def test_large_db_roundtrip_property():
    """Full-size style check without the oracle: every DB row queried against the DB returns itself first."""
    import torch
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    g = torch.Generator(device="cuda").manual_seed(1)
    D = torch.randn((1000000, 32), generator=g, device="cuda")
    db = EmbeddingDB(D)
    rows = torch.arange(0, 1000000, 9973, device="cuda")
    s, i = db.topk_device(D[rows].contiguous(), 4)
    assert (i[:, 0] == rows).all()
    assert (s[:, 0] > 0.9999).all() and (s[:, :-1] >= s[:, 1:]).all()


@pytest.mark.parametrize("path", ["exact", "tc"])
@pytest.mark.parametrize("n,nq,k", [(300, 5, 7), (5000, 130, 25), (66000, 257, 32), (1000, 3, 1)])
def test_topk_both_kernel_paths_bit_exact(monkeypatch, path, n, nq, k):
    """The tensor-core pre-filter path and the streaming fp32 path must both equal the oracle."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    monkeypatch.setenv("ASR_TOPK_PATH", path)
    D, Q = _db(n, 21), _db(nq, 22, unit=False)
    D[n // 2] = D[3]
    D[7] = 0
    s_ref, i_ref = clib.topk(Q, D, k)
    s, i = EmbeddingDB(D, idx_base=77).topk(Q, k)
    assert (i - 77 == i_ref).all() and (s == s_ref).all()


@pytest.mark.parametrize("n,nq,k", [(300001, 20, 25), (300001, 50, 32), (150000, 100, 25), (800003, 7, 1)])
def test_prefilter_few_queries_many_slices_bit_exact(monkeypatch, n, nq, k):
    """The pre-filter kernel's few-query regimes: the query tile replicated over the TMEM lane quarters (4x for
    <= 32 queries, 2x for <= 64), one list per (query, slice) with the bound the slices share (the m-th largest of
    the slices' published entries), a partial last tile, duplicated and zero rows."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    monkeypatch.setenv("ASR_TOPK_PATH", "tc")
    D, Q = _db(n, 41), _db(nq, 42, unit=False)
    D[n - 1] = D[5] = D[n // 3]
    D[11] = 0
    Q[0] = D[n // 3] * 3.0
    s_ref, i_ref = clib.topk(Q, D, k)
    s, i = EmbeddingDB(D, idx_base=5).topk(Q, k)
    assert (i - 5 == i_ref).all() and (s == s_ref).all()


def test_prefilter_identical_rows_pure_index_order(monkeypatch):
    """Every score equal: the shared bounds and the first-tile floor all coincide with the scores themselves; the
    result must be the k lowest indices, for every query."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    monkeypatch.setenv("ASR_TOPK_PATH", "tc")
    row = _db(1, 43)
    D = np.repeat(row, 70000, axis=0)
    Q = _db(40, 44, unit=False)
    s_ref, i_ref = clib.topk(Q, D, 25)
    s, i = EmbeddingDB(D).topk(Q, 25)
    assert (i == i_ref).all() and (s == s_ref).all()
    assert (i == np.arange(25)[None, :]).all()


@pytest.mark.parametrize("path", ["exact", "tc"])
def test_topk_near_ties_within_prefilter_margin(monkeypatch, path):
    """Adversarial for the approximate pre-filter: thousands of rows whose exact scores differ by far
    less than the tf32 error (clusters of 1e-5-perturbed copies).  Exact re-scoring must still order them."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    monkeypatch.setenv("ASR_TOPK_PATH", path)
    rng = np.random.RandomState(5)
    centers = _db(40, 31)
    D = np.repeat(centers, 500, axis=0) + rng.normal(0, 1e-5, (20000, 32)).astype(np.float32)
    D = np.concatenate([D, _db(30000, 32)]).astype(np.float32)
    Q = centers + rng.normal(0, 1e-3, centers.shape).astype(np.float32)
    s_ref, i_ref = clib.topk(Q, D, 25)
    s, i = EmbeddingDB(D).topk(Q, 25)
    assert (i == i_ref).all() and (s == s_ref).all()


@pytest.mark.parametrize("nq,k", [(1, 25), (3, 27), (8, 1), (2, 25)])
def test_many_slices_select_merge_ties_and_degenerate_rows(nq, k):
    """Few queries over a DB large enough for one list per resident CTA (444 slices): the radix-select
    merge.  Duplicated rows (ties resolve by index), zero rows (NaN -> -inf, never selected), an index
    base, and a DB of identical rows (every score equal: pure index order)."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    n = 400000
    D = _db(n, 31)
    D[123456] = D[7] = D[399999] = D[250000]      # exact duplicates across slices
    D[1000:1100] = 0
    Q = np.concatenate([D[250000:250001], _db(nq, 32, unit=False)])[:nq]
    s_ref, i_ref = clib.topk(Q, D, k)
    db = EmbeddingDB(D, idx_base=1 << 33)
    s, i = db.topk(Q, k)
    assert (i - (1 << 33) == i_ref).all() and (s.view(np.uint32) == s_ref.view(np.uint32)).all()
    if k >= 4:
        assert list(i[0, :4] - (1 << 33)) == [7, 123456, 250000, 399999]
    E = np.tile(_db(1, 33), (n, 1))
    s2, i2 = EmbeddingDB(E).topk(Q[:1], k)
    assert list(i2[0]) == list(range(k)) and (s2[0] == s2[0, 0]).all()


def test_config4_full_size_default_dispatch_bit_exact():
    """BASELINE config 4 at its full size through the path asr_topk picks BY ITSELF (nothing forced): 10 000 queries x
    10^6 rows, k = 25 -> the tf32 pre-filter with many query tiles x several DB slices + the slice merge.  A 64-query
    sample of the result (indices and scores) must equal the pinned-order C oracle, and the piece vote on the full
    result must find every recording's piece."""
    import torch
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB, vote_device
    n_db, n_rec, win, k = 1000000, 100, 100, 25
    g = torch.Generator(device="cuda").manual_seed(1)
    D = torch.randn((n_db, 32), generator=g, device="cuda")
    D = D / D.norm(dim=1, keepdim=True)
    ids = (torch.arange(n_db, device="cuda") // 100).to(torch.int32)
    g2 = torch.Generator(device="cuda").manual_seed(2)
    true_piece = torch.randint(0, n_db // 100, (n_rec,), generator=g2, device="cuda")
    rows = (true_piece[:, None] * 100 + torch.randint(0, 100, (n_rec, win), generator=g2, device="cuda")).view(-1)
    Q = (D[rows] + 0.12 * torch.randn((n_rec * win, 32), generator=g2, device="cuda")).contiguous()
    assert Q.shape[0] * n_db >= 1.6e8 and Q.shape[0] > 24          # the dispatch rule sends this to the pre-filter
    db = EmbeddingDB(D)
    s, i = db.topk_device(Q, k)
    sample = np.linspace(0, Q.shape[0] - 1, 64).astype(int)
    s_ref, i_ref = clib.topk(Q[sample].cpu().numpy(), D.cpu().numpy(), k)
    assert (i[sample].cpu().numpy() == i_ref).all() and (s[sample].cpu().numpy() == s_ref).all()
    pid, cnt = vote_device(i.view(n_rec, -1), ids, 5)
    assert (pid[:, 0].long() == true_piece).all()
    # the same DB donated (rows normalised in place, no second copy) and with a small query workspace (chunked calls)
    db2 = EmbeddingDB(D.clone(), normalise_in_place=True, max_queries=3000)
    s2, i2 = db2.topk_device(Q, k)
    assert torch.equal(i2, i) and torch.equal(s2, s)
    with pytest.raises(Exception):
        db2.topk_device(Q[:4], k, normalise=False)                   # the raw rows are gone


def test_db_without_cosine_copy_matches():
    """ASR_DB_NO_COSINE_COPY: no normalised rows are kept, cosine queries normalise in-kernel -- same bits."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    D, Q = _db(50000, 41, unit=False), _db(19, 42, unit=False)
    D[4000] = D[17]
    s_ref, i_ref = clib.topk(Q, D, 25)
    s, i = EmbeddingDB(D, cosine_copy=False).topk(Q, 25)
    assert (i == i_ref).all() and (s == s_ref).all()
    s, i = EmbeddingDB(D, cosine_copy=False).topk(Q[:1], 25)
    assert (i == i_ref[:1]).all() and (s == s_ref[:1]).all()
