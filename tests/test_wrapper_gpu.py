"""The reference-facing interface end to end: RetrievalWrapper, run_eval, refine_cca, server."""
import os

import numpy as np
import pytest
import yaml

from oracle import cca as occa
from oracle import metrics
from oracle.encoders import OracleNet, load_param_list, synth_inputs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PKL = os.path.join(GOLDEN, "params_all_split_mutopia_full_aug.pkl")


def _cos(a, b):
    return (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)


def test_retrieval_wrapper_tutorial_contract():
    """Embedding Tutorial.ipynb cells 20-33: shapes, dtypes, attributes."""
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from audio_sheet_retrieval_b200.retrieval_wrapper import RetrievalWrapper
    w = RetrievalWrapper(model, PKL, prepare_view_1=model.prepare, prepare_view_2=None)
    assert w.code_dim == 32 and tuple(w.shape_view1) == (1, 80, 100) and tuple(w.shape_view2) == (1, 92, 42)
    assert model.INPUT_SHAPE_1[1:] == [160, 200] and model.INPUT_SHAPE_2[1:] == [92, 42]
    X1, X2 = synth_inputs(100, seed=8)
    c1, c2 = w.compute_view_1(X1), w.compute_view_2(X2)
    assert c1.shape == (100, 32) and c2.shape == (100, 32) and c1.dtype == np.float32
    onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL))
    assert _cos(c1, onet.compute_view_1(X1)).min() >= 0.999
    assert _cos(c2, onet.compute_view_2(X2)).min() >= 0.999
    # audio-to-audio tutorial: both prepare_* None, view-1 input already prepared
    w2 = RetrievalWrapper(model, PKL)
    c1b = w2.compute_view_1(model.prepare(X1))
    assert _cos(c1b, c1).min() > 0.99999
    # a custom host-side prepare callable goes through the generic batch loop
    w3 = RetrievalWrapper(model, PKL, prepare_view_1=lambda x: model.prepare(x))
    assert _cos(w3.compute_view_1(X1[:23]), c1[:23]).min() > 0.99999
    # two-input compiled functions (run_eval.py:91-95)
    r = w.compute_v2_latent(w.dummy_in_v1, X2[:5])
    assert (r == c2[:5]).all()


def test_run_eval_metrics_vs_oracle(tmp_path, capsys):
    """Config 1 (reduced n): R@k / MRR / MR within 0.5 % absolute of the oracle's eval on oracle codes."""
    from audio_sheet_retrieval_b200 import run_eval
    from audio_sheet_retrieval_b200.utils.mutopia_data import SyntheticPairPool
    n = 300
    res = run_eval.main(["--model", "mutopia_ccal_cont_rsz", "--data", "synthetic", "--n_test", str(n),
                         "--param_file", PKL])
    out = capsys.readouterr().out
    assert "Median Rank" in out and "MAP" in out
    pool = SyntheticPairPool(2000, seed=25)
    idx = np.linspace(0, 1999, n).astype(int)
    X1, X2 = pool[idx]
    onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL))
    mr, med, md, hr, mrr = metrics.eval_retrieval_ref(onet.compute_view_1(X1), onet.compute_view_2(X2))
    assert abs(res["map"] - mrr) <= 0.005
    assert abs(res["med_rank"] - med) <= max(1.0, 0.005 * n)
    for k in (1, 5, 10, 25):
        assert abs(res["recall_at_k"]["%d" % k] - 100.0 * hr[k] / n) <= 0.5 + 100.0 / n


def test_run_eval_yaml_schema(tmp_path):
    from audio_sheet_retrieval_b200 import run_eval
    import shutil
    pf = tmp_path / "params_all_split_mutopia_full_aug.pkl"
    shutil.copy(PKL, pf)
    run_eval.main(["--model", "mutopia_ccal_cont_rsz", "--data", "synthetic", "--n_test", "40", "--param_file", str(pf),
                   "--dump_results", "--V2_to_V1", "--max_dim", "16"])
    # synthetic pairs never write under the real data set's file name and are tagged in the file
    assert not (tmp_path / "eval_all_split_mutopia_full_aug_A2S.yaml").exists()
    res = yaml.safe_load(open(tmp_path / "eval_all_split_mutopia_full_aug_A2S_synthetic.yaml"))
    assert res.pop("data") == "synthetic"
    assert set(res) == {"map", "med_rank", "recall_at_k"} and set(res["recall_at_k"]) == {"1", "5", "10", "25"}
    with pytest.raises(RuntimeError, match="msmd"):
        run_eval.main(["--model", "mutopia_ccal_cont_rsz", "--data", "mutopia", "--n_test", "40", "--param_file", str(pf)])


def test_refine_cca_roundtrip(tmp_path):
    """refine_cca: pickle in -> pickle out, entries 90-93 replaced, fit equals the oracle's fit on
    the same latents, invariants hold on the library's own latents."""
    from audio_sheet_retrieval_b200 import network, refine_cca
    from audio_sheet_retrieval_b200.params import load_params
    out = tmp_path / "out" / "params.pkl"
    refine_cca.main(["--model", "mutopia_ccal_cont_rsz", "--data", "synthetic", "--n_train", "400", "--param_file", PKL,
                     "--out_file", str(out)])
    old, new = load_params(PKL), load_params(str(out))
    assert len(new) == 97
    for i in range(97):
        same = (old[i] == new[i]).all()
        assert same == (i not in (90, 91, 92, 93)), i
    # oracle fit on the library's latents
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from audio_sheet_retrieval_b200.utils.mutopia_data import SyntheticPairPool
    layers = model.build_model(False)
    network.set_all_param_values(layers, old)
    net = layers[0].net
    X1, X2 = SyntheticPairPool(25000, seed=23)[0:400]
    l1 = net.encoder(1, model.prepare.asr_prepare_mode).embed_host(X1, want="latents")
    l2 = net.encoder(2, 0).embed_host(X2, want="latents")
    o = occa.CCA()
    sig = o.fit(l1, l2)
    U, V = new[90].astype(np.float64), new[91].astype(np.float64)
    np.testing.assert_allclose(new[92], o.m1, atol=1e-6)
    np.testing.assert_allclose(np.abs(np.diag(U.T @ o.S12 @ V)), sig, atol=5e-4)
    np.testing.assert_allclose(U.T @ o.S11 @ U, np.eye(32), atol=5e-3)


def test_detect_score_end_to_end():
    """detect_score on a synthetic recording whose windows were put into the DB: the true piece wins."""
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    rng = np.random.RandomState(0)
    srv = AudioSheetServer()
    srv.initialize_embedding_network(model, PKL)
    specs = [np.abs(rng.normal(0, 0.3, (92, 300))).astype(np.float32) * (rng.rand(92, 1) > 0.7) for _ in range(4)]
    codes, ids = [], []
    for pid, sp in enumerate(specs):
        ex = np.stack([sp[None, :, s:s + 42] for s in range(0, 300 - 42, 5)])
        codes.append(srv.embed_network.compute_view_2(ex))
        ids += [pid] * len(ex)
    # audio -> audio DB: identify recording 2 from its own spectrogram
    srv.set_sheet_db(np.concatenate(codes), np.array(ids), dict((i, "rec%d" % i) for i in range(4)))
    names, votes = srv.detect_score(specs[2], top_k=3, n_candidates=5)
    assert names[0] == "rec2" and votes[0] > 0.5 and abs(votes.sum() - 1.0) < 1e-12


def test_detect_on_device_windows_equals_host_window_loop():
    """detect_score / detect_performance cut their 100 windows on the device (one upload, asr_extract_windows) and
    keep the codes there; the reference's loop (audio_sheet_server.py:216-223, 260-271: host slices -> compute_view_x
    -> per-window retrieval -> vote) must give the same names and vote shares."""
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    rng = np.random.RandomState(7)
    srv = AudioSheetServer()
    srv.initialize_embedding_network(model, PKL)
    specs = [np.abs(rng.normal(0, 0.3, (92, 420))).astype(np.float32) * (rng.rand(92, 1) > 0.6) for _ in range(5)]
    sheets = [(rng.rand(180, 1400) > 0.8).astype(np.uint8) * 255 for _ in range(5)]
    srv.initialize_audio_db_from_specs(["p%d" % i for i in range(5)], specs)
    a_codes, a_ids = srv.perform_excerpt_codes, srv.perform_excerpt_ids
    srv.initialize_sheet_db_from_imges(["p%d" % i for i in range(5)], sheets)
    srv.set_sheet_db(a_codes, a_ids, srv.id_to_perform)          # audio windows as the DB of detect_score
    sh_codes = srv.embed_network.compute_view_1(np.stack(
        [sheets[i][None, 10:170, s:s + 200] for i in range(5) for s in range(0, 1200, 50)]).astype(np.float32))
    srv.set_audio_db(sh_codes, np.repeat(np.arange(5), 24), srv.id_to_perform)   # sheet windows as the DB of detect_performance
    # host-window restatement of the two reference loops
    for which, src, shape in (("score", specs[3], (92, 42)), ("perf", sheets[1].astype(np.float32), (160, 200))):
        starts = np.linspace(0, src.shape[1] - shape[1], 100).astype(int)
        r0 = 0 if which == "score" else src.shape[0] // 2 - 80
        wins = np.stack([src[None, r0:r0 + shape[0], s:s + shape[1]] for s in starts]).astype(np.float32)
        if which == "score":
            codes = srv.embed_network.compute_view_2(wins)
            ref = srv._vote(srv._sheet_db, codes, srv.id_to_piece, 3, 5, False)
            got = srv.detect_score(src, top_k=3, n_candidates=5)
        else:
            codes = srv.embed_network.compute_view_1(wins)
            ref = srv._vote(srv._audio_db, codes, srv.id_to_perform, 3, 5, False)
            got = srv.detect_performance(src, top_k=3, n_candidates=5)
        assert got[0] == ref[0] and np.array_equal(got[1], ref[1])
    assert srv.detect_score(specs[3], 1, 5)[0][0] == "p3" and srv.detect_performance(sheets[1], 1, 5)[0][0] == "p1"


def test_headless_run_equals_process_frame_sequence():
    """AudioSheetServer.run over a recorded spectrogram (the reference's loop without GUI, microphone and music
    detector) ends with the ranking of the frame-by-frame calls and reports a frame rate."""
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    rng = np.random.RandomState(8)
    srv = AudioSheetServer()
    srv.initialize_embedding_network(model, PKL)
    srv.set_sheet_db(rng.normal(size=(400, 32)).astype(np.float32), np.repeat(np.arange(8), 50),
                     dict((i, "p%d" % i) for i in range(8)))
    spec = np.abs(rng.normal(0, 0.3, (92, 60))).astype(np.float32)
    names, probs, fps = srv.run(spec, top_k=3, n_candidates=5, running_frames=10, verbose=False)
    srv.reset_stream()
    running = np.zeros((92, 42), np.float32)
    ref = None
    for f in range(60):
        running = np.hstack((running[:, 1:], spec[:, f:f + 1]))
        if f >= 42:
            ref = srv.process_frame(running, top_k=3, n_candidates=5, running_frames=10)
    assert names == ref[0] and np.array_equal(probs, ref[1]) and fps > 0


def test_db_from_raw_material_matches_reference_loop(tmp_path):
    """initialize_*_db_from_* (audio_sheet_server.py:403-494): same window grid, same codes, same ids."""
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer, extract_windows_device
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    import torch
    rng = np.random.RandomState(3)
    srv = AudioSheetServer()
    srv.initialize_embedding_network(model, PKL)
    specs = [np.abs(rng.normal(0, 0.3, (92, T))).astype(np.float32) for T in (300, 173, 95)]
    scores = [rng.randint(0, 256, (H, Wd)).astype(np.uint8) for H, Wd in ((180, 900), (160, 451), (200, 260))]
    # windows are cut bit-exactly
    w = extract_windows_device(torch.as_tensor(scores[0]).cuda(), [0, 50, 700], 10, 160, 200).cpu().numpy()
    assert (w[1, 0] == scores[0][10:170, 50:250]).all() and (w[2, 0] == scores[0][10:170, 700:900]).all()
    srv.initialize_audio_db_from_specs(["a", "b", "c"], specs)
    srv.initialize_sheet_db_from_imges(["a", "b", "c"], scores)
    onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL))
    ref_codes, ref_ids = [], []
    for pi, sp in enumerate(specs):
        idx = np.arange(0, sp.shape[1] - 42, 42 // 4)
        ref_codes.append(onet.compute_view_2(np.stack([sp[None, :, i:i + 42] for i in idx])))
        ref_ids += [pi] * len(idx)
    ref_codes = np.concatenate(ref_codes)
    assert (srv.perform_excerpt_ids == np.array(ref_ids)).all()
    assert _cos(srv.perform_excerpt_codes, ref_codes).min() >= 0.999
    ref_codes, ref_ids = [], []
    for pi, im in enumerate(scores):
        idx = np.arange(0, im.shape[1] - 200, 200 // 4)
        r0 = im.shape[0] // 2 - 80
        ref_codes.append(onet.compute_view_1(np.stack([im[None, r0:r0 + 160, i:i + 200] for i in idx]).astype(np.float32)))
        ref_ids += [pi] * len(idx)
    assert (srv.sheet_snippet_ids == np.array(ref_ids)).all()
    assert _cos(srv.sheet_snippet_codes, np.concatenate(ref_codes)).min() >= 0.999
    # DB directory round trip, sharded
    AudioSheetServer.save_db_dir(str(tmp_path / "db"), srv.sheet_snippet_codes, srv.sheet_snippet_ids, srv.id_to_piece)
    parts = [AudioSheetServer.load_db_dir(str(tmp_path / "db"), r, 3) for r in range(3)]
    assert (np.concatenate([p[0] for p in parts]) == srv.sheet_snippet_codes).all()
    assert parts[1][3] == len(srv.sheet_snippet_ids) // 3 and parts[0][2] == {0: "a", 1: "b", 2: "c"}


def test_streaming_vote_matches_reference_loop():
    """The vote of AudioSheetServer.run (:118-138) restated with NumPy vs process_frame."""
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from oracle import search
    rng = np.random.RandomState(4)
    srv = AudioSheetServer()
    srv.initialize_embedding_network(model, PKL)
    db = rng.normal(size=(600, 32)).astype(np.float32)
    ids = np.repeat(np.arange(12), 50)
    srv.set_sheet_db(db, ids, dict((i, "p%d" % i) for i in range(12)))
    spec = np.abs(rng.normal(0, 0.3, (92, 42 + 12))).astype(np.float32)
    all_ids = np.zeros(0, dtype=np.int64)
    srv.reset_stream()
    for f in range(12):
        win = spec[:, f:f + 42]
        names, probs = srv.process_frame(win, top_k=4, n_candidates=5, running_frames=6)
        code = srv.embed_network.compute_view_2(win[None, None])
        _, idx = search.pinned_topk(code, db, 5)
        all_ids = np.concatenate((all_ids, ids[idx[0]]))[-30:]
        ref_ids, ref_votes, _ = search.vote_ref(all_ids, 4)
        assert names == ["p%d" % i for i in ref_ids]
        np.testing.assert_allclose(probs, ref_votes / float(len(all_ids)), atol=1e-12)


def test_config1_full_size_metrics_vs_oracle():
    """BASELINE config 1 at its full size: 2 000 synthetic pairs, tutorial weights, both directions.
    R@1/R@5/R@10/R@25 within 0.5 % absolute, MRR within 0.005, median rank within 0.5 % of N vs the
    oracle's eval_retrieval on oracle embeddings (north_star tolerance)."""
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from audio_sheet_retrieval_b200.retrieval_wrapper import RetrievalWrapper
    from audio_sheet_retrieval_b200.utils.train_dcca_pool import eval_retrieval
    from audio_sheet_retrieval_b200.utils.mutopia_data import SyntheticPairPool
    n = 2000
    X1, X2 = SyntheticPairPool(n, seed=23)[0:n]
    w = RetrievalWrapper(model, PKL, prepare_view_1=model.prepare, prepare_view_2=None)
    c1, c2 = w.compute_view_1(X1), w.compute_view_2(X2)
    onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL))
    r1, r2 = onet.compute_view_1(X1), onet.compute_view_2(X2)
    assert _cos(c1, r1).min() >= 0.999 and _cos(c2, r2).min() >= 0.999
    for a, b, ra, rb in ((c1, c2, r1, r2), (c2, c1, r2, r1)):            # S2A and A2S
        mr, med, md, hr, mrr = eval_retrieval(a, b)
        mr_o, med_o, md_o, hr_o, mrr_o = metrics.eval_retrieval_ref(ra, rb)
        for k in (1, 5, 10, 25):
            assert abs(100.0 * hr[k] / n - 100.0 * hr_o[k] / n) <= 0.5
        assert abs(mrr - mrr_o) <= 0.005
        assert abs(med - med_o) <= 0.005 * n
        assert abs(md - md_o) <= 5e-3          # mean diagonal cosine distance: bounded by the per-row embedding tolerance
