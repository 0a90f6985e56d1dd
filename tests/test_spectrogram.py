"""Spectrogram front-end: the oracle's restatement of the madmom chain of notebook cell 28, and the
CUDA kernel against it.  madmom is absent, so the pins are the notebook's own numbers (92 bands) and known-answer
signals (a pure tone lands in the band whose centre bin it is; silence gives log10(1) = 0)."""
import numpy as np
import pytest

from oracle import spectrogram as osp


def test_filterbank_has_the_notebooks_92_bands_and_unit_area():
    fb = osp.log_filterbank()
    assert fb.shape == (1024, 92)                            # "92 Frequency Bins", excerpt shape (100, 1, 92, 42)
    np.testing.assert_allclose(fb.sum(0), 1.0, rtol=1e-6)    # norm_filters=True
    centres = fb.argmax(0)
    assert (np.diff(centres) > 0).all() and centres[0] == 4 and centres[-1] == 527      # unique, ascending FFT bins
    f = osp.log_frequencies(16, 30, 6000)
    assert len(f) == 122 and f[0] >= 30 and f[-1] <= 6000
    np.testing.assert_allclose(f[1:] / f[:-1], 2.0 ** (1 / 16.0), rtol=1e-12)


def test_frame_count_and_known_answers():
    assert osp.num_frames(22050) == 20 and osp.num_frames(22051) == 21 and osp.num_frames(1) == 1
    assert osp.log_spectrogram(np.zeros(5000)).shape == (92, 5) and not osp.log_spectrogram(np.zeros(5000)).any()
    t = np.arange(22050 * 2) / 22050.0
    S = osp.log_spectrogram(0.5 * np.sin(2 * np.pi * 440.0 * t))
    fb = osp.log_filterbank()
    band_of_440 = int(np.argmin(np.abs(fb.argmax(0) - 440.0 / (22050 / 2048.0))))
    assert S.shape == (92, 40) and (S[:, 5:30].argmax(0) == band_of_440).all()


def test_product_filterbank_equals_oracle():
    from audio_sheet_retrieval_b200.utils.spectrogram import logarithmic_filterbank
    assert (logarithmic_filterbank() == osp.log_filterbank()).all()
    assert (logarithmic_filterbank(44100, 4096, 24, 30, 8000) == osp.log_filterbank(44100, 4096, 24, 30, 8000)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n_samples", [1, 2047, 22050, 100003])
def test_kernel_matches_oracle(n_samples):
    """fp32 FFT on the device vs float64 NumPy: log-magnitudes within 2e-4 absolute (values are 0 .. ~3)."""
    from audio_sheet_retrieval_b200.utils.spectrogram import LogSpectrogramProcessor
    rng = np.random.RandomState(n_samples)
    t = np.arange(n_samples) / 22050.0
    x = (0.3 * np.sin(2 * np.pi * 220.0 * t) + 0.2 * np.sin(2 * np.pi * 1760.0 * t * (1 + 0.1 * t)) +
         0.05 * rng.normal(size=n_samples)).astype(np.float32)
    proc = LogSpectrogramProcessor()
    S = proc.process(x)
    R = osp.log_spectrogram(x)
    assert S.shape == R.shape == (92, osp.num_frames(n_samples))
    assert np.abs(S - R).max() <= 2e-4
    xi = np.round(x * 32767).astype(np.int16)                  # int16 input is scaled like madmom scales its window
    np.testing.assert_allclose(proc.process(xi), osp.log_spectrogram(xi / 32768.0), atol=2e-4)


@pytest.mark.gpu
def test_audio_to_piece_identification_end_to_end():
    """PCM samples -> spectrogram on the device -> audio DB -> detect_score on an excerpt of one recording."""
    import os
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from audio_sheet_retrieval_b200.utils.spectrogram import LogSpectrogramProcessor
    pkl = os.path.join(os.path.dirname(__file__), "golden", "params_all_split_mutopia_full_aug.pkl")
    rng = np.random.RandomState(1)
    proc = LogSpectrogramProcessor()
    recs = []
    for p in range(3):                                        # three "pieces": different note sequences
        notes = 110.0 * 2.0 ** (rng.randint(0, 36, 40) / 12.0)
        t = np.arange(22050 // 2) / 22050.0
        recs.append(np.concatenate([0.4 * np.sin(2 * np.pi * f * t) * np.exp(-3 * t) for f in notes]).astype(np.float32))
    specs = [proc.process(r) for r in recs]
    assert all(s.shape == (92, 400) for s in specs)
    srv = AudioSheetServer()
    srv.initialize_embedding_network(model, pkl)
    srv.initialize_audio_db_from_specs(["a", "b", "c"], specs)
    srv.set_sheet_db(srv.perform_excerpt_codes, srv.perform_excerpt_ids, srv.id_to_perform)
    names, votes = srv.detect_score(proc.process(recs[1][22050 * 3:22050 * 15]), top_k=3, n_candidates=5)
    assert names[0] == "b" and votes[0] > 0.5
