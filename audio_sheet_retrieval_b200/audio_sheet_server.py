"""AudioSheetServer: piece identification by brute-force cosine search + per-piece vote.

Mirror of the retrieval part of audio_sheet_retrieval/audio_sheet_server.py: `_retrieve_*`
(:530-563), `detect_score` (:213-253), `detect_performance` (:255-300), DB containers and their
pickle files (:496-522).  The DB lives in HBM (EmbeddingDB); each `detect_*` call runs ONE fused
top-k launch for its 100 windows (the reference loops `cdist` + full `argsort` per window) and one
vote kernel.  The streaming demo GUI (`run`, :83-211) and MSMD-based DB construction (:309-494) are
out of scope; `set_sheet_db` / `set_audio_db` take codes computed with `embed_network`.
"""
from __future__ import print_function

import json
import os
import pickle

import numpy as np
import torch

from . import _lib
from .retrieval import EmbeddingDB, _as_codes
from .retrieval_wrapper import RetrievalWrapper


def extract_windows_device(src, starts, r0, win_h, win_w):
    """src: 2-D CUDA tensor (uint8 or float32) holding one unrolled sheet / spectrogram.
    -> (n, 1, win_h, win_w) CUDA tensor of the same dtype (asr_extract_windows)."""
    assert src.is_cuda and src.dim() == 2 and src.is_contiguous() and src.dtype in (torch.uint8, torch.float32)
    starts = np.asarray(starts, dtype=np.int32)
    if len(starts) and (starts.min() < 0 or starts.max() > src.shape[1] - win_w):
        raise ValueError("window start out of range")
    st = torch.as_tensor(starts).to(src.device)
    out = torch.empty((len(starts), 1, win_h, win_w), dtype=src.dtype, device=src.device)
    _lib.check(_lib.lib.asr_extract_windows(_lib.dptr(src), _lib.IN_U8 if src.dtype == torch.uint8 else _lib.IN_F32,
                                            int(src.shape[0]), int(src.shape[1]), _lib.dptr(st), len(starts), int(r0),
                                            int(win_h), int(win_w), _lib.dptr(out), _lib.stream_ptr()))
    return out


class AudioSheetServer(object):
    """ Audio-to-sheet / sheet-to-audio retrieval server (retrieval + voting part) """

    def __init__(self, spec_shape=(92, 42), sheet_shape=(160, 200)):
        self.spec_shape = tuple(spec_shape)
        self.sheet_shape = tuple(sheet_shape)
        self.embed_network = None
        self.sheet_snippet_codes = self.sheet_snippet_ids = self.id_to_piece = self.sheet_snippets = None
        self.perform_excerpt_codes = self.perform_excerpt_ids = self.id_to_perform = self.perform_excerpts = None
        self._sheet_db = self._audio_db = None

    # -- embedding network (:302-307) ---------------------------------------------------------
    def initialize_embedding_network(self, model, param_file):
        """ load cross modality retrieval model """
        self.embed_network = RetrievalWrapper(model, param_file, prepare_view_1=model.prepare, prepare_view_2=None)
        self.snippet_shape = self.embed_network.shape_view1[1:]
        self.excerpt_shape = self.embed_network.shape_view2[1:]

    # -- spectrogram front-end (the madmom chain of the tutorial, cell 28, and of `_mic_spec_gen`, :44-60) ----------
    def compute_spectrogram(self, samples, sample_rate=22050):
        """mono PCM samples (float in [-1, 1) or int16) -> (92, n_frames) float32 log-frequency spectrogram, computed
        on the device (frame 2048, 20 fps, LogarithmicFilterbank(16 bands / octave, 30 Hz - 6 kHz), log10(1 + x))."""
        from .utils.spectrogram import LogSpectrogramProcessor
        key = int(sample_rate)
        if getattr(self, "_spec_proc", None) is None or self._spec_proc.sample_rate != key:
            self._spec_proc = LogSpectrogramProcessor(sample_rate=key)
        return self._spec_proc.process(samples)

    # -- DB containers ------------------------------------------------------------------------
    def set_sheet_db(self, codes, ids, id_to_piece, snippets=None):
        self.sheet_snippet_codes = np.ascontiguousarray(codes, np.float32)
        self.sheet_snippet_ids = np.asarray(ids)
        self.id_to_piece = id_to_piece
        self.sheet_snippets = snippets
        self._sheet_db = EmbeddingDB(self.sheet_snippet_codes, ids=self.sheet_snippet_ids)

    def set_audio_db(self, codes, ids, id_to_perform, excerpts=None):
        self.perform_excerpt_codes = np.ascontiguousarray(codes, np.float32)
        self.perform_excerpt_ids = np.asarray(ids)
        self.id_to_perform = id_to_perform
        self.perform_excerpts = excerpts
        self._audio_db = EmbeddingDB(self.perform_excerpt_codes, ids=self.perform_excerpt_ids)

    # -- DB construction from raw material (:403-494) -------------------------------------------
    def _embed_windows(self, view, src_2d, starts, r0, win_h, win_w):
        """One H2D copy of the whole piece, windows cut and embedded on the device."""
        net = self.embed_network.net
        mode = self.embed_network.prepare_view_1.asr_prepare_mode if view == 1 else _lib.PREP_NONE
        enc = net.encoder(view, mode)
        dev = torch.device("cuda", torch.cuda.current_device())
        src = torch.as_tensor(np.ascontiguousarray(src_2d))
        src = src.to(dev) if src.dtype == torch.uint8 else src.to(dev, torch.float32)
        wins = extract_windows_device(src.contiguous(), starts, r0, win_h, win_w)
        codes = torch.empty((wins.shape[0], 32), dtype=torch.float32, device=dev)
        for s0 in range(0, wins.shape[0], enc.max_batch):
            enc.embed_device(wins[s0:s0 + enc.max_batch], codes=codes[s0:s0 + enc.max_batch])
        return codes

    def initialize_audio_db_from_specs(self, pieces, spectrograms, keep_snippets=False):
        """ init audio data base: excerpts at hop spec_context // 4 of every spectrogram (:403-445) """
        print("Initializing audio db ...")
        id_to_perform, codes, ids = dict(), [], []
        for piece_idx, piece in enumerate(pieces):
            id_to_perform[piece_idx] = piece
            spectrogram = spectrograms[piece_idx]
            indices = np.arange(0, spectrogram.shape[1] - self.spec_shape[1], self.spec_shape[1] // 4)
            codes.append(self._embed_windows(2, spectrogram, indices, 0, self.spec_shape[0], self.spec_shape[1]))
            ids.append(np.ones(len(indices), dtype=np.int64) * piece_idx)
        codes = torch.cat(codes).cpu().numpy() if codes else np.zeros((0, 32), np.float32)
        self.set_audio_db(codes, np.concatenate(ids), id_to_perform)
        print("%s audio excerpts of %d pieces collected" % (codes.shape[0], len(pieces)))

    def initialize_sheet_db_from_imges(self, pieces, scores, keep_snippets=False):
        """ init sheet music data base: snippets at hop sheet_context // 4 from the vertical centre of every
        unrolled score (:447-494; the reference spells it `imges`) """
        print("Initializing sheet music db ...")
        id_to_piece, codes, ids = dict(), [], []
        for piece_idx, piece in enumerate(pieces):
            id_to_piece[piece_idx] = piece
            piece_image = scores[piece_idx]
            indices = np.arange(0, piece_image.shape[1] - self.sheet_shape[1], self.sheet_shape[1] // 4)
            r0 = piece_image.shape[0] // 2 - self.sheet_shape[0] // 2
            codes.append(self._embed_windows(1, piece_image, indices, r0, self.sheet_shape[0], self.sheet_shape[1]))
            ids.append(np.ones(len(indices), dtype=np.int64) * piece_idx)
        codes = torch.cat(codes).cpu().numpy() if codes else np.zeros((0, 32), np.float32)
        self.set_sheet_db(codes, np.concatenate(ids), id_to_piece)
        print("%s sheet snippet codes of %d pieces collected" % (codes.shape[0], len(pieces)))

    # -- memory-mappable DB directories (sharded loading for multi-GPU serving) -------------------
    @staticmethod
    def save_db_dir(path, codes, ids, id_to_name):
        """codes.npy (n,32) float32 | ids.npy (n,) int32 | meta.json; both arrays can be np.load(mmap_mode='r')-ed."""
        os.makedirs(path, exist_ok=True)
        np.save(os.path.join(path, "codes.npy"), np.ascontiguousarray(codes, np.float32))
        np.save(os.path.join(path, "ids.npy"), np.ascontiguousarray(ids, np.int32))
        with open(os.path.join(path, "meta.json"), "w") as fp:
            json.dump({"n": int(len(ids)), "dim": int(codes.shape[1]),
                       "id_to_name": {str(k): v for k, v in id_to_name.items()}}, fp)

    @staticmethod
    def load_db_dir(path, rank=0, world=1):
        """-> (codes rows [lo,hi) of this rank, ALL ids, id_to_name, lo).  Only the shard's rows are read."""
        meta = json.load(open(os.path.join(path, "meta.json")))
        codes = np.load(os.path.join(path, "codes.npy"), mmap_mode="r")
        ids = np.load(os.path.join(path, "ids.npy"))
        lo, hi = meta["n"] * rank // world, meta["n"] * (rank + 1) // world
        return np.ascontiguousarray(codes[lo:hi]), ids, {int(k): v for k, v in meta["id_to_name"].items()}, lo

    def load_sheet_db_file(self, sheet_db_path):
        """ load sheet codes (:496-501) """
        with open(sheet_db_path, 'rb') as fp:
            data = pickle.load(fp, encoding="latin1")
        self.set_sheet_db(*data)

    def save_sheet_db_file(self, sheet_db_path):
        """ preserve sheet codes (:503-508) """
        with open(sheet_db_path, 'wb') as fp:
            pickle.dump([self.sheet_snippet_codes, self.sheet_snippet_ids, self.id_to_piece, self.sheet_snippets],
                        fp, protocol=2)

    def load_audio_db_file(self, audio_db_path):
        """ load audio codes (:510-515) """
        with open(audio_db_path, 'rb') as fp:
            data = pickle.load(fp, encoding="latin1")
        self.set_audio_db(*data)

    def save_audio_db_file(self, audio_db_path):
        """ preserve audio codes (:517-522) """
        with open(audio_db_path, 'wb') as fp:
            pickle.dump([self.perform_excerpt_codes, self.perform_excerpt_ids, self.id_to_perform,
                         self.perform_excerpts], fp, protocol=2)

    # -- search (:530-563) --------------------------------------------------------------------
    def _retrieve_sheet_snippet_ids(self, spectrogram_code, n_candidates=1):
        """ retrieve k most similar sheet music snippets -> (piece ids, row indices) """
        _, idx = self._sheet_db.topk(spectrogram_code, n_candidates)
        sorted_idx = idx[0][idx[0] >= 0]
        return self.sheet_snippet_ids[sorted_idx], sorted_idx

    def _retrieve_perform_excerpt_ids(self, sheet_code, n_candidates=1):
        """ retrieve k most similar performance excerpts -> (performance ids, row indices) """
        _, idx = self._audio_db.topk(sheet_code, n_candidates)
        sorted_idx = idx[0][idx[0] >= 0]
        return self.perform_excerpt_ids[sorted_idx], sorted_idx

    # -- detection (:213-300) -----------------------------------------------------------------
    def _vote(self, db, codes, id_to_name, top_k, n_candidates, verbose):
        q = _as_codes(codes, db.device)
        _, idx = db.topk_device(q, n_candidates)
        ids, counts = db.vote_device(idx.view(1, -1), top_k)
        ids, counts = ids[0].cpu().numpy(), counts[0].cpu().numpy()
        keep = ids >= 0
        ids, counts = ids[keep], counts[keep]
        if verbose:
            print("\nRetrieval Ranking:")
            for pid, cnt in zip(ids, counts):
                print("pid: %03d (%03d): %s" % (pid, cnt, id_to_name[pid]))
            print("")
        ret_result = [id_to_name[pid] for pid in ids]
        ret_votes = np.asarray(counts, dtype=float) / np.sum(counts)
        return ret_result, ret_votes

    def detect_score(self, spectrogram, top_k=1, n_candidates=1, verbose=False):
        """ detect piece from audio (:213-253): 100 windows; the spectrogram goes to the device once and the
        windows are cut (asr_extract_windows) and embedded there, the codes never come back to the host """
        n_samples = 100
        start_indices = np.linspace(start=0, stop=spectrogram.shape[1] - self.spec_shape[1], num=n_samples)
        start_indices = start_indices.astype(int)
        spec_codes = self._embed_windows(2, np.asarray(spectrogram, dtype=np.float32), start_indices, 0,
                                         self.spec_shape[0], self.spec_shape[1])
        return self._vote(self._sheet_db, spec_codes, self.id_to_piece, top_k, n_candidates, verbose)

    def detect_performance(self, sheet, top_k=1, n_candidates=1, verbose=False):
        """ detect performance from sheet (:255-300): 100 snippets from the vertical centre of the unrolled sheet,
        cut and embedded on the device like detect_score """
        n_samples = 100
        start_indices = np.linspace(start=0, stop=sheet.shape[1] - self.sheet_shape[1], num=n_samples)
        start_indices = start_indices.astype(int)
        r0 = sheet.shape[0] // 2 - self.sheet_shape[0] // 2
        sheet = np.asarray(sheet)
        if sheet.dtype != np.uint8:
            sheet = sheet.astype(np.float32)      # the reference copies the windows into a float32 array
        sheet_codes = self._embed_windows(1, sheet, start_indices, r0, self.sheet_shape[0], self.sheet_shape[1])
        return self._vote(self._audio_db, sheet_codes, self.id_to_perform, top_k, n_candidates, verbose)

    # -- streaming identification (the vote of `run`, :118-138, without the GUI) ----------------------
    def reset_stream(self):
        self._stream_ids = np.zeros(0, dtype=np.int64)

    def process_frame(self, running_spec, top_k=5, n_candidates=25, running_frames=100):
        """One step of the streaming loop: embed the current (92, 42) window, retrieve n_candidates sheet
        snippets, vote over the ids of the last `running_frames` frames.  -> (names, probabilities)."""
        if not hasattr(self, "_stream_ids"):
            self.reset_stream()
        spec_code = self.embed_network.compute_view_2(running_spec[np.newaxis, np.newaxis, :, :])
        piece_ids, _ = self._retrieve_sheet_snippet_ids(spec_code, n_candidates=n_candidates)
        self._stream_ids = np.concatenate((self._stream_ids, piece_ids))
        first_idx = running_frames * n_candidates if running_frames is not None else 0
        if running_frames is not None and self._stream_ids.shape[0] > first_idx:
            self._stream_ids = self._stream_ids[-first_idx:]
        dev = self._sheet_db.device
        ids_dev = torch.as_tensor(self._stream_ids.astype(np.int32)).to(dev)
        rows = torch.arange(len(self._stream_ids), dtype=torch.int64, device=dev).view(1, -1)
        from .retrieval import vote_device
        out_ids, out_cnt = vote_device(rows, ids_dev, top_k)           # row -> id table = the window itself
        out_ids, out_cnt = out_ids[0].cpu().numpy(), out_cnt[0].cpu().numpy()
        keep = out_ids >= 0
        probs = out_cnt[keep].astype(float) / len(self._stream_ids)
        return [self.id_to_piece[i] for i in out_ids[keep]], probs

    def run(self, spec, top_k=5, n_candidates=5, running_frames=None, gui=False, target_piece=None, verbose=True):
        """ run sheet retrieval service over a recorded spectrogram (:83-211 without the GUI and the microphone;
        the music detector of `_detect_music` is a separate network that is not part of this path, every frame
        counts as music).  One frame = one spectrogram column; from the first full window on, every frame is
        embedded, its n_candidates nearest sheet snippets vote.  Prints the reference's fps line and returns
        (names, probabilities, frames per second over the frames that did retrieval). """
        import sys
        import time
        if gui:
            raise NotImplementedError("the matplotlib GUI of the reference is not part of this package")
        if verbose:
            print("Running server ...")
        running_spec = np.zeros((self.spec_shape[0], self.spec_shape[1]), dtype=np.float32)
        self.reset_stream()
        frame_times = np.zeros(10)
        names, probs, t_retrieval, n_retrieval = [], np.zeros(0), 0.0, 0
        for i_frame in range(spec.shape[1]):
            start = time.time()
            running_spec = np.hstack((running_spec[:, 1::], spec[:, i_frame:i_frame + 1])).astype(np.float32)
            if i_frame >= running_spec.shape[1]:
                names, probs = self.process_frame(running_spec, top_k=top_k, n_candidates=n_candidates,
                                                  running_frames=running_frames)
                t_retrieval += time.time() - start
                n_retrieval += 1
            stop = time.time()
            frame_times[1:] = frame_times[0:-1]
            frame_times[0] = stop - start
            fps = 1.0 / np.mean(frame_times)
            if verbose:
                print("Server is running at %.2f fps." % fps, end='\r')
                sys.stdout.flush()
        if verbose:
            print("")
        return names, probs, (n_retrieval / t_retrieval if t_retrieval > 0 else 0.0)

    # -- batched identification over many recordings (config "piece identification") -----------
    def identify_from_codes(self, query_codes, n_recordings, top_k=1, n_candidates=25, direction="A2S"):
        """query_codes: (n_recordings * windows, 32).  -> (piece ids (n_rec, top_k), counts) NumPy."""
        db = self._sheet_db if direction == "A2S" else self._audio_db
        q = _as_codes(query_codes, db.device)
        _, idx = db.topk_device(q, n_candidates)
        ids, counts = db.vote_device(idx.view(n_recordings, -1), top_k)
        return ids.cpu().numpy(), counts.cpu().numpy()
