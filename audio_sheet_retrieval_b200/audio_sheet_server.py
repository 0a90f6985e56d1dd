"""AudioSheetServer: piece identification by brute-force cosine search + per-piece vote.

Mirror of the retrieval part of audio_sheet_retrieval/audio_sheet_server.py: `_retrieve_*`
(:530-563), `detect_score` (:213-253), `detect_performance` (:255-300), DB containers and their
pickle files (:496-522).  The DB lives in HBM (EmbeddingDB); each `detect_*` call runs ONE fused
top-k launch for its 100 windows (the reference loops `cdist` + full `argsort` per window) and one
vote kernel.  The streaming demo GUI (`run`, :83-211) and MSMD-based DB construction (:309-494) are
out of scope; `set_sheet_db` / `set_audio_db` take codes computed with `embed_network`.
"""
from __future__ import print_function

import pickle

import numpy as np
import torch

from .retrieval import EmbeddingDB, _as_codes
from .retrieval_wrapper import RetrievalWrapper


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
        """ detect piece from audio """
        n_samples = 100
        start_indices = np.linspace(start=0, stop=spectrogram.shape[1] - self.spec_shape[1], num=n_samples)
        start_indices = start_indices.astype(int)
        spec_excerpts = np.zeros((len(start_indices), 1, self.spec_shape[0], self.spec_shape[1]), dtype=np.float32)
        for i, idx in enumerate(start_indices):
            spec_excerpts[i, 0] = spectrogram[:, idx:idx + self.spec_shape[1]]
        spec_codes = self.embed_network.compute_view_2(spec_excerpts)
        return self._vote(self._sheet_db, spec_codes, self.id_to_piece, top_k, n_candidates, verbose)

    def detect_performance(self, sheet, top_k=1, n_candidates=1, verbose=False):
        """ detect performance from sheet """
        n_samples = 100
        start_indices = np.linspace(start=0, stop=sheet.shape[1] - self.sheet_shape[1], num=n_samples)
        start_indices = start_indices.astype(int)
        r0 = sheet.shape[0] // 2 - self.sheet_shape[0] // 2
        r1 = r0 + self.sheet_shape[0]
        sheet_snippets = np.zeros((len(start_indices), 1, self.sheet_shape[0], self.sheet_shape[1]), dtype=np.float32)
        for i, idx in enumerate(start_indices):
            sheet_snippets[i, 0] = sheet[r0:r1, idx:idx + self.sheet_shape[1]]
        sheet_codes = self.embed_network.compute_view_1(sheet_snippets)
        return self._vote(self._audio_db, sheet_codes, self.id_to_perform, top_k, n_candidates, verbose)

    # -- batched identification over many recordings (config "piece identification") -----------
    def identify_from_codes(self, query_codes, n_recordings, top_k=1, n_candidates=25, direction="A2S"):
        """query_codes: (n_recordings * windows, 32).  -> (piece ids (n_rec, top_k), counts) NumPy."""
        db = self._sheet_db if direction == "A2S" else self._audio_db
        q = _as_codes(query_codes, db.device)
        _, idx = db.topk_device(q, n_candidates)
        ids, counts = db.vote_device(idx.view(n_recordings, -1), top_k)
        return ids.cpu().numpy(), counts.cpu().numpy()
