"""ctypes binding of libasr_b200.so (the C ABI declared in include/asr_b200.h).

The library is built in-tree by `audio_sheet_retrieval_b200.build`.  There is no CPU
fallback: if the shared library is missing and cannot be built, importing this module
raises, and every compute entry point fails without a CUDA device.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

from . import build as _build

ABI_VERSION = 2
N_LAYERS = 9
DIM = 32
MAX_K = 128
IN_F32, IN_U8 = 0, 1
PREP_NONE, PREP_SCALE, PREP_SCALE_HALF = 0, 1, 2
PATH_TCGEN05, PATH_FP32 = 0, 1
CCA_NSUMS = 3136
CCA_COUNT_ON_DEVICE = -1
DB_NORMALISE_IN_PLACE, DB_NO_COSINE_COPY = 1, 2


class AsrError(RuntimeError):
    pass


class EncoderDesc(ctypes.Structure):
    _fields_ = [
        ("in_h", c_int), ("in_w", c_int), ("prepare", c_int), ("flip_filters", c_int),
        ("channels", c_int * N_LAYERS),
        ("W", POINTER(c_float) * N_LAYERS),
        ("beta", POINTER(c_float) * N_LAYERS),
        ("gamma", POINTER(c_float) * N_LAYERS),
        ("mean", POINTER(c_float) * N_LAYERS),
        ("inv_std", POINTER(c_float) * N_LAYERS),
        ("cca_mean", POINTER(c_float)),
        ("cca_proj", POINTER(c_float)),
    ]


def _load():
    # build() rebuilds only when a source or the header is newer than the .so (mtime check), so a stale
    # library is never loaded with newer argtypes; without nvcc (a deployment box) the shipped .so is used.
    try:
        so = _build.build(force=bool(os.environ.get("ASR_REBUILD")))
    except (OSError, RuntimeError, subprocess.CalledProcessError):
        if not os.path.exists(_build.SO):
            raise
        so = _build.SO
    lib = ctypes.CDLL(so)
    lib.asr_last_error.restype = c_char_p
    lib.asr_abi_version.restype = c_int
    if lib.asr_abi_version() != ABI_VERSION:
        raise AsrError("libasr_b200.so has ABI version %d, this binding expects %d: rebuild with "
                       "`python -m audio_sheet_retrieval_b200.build --force`" % (lib.asr_abi_version(), ABI_VERSION))
    lib.asr_launch_count.restype = c_int64
    lib.asr_encoder_create.argtypes = [POINTER(c_void_p), POINTER(EncoderDesc), c_int]
    lib.asr_encoder_destroy.argtypes = [c_void_p]
    lib.asr_encoder_set_cca.argtypes = [c_void_p, POINTER(c_float), POINTER(c_float)]
    lib.asr_encoder_embed.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p]
    lib.asr_encoder_embed_host.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int]
    lib.asr_encoder_debug_activation.argtypes = [c_void_p, c_int, c_int, c_int64, c_void_p,
                                                 POINTER(c_int), POINTER(c_int), POINTER(c_int)]
    lib.asr_cosine_distances.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.asr_cosine_distances.restype = c_int
    lib.asr_dtw.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.asr_dtw.restype = c_int
    lib.asr_extract_windows.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p]
    lib.asr_extract_windows.restype = c_int
    lib.asr_cca_layer_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_void_p,
                                           c_void_p, c_void_p, c_void_p]
    lib.asr_cca_layer_backward.restype = c_int
    lib.asr_spectrogram_num_frames.argtypes = [c_int64, c_int, c_double]
    lib.asr_spectrogram_num_frames.restype = c_int
    lib.asr_log_spectrogram.argtypes = [c_void_p, c_int64, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                        c_void_p]
    lib.asr_log_spectrogram.restype = c_int
    lib.asr_encoder_set_timing.argtypes = [c_void_p, c_int]
    lib.asr_encoder_get_timing.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_double),
                                           POINTER(c_int64)]
    lib.asr_encoder_set_fusion.argtypes = [c_void_p, c_int]
    lib.asr_encoder_set_fusion.restype = c_int
    lib.asr_encoder_get_fusion.argtypes = [c_void_p]
    lib.asr_encoder_get_fusion.restype = c_int
    lib.asr_encoder_flops_per_sample.argtypes = [c_void_p]
    lib.asr_encoder_flops_per_sample.restype = c_double
    lib.asr_db_create.argtypes = [POINTER(c_void_p), c_void_p, c_int64, c_int64]
    lib.asr_db_create_ex.argtypes = [POINTER(c_void_p), c_void_p, c_int64, c_int64, c_int, c_int64]
    lib.asr_db_destroy.argtypes = [c_void_p]
    lib.asr_topk_merge_gathered.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.asr_rank_target_merge.argtypes = [c_void_p, c_int64, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p]
    lib.asr_cca_accumulate_counted.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.asr_topk.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.asr_debug_tc_scores.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.asr_debug_tc_scores.restype = c_int
    lib.asr_topk_merge.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.asr_rank_of_target.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p]
    lib.asr_vote.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.asr_cca_accumulate.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.asr_cca_solve.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_double, c_double, c_double, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.asr_contrastive_loss.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_float, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]
    lib.asr_contrastive_loss.restype = c_int
    for name in ("asr_encoder_create", "asr_encoder_destroy", "asr_encoder_set_cca", "asr_encoder_embed",
                 "asr_encoder_embed_host", "asr_encoder_debug_activation", "asr_encoder_set_timing",
                 "asr_encoder_get_timing", "asr_db_create", "asr_db_destroy",
                 "asr_topk", "asr_topk_merge", "asr_rank_of_target", "asr_vote", "asr_cca_accumulate",
                 "asr_cca_solve", "asr_db_create_ex", "asr_topk_merge_gathered", "asr_rank_target_merge",
                 "asr_cca_accumulate_counted"):
        getattr(lib, name).restype = c_int
    return lib


lib = _load()
SO_PATH = _build.SO


def check(rc):
    if rc != 0:
        raise AsrError("libasr_b200 error %d: %s" % (rc, lib.asr_last_error().decode("utf-8", "replace")))


def launch_count():
    return int(lib.asr_launch_count())


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream)."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def dptr(t):
    """Device (or host) pointer of a torch tensor / numpy array as c_void_p (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if hasattr(t, "data_ptr"):
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)
