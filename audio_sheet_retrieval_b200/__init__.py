"""B200-native retrieval hot path for CPJKU/audio_sheet_retrieval.

Host side mirrors the reference's Python interface (RetrievalWrapper, eval_retrieval, CCA,
AudioSheetServer, model modules); all arithmetic runs in hand-written sm_100a CUDA behind the
C ABI of include/asr_b200.h (libasr_b200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"
