"""A/B helper for the tools in this directory: ASR_SO=<path> makes the binding load that shared library
instead of the in-tree build (import this before anything from the package)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("ASR_SO"):
    import audio_sheet_retrieval_b200.build as _b
    _so = os.path.abspath(os.environ["ASR_SO"])
    _b.SO = _so
    _b.build = lambda force=False, verbose=False: _so
