"""Host-dependent paths (audio_sheet_retrieval/config/settings.py:4-18), overridable by env."""
import os

EXP_ROOT = os.environ.get("ASR_EXP_ROOT", os.path.join(os.path.expanduser("~"), "experiments", "sheet_retrieval"))
DATA_ROOT_MSMD = os.environ.get("ASR_DATA_ROOT_MSMD", "")
