"""Model module `mutopia_ccal_cont`: full-resolution sheet input, 12/24/48/48 filters.

Same module protocol as audio_sheet_retrieval/models/mutopia_ccal_cont.py: callers read
`build_model`, `prepare`, `DIM_LATENT`, `INPUT_SHAPE_1/2`, `BATCH_SIZE`, `EXP_NAME` (injected by
`select_model`).  Hyper-parameters that only matter for training are kept as documentation of the
reference configuration (:23-51); training itself is out of scope.
"""
import sys

from . import _common
from .. import _lib
from ._common import SPEC_CONTEXT

INI_LEARNING_RATE = 0.002
REFINEMENT_STEPS = 5
LR_MULTIPLIER = 0.5
BATCH_SIZE = 100
MOMENTUM = 0.9
MAX_EPOCHS = 1000
PATIENCE = 30
INPUT_SHAPE_1 = [1, 160, 200]
INPUT_SHAPE_2 = [1, 92, SPEC_CONTEXT]

DIM_LATENT = 32

L1 = None
L2 = 0.00001
GRAD_NORM = None

r1 = r2 = 1e-3
rT = 1e-3

FIT_CCA = False
ALPHA = 1.0
WEIGHT_TNO = 0.0
USE_CCAL = True
GAMMA = 0.7

NUM_FILTERS_1 = 12

build_model = _common.make_build_model(sys.modules[__name__], NUM_FILTERS_1, half_res=False)


def prepare(x, y=None):
    """prepare images for the network (host/NumPy form; fused on the device when used through
    RetrievalWrapper / batch_compute)."""
    return _common.prepare_scale(x, y)


prepare.asr_prepare_mode = _lib.PREP_SCALE   # lets the library fuse this function into layer 0
