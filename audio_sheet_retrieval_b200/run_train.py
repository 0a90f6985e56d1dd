"""Glue shared by every script of the reference (audio_sheet_retrieval/run_train.py:19-48):
select_model, select_data, compile_tag.  Training itself is out of scope."""
from __future__ import print_function

import importlib
import os

from .utils import mutopia_data


def select_model(model_path):
    """ select model (run_train.py:19-29; the reference also returns the train function) """
    model_str = os.path.basename(model_path)
    model_str = model_str.split('.py')[0]
    model = importlib.import_module("audio_sheet_retrieval_b200.models." + model_str)
    model.EXP_NAME = model_str
    return model, None


def select_data(data_name, split_file, config_file, seed=23, test_only=False):
    """ select train data (run_train.py:32-41)

    'mutopia' is the MSMD data set of the reference and needs its `msmd` package (absent from this image);
    it is never silently replaced: only the explicit name 'synthetic' returns the seeded synthetic pools
    (same pool protocol, utils/mutopia_data.py), and results computed on them are tagged as such.
    """
    if str(data_name) == "synthetic":
        return mutopia_data.load_audio_score_retrieval(split_file=split_file, config_file=config_file,
                                                       test_only=test_only, seed=seed)
    if str(data_name) == "mutopia":
        return mutopia_data.load_msmd_audio_score_retrieval(split_file=split_file, config_file=config_file,
                                                            test_only=test_only)
    raise ValueError("unknown data set %r (known: 'mutopia' = MSMD via the msmd package, 'synthetic')" % (data_name,))


def is_synthetic(data_name):
    return str(data_name) == "synthetic"


SYNTHETIC_BANNER = ("\n" + "!" * 78 + "\n!! --data synthetic: seeded SYNTHETIC pairs, not MSMD.  Numbers and refitted projections\n"
                    "!! computed here say nothing about real data; output files carry the tag 'synthetic'.\n" + "!" * 78 + "\n")


def compile_tag(train_split, config):
    """ compile model tag from split and config file paths (run_train.py:44-48) """
    tag = os.path.splitext(os.path.basename(train_split))[0]
    tag += "_" + os.path.splitext(os.path.basename(config))[0]
    return tag
