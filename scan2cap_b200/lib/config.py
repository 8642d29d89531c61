"""The four constants of the reference's lib/config.py that the hot path reads (config.py:62-71).
Dataset / output paths of the reference CONF are host-specific and out of scope."""


class _NS(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


CONF = _NS()
CONF.TRAIN = _NS()
CONF.TRAIN.MAX_DES_LEN = 30
CONF.TRAIN.SEED = 42
CONF.TRAIN.OVERLAID_THRESHOLD = 0.5
CONF.TRAIN.MIN_IOU_THRESHOLD = 0.25
CONF.TRAIN.NUM_BINS = 6
CONF.EVAL = _NS()
CONF.EVAL.MIN_IOU_THRESHOLD = 0.5
