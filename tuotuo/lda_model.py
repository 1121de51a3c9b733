from tuotuo_b200.lda_model import DTYPE, SEED, LDASmoothed, np_clip_for_exp  # noqa: F401
