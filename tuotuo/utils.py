from tuotuo_b200.lda_model import np_clip_for_exp  # noqa: F401
from tuotuo_b200.text import text_pipeline  # noqa: F401
