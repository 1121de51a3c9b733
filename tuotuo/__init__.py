"""Drop-in import path: ``from tuotuo.lda_model import LDASmoothed`` etc. (reference readme.md:18,35) resolves to
the B200-native implementation in ``tuotuo_b200``."""
