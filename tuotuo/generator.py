from tuotuo_b200.generator import doc_generator  # noqa: F401
