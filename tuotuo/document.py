from tuotuo_b200.document import Document  # noqa: F401
