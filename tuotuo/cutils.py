from tuotuo_b200.cutils import _dirichlet_expectation_1d, _dirichlet_expectation_2d  # noqa: F401
