from oracle.newton import split_newton  # noqa: F401
