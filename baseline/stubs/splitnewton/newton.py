from oracle.newton import newton  # noqa: F401
