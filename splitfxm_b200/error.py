"""The package's one exception type.  Every error of the reference is ``splitfxm.error.SFXM``; callers catch it by
that name, so it keeps it."""


class SFXM(Exception):
    """Raised for invalid input (sizes, boundary specifications, schemes) and for every CUDA-side failure."""
