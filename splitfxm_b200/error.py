class SFXM(Exception):
    """Error type of the package (same name as the reference's splitfxm/error.py:1)."""
