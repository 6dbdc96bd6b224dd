"""Which end of the 1-D domain a ghost cell belongs to, and the key that end has in a boundary-condition dict
(``{"left": ..., "right": ...}``).  Same names and values as the reference's splitfxm/constants.py."""
import enum


class btype(enum.Enum):
    LEFT = 1
    RIGHT = 2


# dict key of each side in the reference's `bcs` dictionaries (bc.py:40-45)
btype_map = {side: side.name.lower() for side in btype}
