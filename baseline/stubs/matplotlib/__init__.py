"""Stand-in for `matplotlib` (absent from this image): the reference imports pyplot in visualize.py only."""
