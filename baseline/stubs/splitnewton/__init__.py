"""Stand-in for the third-party `splitnewton` package (un-vendored, unpinned, absent - SURVEY.md 8c): the
reference's steady_state (simulation.py:515-523) is driven by this repository's own restatement in
oracle/newton.py.  PARITY UNPINNED against upstream."""
