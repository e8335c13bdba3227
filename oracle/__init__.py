"""CPU oracle of the hot path -- test infrastructure only (see stencils.py, oracle.c)."""
