"""Seeded synthetic weights / inputs live in tools/synth.py (plain data generation, no reference arithmetic) so that
bench.py's product path does not import anything under oracle/; the checkers re-export it here."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import _gen  # noqa: F401
