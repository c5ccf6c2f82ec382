"""B200-native hot path of the KVQ / KSVQE forward: thin host layer over libkvq_b200.so (sm_100a)."""
from . import lib  # noqa: F401
