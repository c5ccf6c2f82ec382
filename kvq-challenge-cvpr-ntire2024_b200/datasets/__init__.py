"""Dataset registry of the drop-in (`getattr(datasets, cfg['data']['val']['type'])(args, None)`, trainer.py:117).

Video decoding (decord / cv2) and the KVQ annotation files are outside the B200 hot path, so the only
dataset shipped is a synthetic one that produces the same item dict as ViewDecompositionDataset_KVQ
(datasets/fusion_datasets.py:930-1050).  Register real datasets with `datasets.register(cls)`."""
from .features import load_motion_features  # noqa: F401
from .synthetic import SyntheticFragmentDataset  # noqa: F401


def register(cls, name=None):
    globals()[name or cls.__name__] = cls
    return cls
