"""Dataset registry of the drop-in (`getattr(datasets, cfg['data']['val']['type'])(args, None)`, trainer.py:117).

Video decoding (decord / cv2) and the KVQ annotation files are outside the B200 hot path, so the
datasets shipped are synthetic ones producing the same item dicts as ViewDecompositionDataset_KVQ
(datasets/fusion_datasets.py:930-1050) and ViewDecompositionDataset_add_forSimpleVQA (:780-930).  Register real datasets with `datasets.register(cls)`.
The view functions those datasets apply to the decoded frames (get_resized_video, get_resizecrop_video,
UnifiedFrameSampler) are in `datasets.views`, running on the GPU."""
from . import views  # noqa: F401
from .features import load_motion_features  # noqa: F401
from .views import UnifiedFrameSampler, get_resized_video, get_resizecrop_video  # noqa: F401
from .synthetic import SyntheticFragmentDataset, SyntheticKSVQEDataset, SyntheticSimpleVQADataset  # noqa: F401


def register(cls, name=None):
    globals()[name or cls.__name__] = cls
    return cls
