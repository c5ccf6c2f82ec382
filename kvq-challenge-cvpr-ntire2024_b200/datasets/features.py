"""Reader for the per-clip motion features SlowFast_features.py writes (`feature_<i>_{slow,fast}_feature.npy`, shapes
[1,2048,1,1,1] / [1,256,1,1,1]) -- the `batch['feat']` input of the SimpleVQA branch
(ViewDecompositionDataset_add_forSimpleVQA.__getitem__, datasets/fusion_datasets.py:859-890)."""
import os

import numpy as np
import torch

WIDTH = {"Slow": 2048, "Fast": 256, "SlowFast": 2048 + 256}


def load_motion_features(folder, feature_type="SlowFast", num_clips=8):
    """-> float32 [num_clips, 2048 | 256 | 2304]: clip i = cat(squeeze(slow_i), squeeze(fast_i)) (:883-890)."""
    if feature_type not in WIDTH:
        raise ValueError(f"feature_type must be one of {sorted(WIDTH)}, got {feature_type!r}")
    out = torch.zeros([num_clips, WIDTH[feature_type]])
    for i in range(num_clips):
        parts = []
        if feature_type in ("Slow", "SlowFast"):
            parts.append(torch.from_numpy(np.load(os.path.join(folder, f"feature_{i}_slow_feature.npy"))).squeeze())
        if feature_type in ("Fast", "SlowFast"):
            parts.append(torch.from_numpy(np.load(os.path.join(folder, f"feature_{i}_fast_feature.npy"))).squeeze())
        out[i] = torch.cat(parts) if len(parts) > 1 else parts[0]
    return out
