"""oracle/fragments.py against the golden vectors the REAL get_spatial_fragments produced (tools/make_golden_extra.py):
the reference draws its offsets from the global torch RNG, so the oracle replays the same seed."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fragments

CASES = sorted(glob.glob(os.path.join(GOLDEN, "fragments_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_fragments_match_reference_golden(path):
    g = np.load(path)
    shape = tuple(int(v) for v in g["shape"])
    kw = {k: int(g[k]) for k in ("fragments_h", "fragments_w", "fsize_h", "fsize_w", "aligned")}
    video = torch.randint(0, 256, shape, generator=torch.Generator().manual_seed(int(g["seed"]))).float()
    torch.manual_seed(int(g["seed"]))
    rnd_h, rnd_w = fragments.draw_offsets(shape[2], shape[3], shape[1], **kw)
    out = fragments.spatial_fragments(video, rnd_h, rnd_w, **kw)
    o = out.numpy()
    assert hashlib.sha256(np.ascontiguousarray(o).tobytes()).hexdigest() == str(g["sha256"])     # bit-exact, whole array
    if g["out"].shape != o.shape:
        o = o[:, ::4, ::4, ::4]
    np.testing.assert_array_equal(o, g["out"].astype(np.float32))


def test_fragment_clip_contract():
    frames = torch.randint(0, 256, (2, 8, 3, 70, 80), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
    offs = torch.zeros(2, 2, 2, 2, 2, dtype=torch.int32)
    offs[1, 0] = 3
    out = fragments.fragment_clip(frames, offs, 2, 2, 32, 4)
    assert out.shape == (2, 3, 8, 64, 64)
    want = (frames[1, 5, 1, 35 + 3, 40].float() - fragments.MEAN[1]) / fragments.STD[1]
    assert abs(out[1, 1, 5, 32, 32].item() - want.item()) < 1e-6
