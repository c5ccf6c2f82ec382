"""Golden vectors for the widened rows (SURVEY.md section 8 'next'): SimpleVQA spatial branch.

Imported by tools/make_golden.py; runs the REAL reference modules on CPU in the authoring container."""
import json
import os

import numpy as np
import torch

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# (name, [B,3,T,H,W], weight seed, input seed)
SIMPLEVQA_CASES = [
    ("simplevqa_t2_64", (1, 3, 2, 64, 64), 31, 41),           # smallest legal map: layer4 output 2x2
    ("simplevqa_t4_96x160", (2, 3, 4, 96, 160), 32, 42),      # ragged H != W, batch of 2
    ("simplevqa_t8_224", (1, 3, 8, 224, 224), 33, 43),        # clip_len 8 (kwai_simpleVQA_test.yml), half resolution
    ("simplevqa_t3_100x76", (1, 3, 3, 100, 76), 34, 44),      # odd sizes: every stride-2 stage rounds
]


def gen_simplevqa(ref):
    for name, shape, wseed, xseed in SIMPLEVQA_CASES:
        sd = synth.simplevqa_network_state_dict(wseed)
        backbone = ref.simplevqa.resnet50(pretrained=False)
        head = ref.head.simpleVQAHead(in_channels=9472, hidden_channels=128)
        bsd = {k[len("simpleVQA_backbone."):]: v for k, v in sd.items() if k.startswith("simpleVQA_backbone.")}
        hsd = {k[len("simpleVQA_head."):]: v for k, v in sd.items() if k.startswith("simpleVQA_head.")}
        missing = backbone.load_state_dict(bsd, strict=False)
        assert not missing.unexpected_keys, missing
        assert all("num_batches_tracked" in k or k.startswith("fc.") for k in missing.missing_keys), missing
        head.load_state_dict(hsd, strict=True)
        backbone.eval()
        head.eval()
        if name == SIMPLEVQA_CASES[0][0]:
            spec = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in backbone.state_dict().items()}
            spec_h = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in head.state_dict().items()}
            with open(os.path.join(GOLD, "state_dict_keys_simplevqa.json"), "w") as f:
                json.dump({"ResNet": spec, "simpleVQAHead": spec_h}, f, indent=0, sort_keys=True)
        x = synth.clip_input(shape, xseed)
        feat3d = synth.motion_features((shape[0], shape[2], 2304), xseed + 1000)
        with torch.no_grad():
            feats = backbone({"simpleVQA": x, "feat": feat3d})
            score = head(feats)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), score=score.numpy(), feat=feats.numpy(),
                            shape=np.array(shape), wseed=wseed, xseed=xseed)
        print(name, tuple(feats.shape), "score", score.flatten().tolist(), "feat absmean",
              feats[..., :7168].abs().mean().item(), "max", feats[..., :7168].abs().max().item())


# (name, [C,T,H,W], kwargs, seed)
FRAGMENT_CASES = [
    ("fragments_7x7_t16_300x270", (3, 16, 300, 270), dict(fragments_h=7, fragments_w=7, fsize_h=32, fsize_w=32, aligned=8), 51),
    ("fragments_2x3_t8_100x150", (3, 8, 100, 150), dict(fragments_h=2, fragments_w=3, fsize_h=32, fsize_w=32, aligned=4), 52),
    ("fragments_upsample_t4_60x90", (3, 4, 60, 90), dict(fragments_h=3, fragments_w=3, fsize_h=32, fsize_w=32, aligned=2), 53),
    ("fragments_tight_t8_64x96", (3, 8, 64, 96), dict(fragments_h=2, fragments_w=3, fsize_h=32, fsize_w=32, aligned=8), 54),
]


def gen_fragments(ref):
    """datasets/fusion_datasets.get_spatial_fragments on uint8-valued frames, offsets drawn from the seeded GLOBAL RNG
    (the reference has no generator argument)."""
    for name, shape, kw, seed in FRAGMENT_CASES:
        video = torch.randint(0, 256, shape, generator=torch.Generator().manual_seed(seed)).float()
        torch.manual_seed(seed)
        out = ref.fusion.get_spatial_fragments(video, **kw)
        import hashlib
        o = out.numpy()
        digest = hashlib.sha256(np.ascontiguousarray(o).tobytes()).hexdigest()
        if o.size > 200000:        # large case: keep every 4th frame / pixel as uint8 plus the digest of the whole array
            assert (o == np.round(o)).all()
            o = o[:, ::4, ::4, ::4].astype(np.uint8)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), out=o, sha256=np.array(digest), shape=np.array(shape),
                            seed=seed, **{k: np.array(v) for k, v in kw.items()})
        print(name, tuple(out.shape), float(out.mean()))


GENERATORS = {"simplevqa": gen_simplevqa, "fragments": gen_fragments}
