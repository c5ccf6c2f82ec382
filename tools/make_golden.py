"""Generate tests/golden/*.npz by running the REAL reference (imported read-only from
/root/reference, see oracle/ref_import.py) on seeded synthetic weights and inputs.

Run in the authoring container only:   python tools/make_golden.py [--only swin|head|frag|...]
The reference tree does not travel to the GPU box; these small fixtures do.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (name, input shape [B,3,T,H,W], weight seed, input seed, model kwargs)
SWIN_CASES = [
    ("swin_t16_64x64", (1, 3, 16, 64, 64), 11, 21, {}),          # clamped windows + H/W padding
    ("swin_t8_96x128", (2, 3, 8, 96, 128), 12, 22, {}),          # D < window, ragged H != W
    ("swin_t18_100x76", (1, 3, 18, 100, 76), 13, 23, {}),        # patch-embed pad, D pad, odd merge
    ("swin_t32_224", (2, 3, 32, 224, 224), 14, 3, {}),           # BASELINE config 2 geometry
    ("swin_t96_224", (1, 3, 96, 224, 224), 15, 25, {}),          # 96-frame val clip
    ("swin_plain_t16_112", (1, 3, 16, 112, 112), 16, 26, {"frag_biases": [0, 0, 0, 0]}),  # swin_3d_tiny
    # swin_tiny_grpb_m (model.py:39-43): window (4,4,4), no fragment tables
    ("swin_m444_t16_96", (1, 3, 16, 96, 96), 17, 27, {"frag_biases": [0, 0, 0, 0], "window_size": (4, 4, 4)}),
    # swin_small (swin_backbone.py:1093-1095): depths [2,2,18,2]
    ("swin_small_t16_64", (1, 3, 16, 64, 64), 18, 28, {"frag_biases": [0, 0, 0, 0], "depths": [2, 2, 18, 2]}),
    # BASELINE config 2 exactly: batch 8 clips of 32x224x224 (round 2: the batch-2 case above left batch 8 unpinned)
    ("swin_b8_t32_224", (8, 3, 32, 224, 224), 19, 29, {}),
]


def gen_swin(ref, case=None):
    keys_written = case is not None          # a single-case run leaves the key list alone
    for name, shape, wseed, xseed, kw in SWIN_CASES:
        if case is not None and name != case:
            continue
        fb = kw.get("frag_biases", [True, True, True, False])
        window = tuple(kw.get("window_size", (8, 7, 7)))
        depths = list(kw.get("depths", [2, 2, 6, 2]))
        m = ref.swin.SwinTransformer3D(pretrained=None, use_checkpoint=False, frag_biases=fb, window_size=window,
                                       depths=depths)
        head = ref.head.VQAHead(in_channels=768, hidden_channels=64)
        sd = synth.synth_state_dict(synth.swin_shapes(frag_biases=fb, window=window, depths=depths), wseed)
        hd = synth.synth_state_dict(synth.vqa_head_shapes(), wseed)
        missing = m.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing
        assert all("relative_position_index" in k for k in missing.missing_keys), missing
        head.load_state_dict(hd, strict=True)
        m.eval()
        head.eval()
        x = synth.clip_input(shape, xseed)
        with torch.no_grad():
            feat = m({"technical": x})
            score = head(feat)
        fs = feat[:, ::16].contiguous() if feat.numel() > 400000 else feat
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), score=score.numpy(),
                            feat=fs.numpy().astype(np.float32), feat_stride=np.int64(16 if fs is not feat else 1),
                            feat_shape=np.array(feat.shape), shape=np.array(shape), wseed=wseed, xseed=xseed,
                            frag_biases=np.array([int(bool(b)) for b in fb]), window=np.array(window),
                            depths=np.array(depths),
                            feat_absmean=feat.abs().mean().numpy())
        print(name, tuple(feat.shape), "score", score.flatten().tolist(), "absmean", feat.abs().mean().item())
        if not keys_written:
            full = m.state_dict()
            spec = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in full.items()}
            spec_h = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in head.state_dict().items()}
            with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
                json.dump({"SwinTransformer3D": spec, "VQAHead": spec_h}, f, indent=0, sort_keys=True)
            keys_written = True


# forward options of SwinTransformer3D.forward (:1044-1080): (name, shape, wseed, xseed, forward kwargs)
OPTION_CASES = [
    ("swinopt_adaptive_t16_112", (1, 3, 16, 112, 112), 31, 41, {"adaptive_window_size": True}),   # window (4,3,3), shift (4,3,3)
    ("swinopt_adaptive_t32_160", (1, 3, 32, 160, 160), 32, 42, {"adaptive_window_size": True}),   # window (8,5,5)
    ("swinopt_layer2_t16_64", (1, 3, 16, 64, 64), 33, 43, {"layer": 2}),
    ("swinopt_multi_t16_64", (1, 3, 16, 64, 64), 34, 44, {"multi": True}),
]


def gen_swin_options(ref):
    for name, shape, wseed, xseed, kw in OPTION_CASES:
        fb = [True, True, True, False]
        m = ref.swin.SwinTransformer3D(pretrained=None, use_checkpoint=False, frag_biases=fb)
        sd = synth.synth_state_dict(synth.swin_shapes(frag_biases=fb), wseed)
        m.load_state_dict(sd, strict=False)
        m.eval()
        x = synth.clip_input(shape, xseed)
        with torch.no_grad():
            out = m({"technical": x}, **kw)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), out=out.numpy().astype(np.float32), shape=np.array(shape),
                            wseed=wseed, xseed=xseed, kwargs=json.dumps(kw))
        print(name, tuple(out.shape), "absmean", out.abs().mean().item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    ap.add_argument("--case", default=None, help="with --only swin: regenerate just this fixture")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_import.load_reference()
    gens = {"swin": gen_swin, "swinopt": gen_swin_options}
    try:
        from tools import make_golden_extra  # widened rows (fragments, simpleVQA, ...) live there
        gens.update(make_golden_extra.GENERATORS)
    except ImportError:
        pass
    for k, fn in gens.items():
        if args.only in ("all", k):
            if k == "swin" and args.case:
                fn(ref, args.case)
            else:
                fn(ref)


if __name__ == "__main__":
    main()
