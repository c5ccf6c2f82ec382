"""Golden vectors for the literal `KSVQE` key (CLIP ViT-B/16 + adapters, QRS region selection,
CONTRIQUE ResNet-50, CDM, Swin3D-GRPB, VQAHead) -- NOT built on the B200 path yet; this pins the target for it.

Runs the REAL reference modules on CPU in the authoring container (oracle/ref_import.py shims; CLIP built from its
class with seeded weights because no checkpoint can be downloaded; CONTRIQUE / Swin checkpoints replaced by seeded
weights keyed by parameter name, tools/synth.fill_like).  Output: tests/golden/ksvqe_t32_288.npz (small tensors and
statistics of the large ones) + tests/golden/state_dict_keys_ksvqe.json.
    python tools/make_golden_ksvqe.py"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from tools import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WSEED, XSEED = 61, 62


def build_reference_ksvqe():
    ref_import._install_shims()
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    try:
        with ref_import._in_scratch_cwd():
            clipb = importlib.import_module("models.backbones.CLIP_backbone")
            clip_model = importlib.import_module("models.backbones.clip.model")
            # clip.load() downloads ViT-B/16: build the same architecture from its class instead (SURVEY 8c recipe)
            clipb.load_clip_to_cpu = lambda *a, **k: clip_model.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval()
            km = importlib.import_module("models.backbones.KSVQE_model")
            head_mod = importlib.import_module("models.head")
            real_load = torch.load

            def patched_load(path, *a, **k):
                p = str(path)
                if "contrique" in p.lower():
                    return km.CONTRIQUE_model(km.get_network("resnet50", pretrained=False), 2048).state_dict()
                if "swin" in p:
                    return {"state_dict": {}}
                return real_load(path, *a, **k)

            torch.load = patched_load
            try:
                # Kwai_KSVQE_test.yml backbone args (+ a1 / a2, which the YAML omits and model.py:66-67 requires)
                m = km.KSVQE(num_samples=1, sample_type="topkpertubation", CLIP_location=8, cls_use=True, tuning_stage=2,
                             a1=1.0, a2=1.0, frozen_stages=-1, use_checkpoint=False)
            finally:
                torch.load = real_load
            head = head_mod.VQAHead(in_channels=768, hidden_channels=64)
    finally:
        sys.path.remove(ref_import.REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
    return m, head


def seed_weights(module, prefix, seed):
    sd = module.state_dict()
    new = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            new[k] = v
            continue
        t = synth.fill_like(prefix + k, tuple(v.shape), seed)
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("a1", "a2"):                       # CDM mixing weights: keep the configured 1.0
            t = torch.ones_like(v)
        elif leaf in ("class_embedding", "positional_embedding", "logit_scale", "sigma"):
            t = v.clone() if leaf in ("logit_scale", "sigma") else t
        new[k] = t.to(v.dtype).reshape(v.shape)
    module.load_state_dict(new, strict=True)
    return new


def stats(t):
    t = t.detach().float()
    return np.array([t.mean().item(), t.abs().mean().item(), t.abs().max().item(), t.std().item()], dtype=np.float64)


# (file name, clips, frames, input seed, dis_label): one 32-frame clip (the training geometry), a batch of two clips with
# different distortion labels, and one 96-frame item (what inferece_test feeds: trainer.py never splits the KSVQE views
# into clips because 'KSVQE' is not a key of the data dict)
CASES = [("ksvqe_t32_288", 1, 32, XSEED, [0]), ("ksvqe_b2_t32_288", 2, 32, XSEED + 1, [0, 3]),
         ("ksvqe_t96_288", 1, 96, XSEED + 2, [0])]


def main():
    m, head = build_reference_ksvqe()
    seed_weights(m, "KSVQE_backbone.", WSEED)
    seed_weights(head, "KSVQE_head.", WSEED)
    m.eval()
    head.eval()
    for i, case in enumerate(CASES):
        run_case(m, head, *case, write_keys=(i == 0))


def run_case(m, head, name, B, T, xseed, labels, write_keys):
    g = torch.Generator().manual_seed(xseed)
    x = {"fragment": torch.randn((B, 3, T, 288, 288), generator=g),           # 9x9 grid of 32x32 patches, normalised
         "resize_video": torch.randn((B, 3, T, 112, 112), generator=g),       # CLIP-normalised 112x112 view
         "dis_label": torch.tensor(labels, dtype=torch.long)}
    cap = {}
    hooks = [m.CLIP_tool.register_forward_hook(lambda mod, i, o: cap.__setitem__("clip", o)),
             m.spa_patchnet.register_forward_hook(lambda mod, i, o: cap.__setitem__("x_sel_ori", o)),
             m.distortion_tool.register_forward_hook(lambda mod, i, o: cap.__setitem__("dist_token", o))]
    for l, layer in enumerate(m.layers):
        hooks.append(layer.register_forward_hook(lambda mod, i, o, l=l: cap.__setitem__(f"stage{l}", o)))
    # CDM internals of both modulated stages (bring-up checkpoints for the B200 path)
    first = lambda o: o[0] if isinstance(o, (tuple, list)) else o
    for mod_name in ("semantic_adapter", "semantic_cross", "semantic_mod", "distortion_adapter", "distortion_cross",
                     "distortion_self", "distortion_mod"):
        for i in range(2):
            hooks.append(getattr(m, mod_name)[i].register_forward_hook(
                lambda mod, inp, o, k=f"{mod_name}{i}": cap.__setitem__(k, first(o))))
    hooks.append(m.dist_adapter.register_forward_hook(lambda mod, i, o: cap.__setitem__("dist_adapter", o)))
    with torch.no_grad():
        feat, loss = m(x)
        score = head(feat)
    for h in hooks:
        h.remove()
    cls_attn = cap["clip"][0]                                                  # [4B, 49] cos(cls, patch) per key frame
    x_sel = cap["x_sel_ori"]                                                   # [B,3,32,224,224] selected 7x7 region
    # which of the 3x3 candidate regions each frame took: compare against the nine crops of the 9x9 fragment grid
    frag = x["fragment"]
    region = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        for t in range(T):
            hit = [(ry, rx) for ry in range(3) for rx in range(3)
                   if torch.equal(frag[b, :, t, 32 * ry:32 * ry + 224, 32 * rx:32 * rx + 224], x_sel[b, :, t])]
            region[b, t] = hit[0][0] * 3 + hit[0][1] if hit else -1
    out = {"score": score.numpy(), "loss": np.array(float(loss)), "feat_stats": stats(feat),
           "feat_slice": feat[0, :8, 0, :, :].numpy(), "cls_attn": cls_attn.numpy(), "region": region,
           "dist_token_stats": stats(cap["dist_token"]), "dist_token_slice": cap["dist_token"][0, 0, :4, :8].numpy(),
           "wseed": WSEED, "xseed": xseed, "B": B, "T": T, "labels": np.array(labels)}
    for l in range(4):
        out[f"stage{l}_stats"] = stats(cap[f"stage{l}"])
    for k, v in cap.items():
        if k[:-1] in ("semantic_adapter", "semantic_cross", "semantic_mod", "distortion_adapter", "distortion_cross",
                      "distortion_self", "distortion_mod") or k == "dist_adapter":
            out[k + "_stats"] = stats(v)
            out[k + "_slice"] = v.detach().float().reshape(-1, v.shape[-1])[:3, :6].numpy()
            out[k + "_shape"] = np.array(v.shape)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    if write_keys:
        spec = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()}
        with open(os.path.join(GOLD, "state_dict_keys_ksvqe.json"), "w") as f:
            json.dump({"KSVQE": spec}, f, indent=0, sort_keys=True)
    print("score", score.flatten().tolist(), "loss", float(loss), "feat", out["feat_stats"], "regions", region[0].tolist())
    print("cls_attn", tuple(cls_attn.shape), "dist_token", tuple(cap["dist_token"].shape), out["dist_token_stats"])
    for l in range(4):
        print("stage", l, tuple(cap[f"stage{l}"].shape), out[f"stage{l}_stats"])


if __name__ == "__main__":
    main()
