"""Import the UNMODIFIED reference (lixinustc/KVQ-Challenge-CVPR-NTIRE2024) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists so that `tools/make_golden.py` can run the real
reference modules on CPU *in the authoring container* and dump golden vectors under
`tests/golden/`.  `/root/reference` does not exist on the GPU box, so nothing in `tests/ -m gpu`,
`bench.py` or `__graft_entry__.smoke()` may import this file.

The reference cannot be imported as shipped:
  * `models/backbones/swin_backbone.py:1108` builds a SwinTransformer3D at import time and
    `torch.load`s `pretrained_weights/swin_tiny_patch244_window877_kinetics400_1k.pth`
    relative to cwd  -> we chdir into a scratch dir holding a stub `{'state_dict': {}}`.
  * `timm`, `thop`, `ftfy`, `decord`, `turtle` (tkinter) are absent and
    `torchvision.io.write_video` was removed -> `sys.modules` shims (identity at eval time).
  * `simpleVQA_model.resnet50(pretrained=True)` downloads weights -> `model_zoo.load_url -> {}`.
Nothing under /root/reference is modified or copied.
"""
import contextlib
import os
import sys
import tempfile
import types

import torch

REFERENCE_ROOT = os.environ.get("KVQ_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _install_shims():
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        registry = types.ModuleType("timm.models.registry")

        class DropPath(nn.Module):  # identity in eval; the oracle only ever runs eval
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
        registry.register_model = lambda fn: fn
        timm.models = timm_models
        timm_models.layers = layers
        timm_models.registry = registry
        sys.modules.update({"timm": timm, "timm.models": timm_models,
                            "timm.models.layers": layers, "timm.models.registry": registry})
    if "thop" not in sys.modules:
        thop = types.ModuleType("thop")
        thop.profile = lambda *a, **k: (0, 0)
        sys.modules["thop"] = thop
    if "ftfy" not in sys.modules:
        ftfy = types.ModuleType("ftfy")
        ftfy.fix_text = lambda s: s
        sys.modules["ftfy"] = ftfy
    try:
        import turtle  # noqa: F401
    except Exception:
        t = types.ModuleType("turtle")
        t.forward = lambda *a, **k: None
        sys.modules["turtle"] = t
    if "decord" not in sys.modules:
        decord = types.ModuleType("decord")
        decord.bridge = types.SimpleNamespace(set_bridge=lambda *_: None)
        decord.VideoReader = object
        decord.cpu = lambda *a: None
        decord.gpu = lambda *a: None
        sys.modules["decord"] = decord
    import torchvision.io as tvio
    if not hasattr(tvio, "write_video"):
        tvio.write_video = lambda *a, **k: None


_REF = {}


@contextlib.contextmanager
def _in_scratch_cwd():
    old = os.getcwd()
    d = tempfile.mkdtemp(prefix="kvq_ref_")
    os.makedirs(os.path.join(d, "pretrained_weights"), exist_ok=True)
    torch.save({"state_dict": {}},
               os.path.join(d, "pretrained_weights", "swin_tiny_patch244_window877_kinetics400_1k.pth"))
    os.chdir(d)
    try:
        yield d
    finally:
        os.chdir(old)


def load_reference():
    """Returns a namespace with the reference's `models` package and fragment sampler."""
    if _REF:
        return types.SimpleNamespace(**_REF)
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_shims()
    # our drop-in package is also called `models`; make sure the reference's wins here
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")
              or k == "datasets" or k.startswith("datasets.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        with _in_scratch_cwd():
            import importlib
            swin = importlib.import_module("models.backbones.swin_backbone")
            svqa = importlib.import_module("models.backbones.simpleVQA_model")
            svqa.model_zoo.load_url = lambda *a, **k: {}
            head = importlib.import_module("models.head")
            model = importlib.import_module("models.model")
            fusion = importlib.import_module("datasets.fusion_datasets")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _REF.update(swin=swin, simplevqa=svqa, head=head, model=model, fusion=fusion)
    # keep the reference modules out of the global namespace so `import models` afterwards
    # resolves to whichever path the caller puts on sys.path
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")
              or k == "datasets" or k.startswith("datasets.")]:
        del sys.modules[k]
    return types.SimpleNamespace(**_REF)


def randomise_(module, seed, table_std=0.2):
    """default init leaves GRPB tables, biases, LN/BN affine and BN stats
    at values that hide bugs; draw all of them from a seeded generator instead."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if "position_bias_table" in name:
                p.copy_(torch.randn(p.shape, generator=g) * table_std)
            elif name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif ("norm" in name or "bn" in name or "downsample.1" in name) and name.endswith(".weight") and p.dim() == 1:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
        for name, b in module.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    return module
