"""VQA_Network with the reference's constructor / forward contract (models/model.py:18-121).

For Swin keys the backbone and head are fused into one libkvq_b200.so call; the module tree
(`<key>_backbone`, `<key>_head`) and state_dict names are the reference's, so `load_state_dict` of a reference
checkpoint (with or without the DataParallel `module.` prefix stripped by the caller) works unchanged."""
from functools import reduce

import torch
import torch.nn as nn

from .backbones.swin_backbone import SwinTransformer3D as VideoBackbone
from .backbones.swin_backbone import swin_3d_small, swin_3d_tiny
from .backbones.KSVQE_model import KSVQE as KSVQE_Backbone
from .backbones.simpleVQA_model import resnet50 as simpleVQA_Backbone
from .head import VQAHead, simpleVQAHead

__all__ = ["VQA_Network", "VideoBackbone", "VQAHead", "simpleVQAHead", "simpleVQA_Backbone", "swin_3d_tiny",
           "swin_3d_small"]

SWIN_KEYS = ("swin_tiny", "swin_tiny_grpb", "swin_tiny_grpb_m", "swin_small")


class VQA_Network(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.key_names = []
        self.multi = False
        self.layer = -1
        # None = follow env KVQ_CUDA_GRAPH; True = replay each (input buffer, shape) forward from a captured CUDA graph
        self.use_cuda_graph = None
        for key, hypers in config["model"]["args"].items():
            if key == "swin_tiny":
                backbone = swin_3d_tiny(**hypers.get("backbone", {}))
            elif key == "swin_tiny_grpb":
                backbone = VideoBackbone()                                                  # model.py:34-38
            elif key == "swin_tiny_grpb_m":
                backbone = VideoBackbone(window_size=(4, 4, 4), frag_biases=[0, 0, 0, 0])    # model.py:39-43
            elif key == "swin_small":
                backbone = swin_3d_small(**hypers.get("backbone", {}))
            elif key == "simpleVQA":
                backbone = simpleVQA_Backbone(pretrained=True)                              # model.py:52-55
            elif key == "KSVQE":                                                            # model.py:56-68
                hb = hypers["backbone"]
                backbone = KSVQE_Backbone(num_samples=hb["num_samples"], sample_type=hb["sample_type"],
                                          CLIP_location=hb["CLIP_location"], cls_use=hb["cls_use"],
                                          tuning_stage=hb["tuning_stage"], a1=hb["a1"], a2=hb["a2"],
                                          frozen_stages=hb["frozen_stages"])
            else:
                raise NotImplementedError(
                    f"kvq_b200: model key '{key}' is not on the B200 hot path yet (DESIGN.md, out-of-scope table)")
            head = simpleVQAHead(**hypers["head"]) if key == "simpleVQA" else VQAHead(**hypers["head"])
            self.key_names.append(key)
            setattr(self, key + "_backbone", backbone)
            setattr(self, key + "_head", head)

    def forward(self, inputs, targets=None, inference=True, return_pooled_feats=False, reduce_scores=False,
                pooled=False, clip_return=False, **kwargs):
        if self.training:
            raise RuntimeError("kvq_b200: inference path only -- call model.eval() (trainer.py:305)")
        scores, feats, dis_contra_loss = [], {}, None
        for key in self.key_names:
            backbone, head = getattr(self, key + "_backbone"), getattr(self, key + "_head")
            # the ResNet reads batch['simpleVQA'] + ['feat'], KSVQE 'fragment' / 'resize_video' / 'dis_label'
            x = inputs if key in ("simpleVQA", "KSVQE") else inputs["technical"]
            feat, score = backbone.forward_with_head(x, head, want_feat=return_pooled_feats, graph=self.use_cuda_graph)
            if key == "KSVQE":
                dis_contra_loss = backbone.last_dis_contra_loss
            scores.append(score)
            if return_pooled_feats:
                feats[key] = feat
        if reduce_scores:
            scores = reduce(lambda a, b: a + b, scores) if len(scores) > 1 else scores[0]
        if return_pooled_feats:                                     # model.py:109-121
            return (scores, feats, dis_contra_loss) if dis_contra_loss is not None else (scores, feats)
        return (scores, dis_contra_loss) if dis_contra_loss is not None else scores
