"""SimpleVQA spatial branch with the reference's module tree and state_dict names
(models/backbones/simpleVQA_model.py:85-264, :307-321).  The modules are parameter containers: the arithmetic
(per-frame ResNet-50 with BatchNorm folded, mean / unbiased-std pools of layer2-4, concatenation with batch['feat'])
runs in libkvq_b200.so (kvq_simplevqa_forward)."""
import torch
import torch.nn as nn

from kvq_b200 import ops

__all__ = ["ResNet", "Bottleneck", "resnet50"]


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)     # stride on the 3x3 (:98)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    def __init__(self, block=Bottleneck, layers=(3, 4, 6, 3), feat3d_dim=2304):
        super().__init__()
        if block is not Bottleneck:
            raise NotImplementedError("kvq_b200: the SimpleVQA path is built for Bottleneck ResNets (resnet50)")
        self.layer_sizes, self.feat3d_dim = tuple(layers), feat3d_dim
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        # the reference backbone carries its own (unused, :258) regressor; kept so checkpoints load strictly
        self.quality = nn.Sequential(nn.Linear(4096 + 2048 + 1024 + 2048 + 256, 128), nn.Linear(128, 1))
        for m in self.modules():                                            # :171-176
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._packed = None
        self._packed_key = None

    def _make_layer(self, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * 4:                      # :196-200
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride, bias=False),
                                 nn.BatchNorm2d(planes * 4))
        mods = [Bottleneck(self.inplanes, planes, stride, down)]
        self.inplanes = planes * 4
        mods += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    def packed(self, head=None):
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("kvq_b200: the SimpleVQA ResNet runs on a CUDA device only (call .to('cuda')); "
                               "there is no CPU fallback")
        tensors = list(self.parameters()) + list(self.buffers()) + (list(head.parameters()) if head is not None else [])
        key = (tuple((t.data_ptr(), t._version) for t in tensors), id(head))
        if self._packed is None or self._packed_key != key:
            sd = dict(self.state_dict())
            hp = None
            if head is not None:
                sd.update({"__head__." + k: v for k, v in head.state_dict().items()})
                hp = "__head__."
            with torch.cuda.device(dev):
                self._packed = ops.SimpleVQAWeights(sd, dev, prefix="", head_prefix=hp, layers=self.layer_sizes,
                                                    feat3d_dim=self.feat3d_dim, eps=self.bn1.eps)
            self._packed_key = key
        return self._packed

    def _run(self, batch, head, graph):
        if self.training:
            raise RuntimeError("kvq_b200: inference path only -- call model.eval() (BatchNorm is folded)")
        x = batch["simpleVQA"]
        if not x.is_cuda:
            raise RuntimeError("kvq_b200: input frames must be CUDA tensors -- this path has no CPU fallback")
        with torch.cuda.device(x.device):
            return self.packed(head).forward(x, batch["feat"], graph=graph)

    def forward(self, batch, multi=None, layer=None):
        """batch['simpleVQA'] f32 [B,3,T,H,W], batch['feat'] [B,T,2304] -> [B,T,9472]  (:220-264)."""
        return self._run(batch, None, False)[0]

    def forward_with_head(self, batch, head, want_feat=False, graph=None):
        """Fused backbone + simpleVQAHead: one C-ABI call, score [B,1] (+ the [B,T,9472] features)."""
        feats, score = self._run(batch, head, graph)
        return (feats if want_feat else None), score.reshape(-1, 1)


def resnet50(pretrained=True, progress=True, **kwargs):
    """:307-321.  There is no network here: `pretrained` checkpoints come in through load_state_dict."""
    return ResNet(Bottleneck, (3, 4, 6, 3), **kwargs)
