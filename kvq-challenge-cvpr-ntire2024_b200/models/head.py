"""VQAHead / simpleVQAHead with the reference's parameter names (models/head.py:10-68)."""
import torch
import torch.nn as nn

from kvq_b200 import ops


class simpleVQAHead(nn.Module):
    """quality = Linear(in,hidden) -> Linear(hidden,1), mean over frames (head.py:10-31).  No activation separates
    the two Linears, so they are folded into one affine map at pack time and run as one row-dot kernel."""

    def __init__(self, in_channels=4096 + 2048 + 1024 + 2048 + 256, hidden_channels=128):
        super().__init__()
        self.quality = nn.Sequential(nn.Linear(in_channels, hidden_channels), nn.Linear(hidden_channels, 1))
        self._packed = None
        self._key = None

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("kvq_b200: features must be CUDA tensors -- this path has no CPU fallback")
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._key != key:
            self._packed = ops.fold_simplevqa_head(dict(self.state_dict()), "")
            self._key = key
        B, T, F = x.shape
        with torch.cuda.device(x.device):
            return ops.rowdot_mean_f32(x.float().contiguous().view(B * T, F), *self._packed, group=T).reshape(B, 1)


class VQAHead(nn.Module):
    """1x1x1 Conv3d C->hidden, GELU, hidden->1, mean over (D,H,W)  (head.py:33-68).  Parameter containers only;
    the arithmetic runs in libkvq_b200.so (fused into the backbone call by VQA_Network, or stand-alone here)."""

    def __init__(self, in_channels=768, hidden_channels=64, num_class=1, dropout_ratio=0.5, pre_pool=False, **kwargs):
        super().__init__()
        if num_class != 1 or pre_pool or hidden_channels != 64:
            raise NotImplementedError("kvq_b200: VQAHead is built for num_class=1, pre_pool=False, hidden_channels=64")
        self.in_channels, self.hidden_channels = in_channels, hidden_channels
        self.fc_hid = nn.Conv3d(in_channels, hidden_channels, (1, 1, 1))
        self.fc_last = nn.Conv3d(hidden_channels, num_class, (1, 1, 1))
        self._packed = None
        self._key = None

    def forward(self, x, rois=None):
        if self.training:
            raise RuntimeError("kvq_b200: inference path only (call .eval(); dropout is identity there)")
        if not x.is_cuda:
            raise RuntimeError("kvq_b200: features must be CUDA tensors -- this path has no CPU fallback")
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._key != key:
            with torch.cuda.device(x.device):
                self._packed = (ops.cast_f16(self.fc_hid.weight.reshape(self.hidden_channels, -1)),
                                self.fc_hid.bias.detach().float().contiguous(),
                                self.fc_last.weight.detach().float().reshape(-1).contiguous(),
                                self.fc_last.bias.detach().float().reshape(-1).contiguous())
            self._key = key
        with torch.cuda.device(x.device):
            return ops.vqa_head(x.contiguous(), *self._packed)
