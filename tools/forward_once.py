"""One seeded Swin3D forward through the C ABI; prints {"score": [...], "feat_sum": ..., "feat_abs_sum": ...} as JSON.
Used by tests/test_gpu_swin.py::test_kernel_path_switches_agree to compare the environment-selected kernel paths
(KVQ_EMBED_EXPLICIT, KVQ_LNQKV_SPLIT, KVQ_FUSED_MLP, KVQ_ATTN_VARIANT), which are read once per process.
Usage: python tools/forward_once.py <golden.npz>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from test_gpu_swin import _weights
    from oracle import synth
    g = np.load(sys.argv[1])
    dev = torch.device("cuda:0")
    wts, _ = _weights(g, dev)
    shape = tuple(int(v) for v in g["shape"])
    x = synth.clip_input(shape, int(g["xseed"])).to(dev)
    feat, score = wts.forward(x, want_feat=True)
    torch.cuda.synchronize()
    f = feat.double()
    print(json.dumps({"score": [float(v) for v in score.reshape(-1).cpu()], "feat_sum": float(f.sum()),
                      "feat_abs_sum": float(f.abs().sum()), "feat_sample": [float(v) for v in f.reshape(-1)[::9973][:64].cpu()]}))


if __name__ == "__main__":
    main()
