"""Per-fixture parity table of the Swin3D path on the GPU: score error and feature max / mean error of the C-ABI
forward against every tests/golden/swin_*.npz (outputs of the REAL reference), written as JSON.
Usage on the GPU box: python tools/parity_table.py gpurun_out/r02_parity.json   (then copy into profiles/)"""
import glob
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
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity.json"
    from test_gpu_swin import _weights          # same weight construction as the tests
    from tools import synth
    dev = torch.device("cuda:0")
    rows = []
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "swin_*.npz"))):
        g = np.load(path)
        wts, _ = _weights(g, dev)
        shape = tuple(int(v) for v in g["shape"])
        x = synth.clip_input(shape, int(g["xseed"])).to(dev)
        feat, score = wts.forward(x, want_feat=True)
        torch.cuda.synchronize()
        st = int(g["feat_stride"])
        fs = (feat[:, ::st] if st > 1 else feat).cpu().numpy()
        ferr = np.abs(fs - g["feat"])
        serr = np.abs(score.cpu().numpy().reshape(-1) - g["score"].reshape(-1))
        rows.append({"fixture": os.path.basename(path)[:-4], "shape": list(shape),
                     "score_ref": [float(v) for v in g["score"].reshape(-1)],
                     "score_err": [float(v) for v in serr], "score_err_max": float(serr.max()), "score_tol": 1e-3,
                     "feat_err_max": float(ferr.max()), "feat_err_mean": float(ferr.mean()),
                     "feat_absmean_ref": float(g["feat_absmean"])})
        print(rows[-1]["fixture"], "score err max %.2e" % serr.max(), "feat max %.2e mean %.2e" % (ferr.max(), ferr.mean()))
    with open(out, "w") as f:
        json.dump({"kernel_path": "kvq_swin3d_forward (C-ABI), attention variant " + os.environ.get("KVQ_ATTN_VARIANT", "default"),
                   "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
