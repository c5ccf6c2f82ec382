"""One warm step, then one step inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_once.py slowfast [batch]
Kernels are launched eagerly (no CUDA graph) so that every launch is a separate ncu row."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "swin"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name]["batch"]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    host, step, modules = bench.build_workload(name, B, dev, 0)
    for m in modules:
        m.use_cuda_graph = False
    x = [t.to(dev) for t in host(3)]
    with torch.no_grad():
        step(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
