#!/usr/bin/env python
"""bench.py -- clips/s of the KSVQE Swin3D-T forward (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3                 # this repo's sm_100a path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference algorithm on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                     # one rank per GPU, clips sharded (weak scaling)

A step = one forward of VQA_Network('swin_tiny_grpb') over a batch of 8 synthetic clips [8,3,32,224,224] per GPU
through the public API (models.VQA_Network.forward(inputs=..., reduce_scores=True)), followed (N > 1) by the one
collective of the path: an NCCL all-gather of the per-clip scores (trainer_ddp.py:262).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

CLIP = (3, 32, 224, 224)
WORKLOAD = "KSVQE Swin3D-T forward (swin_tiny_grpb backbone + VQAHead), batch 8 clips 32x3x224x224 per GPU"
METRIC = "clips/sec KSVQE Swin3D 32x224x224"
MODEL_CFG = {"model": {"type": "swin", "args": {"swin_tiny_grpb": {"head": {"in_channels": 768, "hidden_channels": 64}}}}}
# algorithmic work per clip, SURVEY.md section 8(d) / 8(a6-a8) (FLOP = 2*MAC)
ATTN_GFLOP_PER_CLIP_BLOCK = (7.55, 3.78, 1.89, 0.94)        # QK^T + PV per block, stages 0..3
SWIN_GFLOP_PER_CLIP = 175.53


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def synth_model_state(seed=0):
    from oracle import synth
    return synth.swin_network_state_dict(seed, key="swin_tiny_grpb")


def oracle_clips_per_sec(sd, n_clips, warm=1, want_scores=False):
    """The reference algorithm restated on CPU (oracle/swin3d.py, pinned to the reference by tests/golden),
    all host threads, one 32x224x224 clip per call."""
    import torch
    from oracle import swin3d, synth
    torch.set_num_threads(os.cpu_count() or 1)
    scores = []
    with torch.no_grad():
        for i in range(warm):
            swin3d.vqa_network_swin(sd, synth.clip_input((1,) + CLIP, 1000 + i))
        t0 = time.perf_counter()
        for i in range(n_clips):
            x = synth.clip_input((1,) + CLIP, 3 if i == 0 else 2000 + i)[:1]
            scores.append(float(swin3d.vqa_network_swin(sd, x).reshape(-1)[0]))
        dt = time.perf_counter() - t0
    return n_clips / dt, torch.get_num_threads(), scores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = synth_model_state()
    per_step = 1                                   # bounded sample: one clip of the 8-clip batch per step
    import torch
    from oracle import swin3d, synth
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        for i in range(args.warmup):
            swin3d.vqa_network_swin(sd, synth.clip_input((per_step,) + CLIP, 1000 + i))
        t0 = time.perf_counter()
        for i in range(args.steps):
            swin3d.vqa_network_swin(sd, synth.clip_input((per_step,) + CLIP, 2000 + i))
        dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_step": "1 clip sample of the 8-clip batch (CPU)"},
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} clips of 32x3x224x224, one per step, oracle/swin3d.py fp32 on "
                                       f"{cores} host threads (the Python reference cannot travel to the GPU box)"},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from kvq_b200 import lib
    import models
    from oracle import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    L = lib.load()
    B = args.batch

    model = models.VQA_Network(MODEL_CFG)
    sd = synth_model_state()
    model.load_state_dict(sd, strict=False)
    model = model.to(dev)
    model.eval()
    model.use_cuda_graph = not args.no_graph      # one graph per (input buffer, shape): ~95 launches -> 1 replay

    x = synth.clip_input((B,) + CLIP, 3 + rank).to(dev)          # 154 MB > 126 MB L2
    gathered = torch.empty(world * B, dtype=torch.float32, device=dev)

    def step(inp):
        s = model(inputs={"technical": inp}, reduce_scores=True).reshape(-1)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)             # the one collective of the path (trainer_ddp.py:262)
            return gathered
        return s

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step(x)
        sync_all()
        scores_gpu = out[rank * B:(rank + 1) * B].clone() if world > 1 else out.clone()

        # ---------------- device-resident throughput ----------------
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        from kvq_b200 import ops as kops
        launches0 = kops.kernel_launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(args.steps):
            step(x)
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        launches = kops.kernel_launches() - launches0
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        value = world * B * args.steps / (ms / 1000.0)

        # ---------------- end to end: pinned host clips -> H2D -> forward -> scores D2H ----------------
        xh = [synth.clip_input((B,) + CLIP, 50 + rank).pin_memory(), synth.clip_input((B,) + CLIP, 60 + rank).pin_memory()]
        xd = [torch.empty_like(x), torch.empty_like(x)]
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                xd[i % 2].copy_(xh[i % 2], non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_loop(n):
            for e in consumed:
                e.record()
            upload(0)
            res = None
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)                              # prefetch the next batch while this one computes
                torch.cuda.current_stream().wait_event(ready[i % 2])
                s = step(xd[i % 2])
                consumed[i % 2].record()
                res = s.cpu()                                  # D2H read of the step's result (sync)
            return res

        e2e_loop(2)
        sync_all()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        sync_all()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * B * args.steps / float(t.item())

        # ---------------- per-kernel breakdown (instrumented pass, outside the timed regions) ----------------
        roof, breakdown = None, None
        if rank == 0:
            ncat = L.kvq_profile_num_categories()
            psteps = min(args.steps, 5)
            model.use_cuda_graph = False              # events are recorded between the kernels: eager launches
            L.kvq_profile_enable(1)
            for _ in range(psteps):
                model(inputs={"technical": x}, reduce_scores=True)
            import ctypes
            msb = (ctypes.c_float * ncat)()
            cnt = (ctypes.c_int * ncat)()
            lib.check(min(L.kvq_profile_collect(msb, cnt, ncat), 0), "profile_collect")
            L.kvq_profile_enable(0)
            total = sum(msb)
            rows = [(L.kvq_profile_category_name(i).decode(), msb[i], cnt[i]) for i in range(ncat) if cnt[i] > 0]
            rows.sort(key=lambda r: -r[1])
            breakdown = [{"kernel": n, "ms_per_step": m / psteps, "launches_per_step": c // psteps,
                          "share": m / total} for n, m, c in rows]
            pk = peaks()
            # dominant kernel = the fused window attention; report the heaviest stage instance
            top = next(r for r in rows if r[0].startswith("window_attn"))
            stage = int(top[0][-1])
            avg_ms = top[1] / top[2]
            flops = ATTN_GFLOP_PER_CLIP_BLOCK[stage] * 1e9 * B
            achieved = flops / (avg_ms * 1e-3) / 1e12
            roof = {"kernel": f"window_attn2_kernel ({top[0]})", "bound": "tensor", "achieved": achieved,
                    "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": args.traffic,
                    "avg_launch_ms": avg_ms, "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json)",
                    "algorithmic": f"{ATTN_GFLOP_PER_CLIP_BLOCK[stage]} GFLOP/clip/block (QK^T+PV) x {B} clips per launch",
                    "step_tensor_frac": SWIN_GFLOP_PER_CLIP * 1e9 * value / world / 1e12 / pk["tflops"]}

    # ---------------- CPU baseline (rank 0, N = 1) + score delta ----------------
    cpu = None
    score_delta = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sc = oracle_clips_per_sec(sd, args.cpu_clips, warm=1, want_scores=True)
        cpu = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_clips} clips of 32x3x224x224 (1 warm-up), oracle/swin3d.py fp32, one clip per call"}
        score_delta = abs(sc[0] - float(scores_gpu[0].item()))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"clip-sharded x{world}",
                           "l2": "inputs (154 MB/GPU) and the ~0.9 GB activation workspace exceed the 126 MB L2",
                           "arithmetic": "fp16 operands, fp32 accumulate / softmax / LayerNorm / residual",
                           "launch": "eager" if args.no_graph else "CUDA graph replay per input buffer"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * 3 * 32 * 224 * 224 * 4,
                        "d2h_bytes_per_step": (world if world > 1 else 1) * B * 4},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "score_delta_vs_oracle": score_delta, "breakdown": breakdown}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step")
    ap.add_argument("--cpu-clips", type=int, default=4, help="clips timed by the CPU baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels eagerly instead of CUDA-graph replay")
    ap.add_argument("--traffic", type=float, default=295.4e6,
                    help="dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (stage-0 "
                         "window_attn_fast_kernel at batch 8) from the ncu --set full capture in profiles/r01_summary.md")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
