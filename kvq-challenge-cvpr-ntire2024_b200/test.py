"""Drop-in for the reference test.py (same flags: -o/--opt, -t/--target_set, --gpu_id): YAML -> Trainer ->
`inferece()` -> output.txt.  Launch with torchrun for one process per GPU."""
import argparse
import os
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from trainer import Trainer  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-o", "--opt", type=str, default=os.path.join(HERE, "config", "kvq_swin_grpb_test.yml"))
    parser.add_argument("-t", "--target_set", type=str, default="val")
    parser.add_argument("--gpu_id", type=str, default="0")
    parser.add_argument("--output", type=str, default="output.txt")
    args = parser.parse_args()
    with open(args.opt) as f:
        opt = yaml.safe_load(f)
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch
        import torch.distributed as dist
        # bind the device BEFORE NCCL initialises: --gpu_id defaults to "0", and every rank on cuda:0 fails with a
        # duplicate-GPU error (Trainer applies the same rule)
        gpus = [int(i) for i in str(args.gpu_id).split(",")]
        local, lws = int(os.environ["LOCAL_RANK"]), int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        torch.cuda.set_device(gpus[local] if len(gpus) >= lws else local)
        dist.init_process_group("nccl")
    trainer = Trainer(args, opt)
    for name, score in trainer.inferece(args.output):
        print(name, score)


if __name__ == "__main__":
    main()
