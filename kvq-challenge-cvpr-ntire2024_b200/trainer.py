"""Inference-side drop-in for the reference Trainer (trainer.py:36-74, :298-334).

`test.py` in the reference calls `trainer.inferece()`, which `trainer.Trainer` never defines; the
intended behaviour is `inferece_test`: per video -> reshape [b,c,T,h,w] into clips -> model(inputs=data,
reduce_scores=True) -> mean over clips -> `output.txt` lines `video_name,score`.  This class provides `inferece`
with those semantics on the B200 path.  One process drives ONE GPU; under torchrun the videos are sharded across
ranks and rank 0 writes the file (the reference wraps the model in nn.DataParallel instead, trainer.py:61)."""
import os

import torch

import datasets
from kvq_b200 import ops, parallel
from models import VQA_Network


def split_clips(x, num_clips):
    """[b,c,T,h,w] -> [b*num_clips, c, T/num_clips, h, w]  (trainer.py:308-319)."""
    b, c, t, h, w = x.shape
    return (x.reshape(b, c, num_clips, t // num_clips, h, w).permute(0, 2, 1, 3, 4, 5)
            .reshape(b * num_clips, c, t // num_clips, h, w))


def strip_module_prefix(state_dict):
    """Checkpoints are saved from the DataParallel / DDP wrapper (trainer.py:223-230): keys carry `module.`."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}


def score_video(model, data, key_list, device=None):
    """Body of the inferece_test loop (trainer.py:306-329) for one loader item; returns the video's mean score."""
    if "frames" in data and "technical" not in data:                    # raw frames: the view kernels on the GPU
        fo = data["fragment_opts"]
        frames = data["frames"].to(device, non_blocking=True)
        if frames.dim() == 4:
            frames, offsets = frames[None], data["offsets"][None]
        else:
            offsets = data["offsets"]
        g = lambda v: int(v[0]) if torch.is_tensor(v) else int(v)         # DataLoader collates ints into tensors
        frames = frames.contiguous()
        data["technical"] = ops.fragment_gather_u8(frames, offsets.to(device).int().contiguous(),
                                                   g(fo["fragments_h"]), g(fo["fragments_w"]), g(fo["fsize"]),
                                                   g(fo["aligned"]))
        if "KSVQE" in key_list:
            # the two views of ViewDecompositionDataset_KVQ (fusion_datasets.py:1017-1027): the normalised fragment grid
            # and the frames resized to size_h x size_w, CLIP-normalised -- both from the same decoded frames
            from datasets import views
            ro = data["resize_opts"]
            data["fragment"] = data["technical"]
            aa = ro.get("antialias")                     # None -> datasets.views.RESIZE_ANTIALIAS (torchvision >= 0.17)
            if torch.is_tensor(aa):
                aa = bool(aa.reshape(-1)[0])
            data["resize_video"] = views.resized_video_normalised(frames, g(ro["size_h"]), g(ro["size_w"]), antialias=aa)
    if "KSVQE" in key_list and device is not None:
        # nn.DataParallel scatters these in the reference (trainer.py:61); 'KSVQE' is not a key of the data dict, so
        # the clip reshape below never applies to them (the 96 frames of a video go through as one clip)
        for k in ("fragment", "resize_video", "dis_label"):
            if k in data and torch.is_tensor(data[k]):
                data[k] = data[k].to(device, non_blocking=True)
    for key in key_list:
        if key in data:
            if device is not None:
                data[key] = data[key].to(device)
            nc = data["num_clips"][key]
            nc = int(nc[0]) if torch.is_tensor(nc) else int(nc)
            data[key] = split_clips(data[key], nc).contiguous()
            if key == "simpleVQA" and "feat" in data and torch.is_tensor(data["feat"]):
                # the SlowFast features ride with the frames: one device copy, split into the same clips (the reference
                # head flattens them with view(-1, 2304), head.py:24, so any clip split of [b, T, 2304] is equivalent)
                f = data["feat"].to(device if device is not None else data[key].device, torch.float32)
                data["feat"] = f.reshape(data[key].shape[0], data[key].shape[2], f.shape[-1]).contiguous()
    with torch.no_grad():
        pred = model(inputs=data, reduce_scores=True)
        if isinstance(pred, tuple):            # KSVQE key returns (scores, dis_contra_loss)  (trainer.py:323-325)
            pred = pred[0]
    return pred.float().mean(0).reshape(-1)[0]


class Trainer:
    def __init__(self, args, config):
        self.args, self.config = args, config
        self.gpu_list = [int(i) for i in str(args.gpu_id).split(",")]
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if "LOCAL_RANK" in os.environ:
            # one process per GPU under torchrun: with fewer --gpu_id entries than local ranks (test.py defaults to
            # "0") every rank would land on the same device, so fall back to cuda:LOCAL_RANK
            lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
            gpu = self.gpu_list[local] if len(self.gpu_list) >= lws else local
            if gpu >= torch.cuda.device_count():
                raise RuntimeError(f"kvq_b200: local rank {local} has no GPU (visible devices: {torch.cuda.device_count()}, "
                                   f"--gpu_id {args.gpu_id})")
        else:
            gpu = self.gpu_list[0]
        self.device = torch.device("cuda", gpu)
        torch.cuda.set_device(self.device)
        self.key_list = ["technical"] if any(k.startswith("swin") for k in config["model"]["args"]) \
            else config["model"]["type"].split(",")
        self.build_datasets()
        self.build_models()

    def build_datasets(self):
        if "val" in self.config["data"]:
            v = self.config["data"]["val"]
            self.val_dataset = getattr(datasets, v["type"])(v["args"], None)

    def build_models(self):
        self.model = VQA_Network(self.config).to(self.device)
        path = self.config.get("load_path")
        if path is not None:
            sd = torch.load(path, map_location="cpu")
            sd = sd.get("state_dict", sd)
            msg = self.model.load_state_dict(strip_module_prefix(sd), strict=False)
            print("load", path, msg)
        self.model.eval()

    def inferece(self, output_path="output.txt"):
        import torch.distributed as dist
        ddp = dist.is_available() and dist.is_initialized()
        world, rank = (dist.get_world_size(), dist.get_rank()) if ddp else (1, 0)
        n = len(self.val_dataset)
        lo, hi = parallel.shard_bounds(n, world, rank)
        loader = torch.utils.data.DataLoader(torch.utils.data.Subset(self.val_dataset, range(lo, hi)), batch_size=1,
                                             num_workers=self.config.get("num_workers", 0), pin_memory=True)
        names, scores = [], []
        for data in loader:
            names.append(data["video_name"][0])
            scores.append(score_video(self.model, data, self.key_list, self.device))
        local = torch.stack(scores) if scores else torch.zeros(0, device=self.device)
        allscores = parallel.all_gather_scores(local, n).cpu().tolist()
        if ddp:
            gathered = [None] * world
            dist.all_gather_object(gathered, names)
            names = [x for part in gathered for x in part]
        if rank == 0:
            with open(output_path, "w") as f:
                for name, s in zip(names, allscores):
                    f.write(f"{name},{s}\n")
        return list(zip(names, allscores))

    inferece_test = inferece

    @staticmethod
    def rescale(pr, gt=None):
        """trainer.py:356-361: z-score the predictions, optionally onto the labels' mean / std."""
        import numpy as np
        pr = np.asarray(pr, dtype=np.float64)
        pr = (pr - np.mean(pr)) / np.std(pr)
        if gt is not None:
            pr = pr * np.std(gt) + np.mean(gt)
        return pr

    def inferece_val(self):
        """trainer.py:251-296: scores of the labelled validation set -> SRCC / PLCC / KRCC / RMSE after `rescale`."""
        import numpy as np
        from scipy.stats import kendalltau, pearsonr, spearmanr
        loader = torch.utils.data.DataLoader(self.val_dataset, batch_size=1, num_workers=self.config.get("num_workers", 0),
                                             pin_memory=True)
        preds, labels = [], []
        for data in loader:
            preds.append(float(score_video(self.model, data, self.key_list, self.device)))
            labels.append(float(data["label"].reshape(-1)[0]))
        labels = np.asarray(labels, dtype=np.float64)
        preds = self.rescale(preds, labels)
        s, p, k = spearmanr(labels, preds)[0], pearsonr(labels, preds)[0], kendalltau(labels, preds)[0]
        r = float(np.sqrt(((labels - preds) ** 2).mean()))
        print("SRCC{}PLCC{}KRCC{}RMSE{}".format(s, p, k, r))
        return {"SRCC": float(s), "PLCC": float(p), "KRCC": float(k), "RMSE": r}
