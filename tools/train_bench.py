#!/usr/bin/env python
"""BASELINE.json config 5, hot-path share: the data-parallel TRAINING step of the hot path at 256x256 crops,
8 images per GPU (batch 64 on 8 GPUs) -- forward + backward of the 4 masked window attentions, 6 GDN/IGDN and the
rounding of one AutoEncoderRGB_Journal step, then the gradient exchange (the path's only collective).

    python tools/train_bench.py [--steps 20]                                       1 GPU (no exchange)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py

Two exchanges are timed: the hot-path parameters themselves (what this repo owns: ~0.55 M floats) and, for scale,
a flat buffer of the full model's 34.07 M fp32 gradients (136 MB) through the same bucketed
`GradientAllReduce`.  One JSON line on rank 0; device time, max over ranks.  Development / evidence tool, not the
driver's bench (that is bench.py at the repo root).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mwa_b200 as pkg  # noqa: E402

CROP, PER_GPU = 256, 8
FULL_MODEL_PARAMS = 34_070_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    bench.IMG_H = bench.IMG_W = CROP                       # same call sites, 256x256 crops
    ops = bench.to_device(bench.build_workload(pkg, dev, PER_GPU, seed0=PER_GPU * rank), dev)
    params = [p for op in ops if "mod" in op for p in op["mod"].parameters()]
    for op in ops:
        if op["kind"] in ("attn", "gdn"):
            op["x"].requires_grad_(True)
            op["gy"] = torch.randn_like(op["x"])
        else:
            for k in ("y", "mu", "lrp", "z"):
                op[k].requires_grad_(True)
    ar = pkg.GradientAllReduce(params, clip_value=5.0)
    full = [torch.nn.Parameter(torch.zeros(FULL_MODEL_PARAMS, device=dev))]
    full[0].grad = torch.randn(FULL_MODEL_PARAMS, device=dev)
    ar_full = pkg.GradientAllReduce(full, clip_value=5.0)

    def fwd_bwd():
        for op in ops:
            if op["kind"] == "attn":
                op["mod"](op["x"], op["alpha"]).backward(op["gy"])
            elif op["kind"] == "gdn":
                op["mod"](op["x"]).backward(op["gy"])
            else:
                outs = bench.run_op(pkg, op)
                torch.autograd.backward([o for o in outs if o.requires_grad],
                                        [torch.ones_like(o) for o in outs if o.requires_grad])
        for t in [op[k] for op in ops for k in ("x", "y", "mu", "lrp", "z") if k in op]:
            t.grad = None

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # per-op forward + backward times (rank 0, informational)
    per_op = {}
    for op in ops:
        if op["kind"] == "round":
            continue
        def one(op=op):
            y = op["mod"](op["x"], op["alpha"]) if op["kind"] == "attn" else op["mod"](op["x"])
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); y.backward(op["gy"]); e1.record(); torch.cuda.synchronize()
            op["x"].grad = None
            return e0.elapsed_time(e1)
        one(); one()
        per_op[op["name"] + " bwd"] = round(min(one() for _ in range(3)), 3)
    ms_fb = timed(fwd_bwd)
    ms_ar = timed(ar)
    ms_full = timed(ar_full)
    if rank == 0:
        step_ms = ms_fb + ms_ar
        print(json.dumps({
            "metric": "256x256 RGBA crops/s, hot-path training step (fwd + bwd + gradient exchange + clip)",
            "value": PER_GPU * world / (step_ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "config": {"workload": "C5-hotpath: 4 masked window attention + 6 GDN/IGDN + rounding, forward and "
                                   "backward, 8 crops of 256x256 per GPU", "global_batch": PER_GPU * world},
            "fwd_bwd_ms": ms_fb, "hot_path_grad_exchange_ms": ms_ar, "hot_path_grad_bytes": ar.bytes_per_step,
            "per_op_backward_ms": per_op, "full_model_grad_exchange_ms": ms_full, "full_model_grad_bytes": 4 * FULL_MODEL_PARAMS,
            "full_model_busbw_gbs": (2 * (world - 1) / world * 4 * FULL_MODEL_PARAMS / (ms_full * 1e-3) / 1e9)
            if world > 1 else None,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
