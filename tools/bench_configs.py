#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5: the front-end feeding its two consumers on the device.

  config 4  predict.py path: normalise -> centred STFT -> |z|^2 -> mel (160, 513, 1) -> badwinner2 inference
  config 5  training input pipeline: normalise -> raw_to_mel image (160, 513, 3) -> PCEN -> wr_resnet_bird training step
            (forward + backward + SGD), one process per GPU (DDP over NCCL when launched with torchrun)

Reports front-end-only, model-only and end-to-end clips/s per configuration and whether the front-end keeps the model
fed (front-end clips/s / model clips/s).  The models are torch restatements with random weights
(audio-training_b200/consumers.py, pinned against the executed reference graph by tests/test_consumers.py); bf16 autocast, channels_last -- the consumer's precision is the
consumer's choice, the features stay FP32.  CUDA-event timed, max over ranks.

  python tools/bench_configs.py [--config 4|5|all] [--batch N] [--steps K]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --config 5
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt            # noqa: E402
from audio_training_b200 import consumers as cs           # noqa: E402


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def synth(batch, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(144000, device=device, dtype=torch.float32) / 48000.0
    f0 = 300.0 + 10700.0 * torch.rand((batch, 1), generator=g, device=device)
    return (0.3 * (2 * torch.rand((batch, 144000), generator=g, device=device) - 1) +
            0.4 * torch.sin(2 * torch.pi * f0 * t)).contiguous()


def config4(args, dev, world):
    plan = rt.Plan(rt.FrontendConfig(framing="center_zero", power=2, channels=1, normalize=True), dev)
    model = cs.build_model((160, 513, 1), None, args.labels).cuda(dev).eval().to(memory_format=torch.channels_last)
    x = synth(args.batch, f"cuda:{dev}", 4 + dev)
    feat = plan.frontend(x)

    def run_model(f):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(f)

    fe = timed(lambda: plan.frontend(x), args.steps)
    mo = timed(lambda: run_model(feat), args.steps)
    e2e = timed(lambda: run_model(plan.frontend(x)), args.steps)
    k = args.batch * world * 1e3
    return {"config": "4: path B front-end -> badwinner2 inference (predict.py path)", "n_gpus": world, "batch_per_gpu": args.batch,
            "frontend_clips_per_s": k / fe, "model_clips_per_s": k / mo, "end_to_end_clips_per_s": k / e2e,
            "frontend_over_model": mo / fe, "model": "BadWinner2 torch restatement, bf16 autocast, random weights",
            "frontend_share_of_step": fe / e2e}


def config5(args, dev, world):
    plan = rt.Plan(rt.FrontendConfig(framing="tf_pad_end", power=2, channels=3, normalize=True), dev)
    model = cs.WRResNet((160, 513, 3), args.labels).cuda(dev).to(memory_format=torch.channels_last)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev])
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    x = synth(args.batch, f"cuda:{dev}", 5 + dev)
    y = (torch.rand((args.batch, args.labels), device=f"cuda:{dev}") > 0.8).float()
    lossf = torch.nn.BCELoss()

    def features():
        return plan.pcen(plan.frontend(x), time_axis=2)      # image [B, M, T, 3]: EMA along T per mel and channel

    feat = features()

    def train(f):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(f)
        loss = lossf(out.float(), y)
        loss.backward()
        opt.step()

    fe = timed(features, args.steps)
    mo = timed(lambda: train(feat), args.steps)
    e2e = timed(lambda: train(features()), args.steps)
    k = args.batch * world * 1e3
    return {"config": "5: raw_to_mel (C=3) -> PCEN -> wr_resnet_bird training step", "n_gpus": world, "batch_per_gpu": args.batch,
            "frontend_clips_per_s": k / fe, "model_clips_per_s": k / mo, "end_to_end_clips_per_s": k / e2e,
            "frontend_over_model": mo / fe, "model": "WRResNetBird torch restatement, bf16 autocast, SGD, "
            + ("DDP/NCCL" if world > 1 else "single process"), "frontend_share_of_step": fe / e2e}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--labels", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    for name, fn in (("4", config4), ("5", config5)):
        if args.config in ("all", name):
            res = fn(args, dev, world)
            if dev == 0:
                print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
