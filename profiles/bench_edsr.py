#!/usr/bin/env python
"""EDSR x4 (reference configs/{train,test}/edsr_net/exp1_x4.yaml: 32 residual blocks x 256 features, 43 M
parameters) on the RefineNet conv core - SURVEY section 8 f3.  Not the headline bench (bench.py); same timing rules:
CUDA events on the launching stream, >= 3 warm-up steps, 256 MiB L2 flush between steps.

  python profiles/bench_edsr.py [--frames 60] [--steps 10]        (contract-shaped line + CPU leg: bench.py --workload edsr_x4)

inference: one step = `--frames` ACDCSR-shaped LR frames 54x63 -> SR frames 216x252 through the public module call
training : one step = forward + L1 + backward + fused Adam on N = 16 patches of 32x32 (HR 128x128)
FLOPs    : 2 * 9 * C_in * C_out per output pixel of every conv (unpadded channels), backward = 2 x forward minus the
           head conv's data gradient
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT, os.path.join(PKG, "csrc")):
    sys.path.insert(0, p)
import torch  # noqa: E402

KW = dict(in_channels=1, out_channels=1, num_resblocks=32, num_features=256, upscale_factor=4, res_scale=0.1)


def conv_flops(n, h, w, F=256, R=32, factors=(2, 2)):
    px = n * h * w
    fwd = 2 * 9 * (1 * F + (2 * R + 1) * F * F) * px
    hh, ww = h, w
    for r in factors:
        fwd += 2 * 9 * F * F * r * r * n * hh * ww
        hh, ww = hh * r, ww * r
    fwd += 2 * 9 * F * 1 * n * hh * ww
    return fwd


def timed(fn, steps, warmup, flush):
    for _ in range(warmup):
        flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        flush.fill_(1)
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def measure(args):
    """Device-timed EDSR x4 inference + training step; returns the result dict (no CPU leg: the CPU oracle is timed by
    `bench.py --workload edsr_x4`, the only bench that may execute oracle/)."""
    import build as pvsr_build
    pvsr_build.build()
    from pvsr.optim import FusedAdam
    from src.model.nets import EDSRNet
    sys.path.insert(0, ROOT)
    from bench import measured_peaks
    dev = torch.device("cuda", 0)
    peaks = measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"model": "EDSRNet x4, 32 resblocks x 256 features (43.08 M parameters), random init", "dtype": "bf16",
           "data": "synthetic", "peak_tflops": peak, "peak_source": peaks["_source"] + ", sustained bf16"}

    # ---- inference
    torch.manual_seed(0)
    net = EDSRNet(**KW).to(dev).eval()
    net.reuse_output_buffers = True
    g = torch.Generator().manual_seed(1234)
    x_h = torch.randn(args.frames, 1, 54, 63, generator=g).pin_memory()
    x_d = x_h.to(dev)
    y_h = torch.empty(args.frames, 1, 216, 252).pin_memory()
    with torch.no_grad():
        ms = timed(lambda: net(x_d), args.steps, max(args.warmup, 3), flush)

        def e2e():
            y_h.copy_(net(x_h.to(dev, non_blocking=True)), non_blocking=True)
        ms_e2e = timed(e2e, args.steps, 3, flush)
    fl = conv_flops(args.frames, 54, 63)
    out["inference"] = {"frames_per_step": args.frames, "ms_per_step": ms, "frames_per_s": args.frames / ms * 1e3,
                        "e2e_frames_per_s": args.frames / ms_e2e * 1e3, "tflop_per_step": fl / 1e12,
                        "tflops": fl / ms / 1e9, "frac_of_peak": fl / ms / 1e9 / peak,
                        "launches_per_step": 2 * 32 + 7, "ms_per_step_e2e": ms_e2e,
                        "h2d_bytes_per_step": x_h.numel() * 4, "d2h_bytes_per_step": y_h.numel() * 4}
    del net
    torch.cuda.empty_cache()

    # ---- training step
    torch.manual_seed(0)
    net = EDSRNet(**KW).to(dev).train()
    opt = FusedAdam.for_net(net, lr=1e-4)
    x = torch.randn(16, 1, 32, 32, generator=g).to(dev)
    t = torch.randn(16, 1, 128, 128, generator=g).to(dev)
    losses = []

    def step():
        loss, _ = net.engine.loss_and_grads(x, t)
        opt.step()
        losses.append(loss)
    ms = timed(step, args.steps, max(args.warmup, 3), flush)
    fwd = conv_flops(16, 32, 32)
    total = 3 * fwd - 2 * 9 * 256 * 16 * 32 * 32         # dgrad + wgrad of every conv, no dgrad through the head conv
    out["train_step"] = {"batch": 16, "ms_per_step": ms, "images_per_s": 16 / ms * 1e3, "tflop_per_step": total / 1e12,
                         "tflops": total / ms / 1e9, "frac_of_peak": total / ms / 1e9 / peak,
                         "loss_first": float(losses[0]), "loss_last": float(losses[-1])}

    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    out = measure(args)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
