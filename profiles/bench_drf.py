#!/usr/bin/env python
"""DRFNet x4 (reference src/model/nets/drf_net.py; no reference config instantiates it - 64 features, 6 projection
groups as in the SRFBN it derives from) on the RefineNet conv core - SURVEY section 8 f3.  Not the headline bench
(bench.py); same timing rules: CUDA events on the launching stream, >= 3 warm-up steps, 256 MiB L2 flush between steps.

  python profiles/bench_drf.py [--seqs 8] [--frames 30] [--steps 5]   (contract-shaped line + CPU leg: bench.py --workload drfnet_x4)

inference: one step = `--seqs` ACDCSR-shaped cine sequences of `--frames` LR frames 54x63 -> SR frames 216x252 through
           the public module call (the frame recurrence is sequential; the `--seqs` sequences are batched)
training : one step = forward + frame-averaged L1 + BPTT backward + fused Adam on N = 16 sequences of 7 frames 32x32
FLOPs    : ALGORITHMIC, as the reference evaluates the net: 2 * k*k * C_in * C_out per output pixel of a Conv2d and per
           INPUT pixel of a ConvTranspose2d; `executed` counts the phase-stacked 3x3 forms (9 * P instead of k*k taps).
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

KW = dict(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=4)


def conv_flops(frames, h, w, F=64, G=6, s=4, executed=False):
    """Forward FLOPs of `frames` LR frames h x w."""
    k, P = s + 4, s * s
    proj = 9 * P if executed else k * k
    per_lr = 9 * 4 * F + 4 * F * F + 2 * F * F                     # in_block conv1 / conv2, f_block.in_block
    for i in range(G):
        if i > 0:
            per_lr += (i + 1) * F * F + (i + 1) * F * F * P           # 1x1 convs over the LR / HR concatenations
        per_lr += 2 * proj * F * F                                    # deconv + strided conv
    per_lr += G * F * F                                               # f_block.out_block
    per_lr += 9 * F * 4 * F + 4 * 9 * F * 4 * F + 16 * 9 * F         # _OutBlock x4: two shuffle convs + the last conv
    return 2.0 * per_lr * frames * h * w


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
    import build as pvsr_build
    pvsr_build.build()
    from pvsr.optim import FusedAdam
    from src.model.nets import DRFNet
    sys.path.insert(0, ROOT)
    from bench import measured_peaks
    dev = torch.device("cuda", 0)
    peaks = measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"model": "DRFNet x4, 64 features, 6 projection groups (3.66 M parameters), random init", "dtype": "bf16",
           "data": "synthetic", "peak_tflops": peak, "peak_source": peaks["_source"] + ", sustained bf16"}
    T, n = args.frames, args.seqs

    # ---- inference
    torch.manual_seed(0)
    net = DRFNet(**KW).to(dev).eval()
    net.reuse_output_buffers = True
    g = torch.Generator().manual_seed(1234)
    x_h = torch.randn(T, n, 1, 54, 63, generator=g).pin_memory()
    x_d = [x.to(dev) for x in x_h]
    y_h = torch.empty(T, n, 1, 216, 252).pin_memory()
    with torch.no_grad():
        ms = timed(lambda: net(x_d), args.steps, max(args.warmup, 3), flush)

        def e2e():
            xs = x_h.to(dev, non_blocking=True)
            net([xs[t] for t in range(T)])           # reuse_output_buffers: the T outputs are views of one [T, n, 1, H, W] buffer
            y_h.copy_(net.engine.geoms[(T, n, 54, 63, False)].out, non_blocking=True)
        ms_e2e = timed(e2e, args.steps, 3, flush)
    fl, fx = conv_flops(T * n, 54, 63), conv_flops(T * n, 54, 63, executed=True)
    launches = 2 + T * (2 + 4 * 6 - 2 + 1) + 3 + 2
    out["inference"] = {"sequences_per_step": n, "frames_per_step": T * n, "ms_per_step": ms,
                        "frames_per_s": T * n / ms * 1e3, "e2e_frames_per_s": T * n / ms_e2e * 1e3,
                        "tflop_per_step": fl / 1e12, "executed_tflop_per_step": fx / 1e12, "tflops": fl / ms / 1e9,
                        "executed_tflops": fx / ms / 1e9, "frac_of_peak": fl / ms / 1e9 / peak,
                        "executed_frac_of_peak": fx / ms / 1e9 / peak, "launches_per_step": launches,
                        "ms_per_step_e2e": ms_e2e, "h2d_bytes_per_step": x_h.numel() * 4,
                        "d2h_bytes_per_step": y_h.numel() * 4}
    del net
    torch.cuda.empty_cache()

    # ---- training step
    torch.manual_seed(0)
    net = DRFNet(**KW).to(dev).train()
    opt = FusedAdam.for_net(net, lr=1e-4)
    Tt, N = 7, 16
    xs = [torch.randn(N, 1, 32, 32, generator=g).to(dev) for _ in range(Tt)]
    ts = [torch.randn(N, 1, 128, 128, generator=g).to(dev) for _ in range(Tt)]
    losses = []

    def step():
        loss, _ = net.engine.loss_and_grads(xs, ts)
        opt.step()
        losses.append(loss)
    ms = timed(step, args.steps, max(args.warmup, 3), flush)
    total = 3 * conv_flops(Tt * N, 32, 32) - 2 * 9 * 4 * 64 * Tt * N * 32 * 32     # no data gradient through in_block.conv1
    out["train_step"] = {"batch": N, "frames": Tt, "ms_per_step": ms, "target_frames_per_s": Tt * N / ms * 1e3,
                         "tflop_per_step": total / 1e12, "tflops": total / ms / 1e9,
                         "frac_of_peak": total / ms / 1e9 / peak, "loss_first": float(losses[0]),
                         "loss_last": float(losses[-1])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=8)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    print(json.dumps(measure(args)), flush=True)


if __name__ == "__main__":
    main()
