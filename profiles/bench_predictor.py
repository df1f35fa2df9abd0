#!/usr/bin/env python
"""Whole test loop (`python -m src.main <test config> --test` = AcdcVSRRefineNetPredictor.predict): dataset ->
RefineNet x4 -> per-frame L1 / PSNR / SSIM -> log, on synthetic ACDCSR-shaped cine sequences.  What a user of the
reference actually waits for; the net alone is bench.py.

  python profiles/bench_predictor.py [--sequences 64] [--per-launch 16] [--loader host|device]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT, os.path.join(PKG, "csrc")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=64)
    ap.add_argument("--per-launch", type=int, default=16)
    args = ap.parse_args()
    import build as pvsr_build
    pvsr_build.build()
    from src.data.dataloader import Dataloader, DeviceDataloader
    from src.data.datasets import SyntheticCineDataset
    from src.model.metrics import PSNR, SSIM
    from src.model.nets import RefineNet
    from src.runner.predictors import AcdcVSRRefineNetPredictor
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], upscale_factor=4, num_stages=3,
              update_memory=True, num_updated_frames=6, refine_window_size=5, positional_encoding=True)
    out = {"workload": f"{args.sequences} ACDCSR x4 sequences (30 SR frames 216x252 each), {args.per_launch} per launch, "
                       "L1 + PSNR + SSIM per frame"}
    for name, cls in (("host_loader", Dataloader), ("device_loader", DeviceDataloader)):
        torch.manual_seed(0)
        net = RefineNet(**kw)
        ds = SyntheticCineDataset(type='test', downscale_factor=4, num_sequences=args.sequences, num_phases=30,
                                  lr_size=(54, 63), num_frames=7, num_updated_frames=6)
        pred = AcdcVSRRefineNetPredictor(device=torch.device('cuda:0'), test_dataloader=cls(ds, batch_size=1), net=net,
                                         loss_fns=[torch.nn.L1Loss()], loss_weights=[1.0], metric_fns=[PSNR(), SSIM()],
                                         sequences_per_launch=args.per_launch)
        pred.predict()                       # warm-up: plans, graphs, residency
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        log = pred.predict()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"frames_per_s": round(args.sequences * 30 / dt, 1), "s": round(dt, 3),
                     "log": {k: round(v, 5) for k, v in log.items()}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
