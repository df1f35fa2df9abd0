#!/usr/bin/env python
"""Whole training loop (`python -m src.main <train config>` = AcdcVSRRefineNetTrainer._run_epoch('training')):
dataset -> batch -> fused RefineNet x4 step (forward + multi-stage L1 + backward + Adam) -> PSNR / SSIM -> log, at the
reference's training shapes (N = 16, 7 target frames + 2 x 6 warm-up frames, 32x32 LR patches).  The step alone is
bench.py --mode train.

  python profiles/bench_trainer.py [--steps 60]
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
    ap.add_argument("--steps", type=int, default=60)
    args = ap.parse_args()
    import build as pvsr_build
    pvsr_build.build()
    from pvsr.optim import FusedAdam
    from src.callbacks.monitor import Monitor
    from src.data.dataloader import Dataloader, DeviceDataloader
    from src.data.datasets import SyntheticCineDataset
    from src.model.metrics import PSNR, SSIM
    from src.model.nets import RefineNet
    from src.runner.trainers import AcdcVSRRefineNetTrainer
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], upscale_factor=4, num_stages=3,
              update_memory=True, num_updated_frames=6, refine_window_size=5, positional_encoding=True)
    dev = torch.device('cuda:0')
    n_seq = max(2, (args.steps + 2) * 16 // 30 + 1)
    out = {"workload": f"{args.steps} training steps, N=16, 19 LR frames 32x32 -> 7 HR targets 128x128, fused L1 + FusedAdam, "
                       "PSNR + SSIM per step"}
    import tempfile
    for name, cls, kwargs in (("host_loader_workers8", Dataloader, dict(num_workers=8, pin_memory=True)),
                              ("device_loader", DeviceDataloader, {})):
        torch.manual_seed(0)
        net = RefineNet(**kw).to(dev)
        opt = FusedAdam.for_net(net, lr=1e-4)
        ds = SyntheticCineDataset(type='train', downscale_factor=4, num_sequences=n_seq, num_phases=30,
                                  lr_size=(32, 32), num_frames=7, num_updated_frames=6)
        dl = cls(ds, batch_size=16, shuffle=True, drop_last=True, **kwargs)
        with tempfile.TemporaryDirectory() as tmp:
            tr = AcdcVSRRefineNetTrainer(device=dev, train_dataloader=dl, valid_dataloader=dl, net=net,
                                         loss_fns=[torch.nn.L1Loss()], loss_weights=[1.0], metric_fns=[PSNR(), SSIM()],
                                         optimizer=opt, lr_scheduler=None, logger=None,
                                         monitor=Monitor(tmp, 'min', 'Loss', 1000), num_epochs=1)
            tr._run_epoch('training')          # warm-up epoch: plans, graphs, worker start-up, residency
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            log, _, _ = tr._run_epoch('training')
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[name] = {"steps": len(dl), "ms_per_step": round(1e3 * dt / len(dl), 2),
                     "target_frames_per_s": round(len(dl) * 16 * 7 / dt, 1), "log": {k: round(v, 4) for k, v in log.items()}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
