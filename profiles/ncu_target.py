"""Profiling target: exactly two eager (no CUDA graph) passes of the bench workload, so that an ncu launch list of
this command shows every kernel of a step by name and `profiles/ncu_capture.sh` can pick representative launches
for the `--set full` captures.  python profiles/ncu_target.py infer|train|edsr_infer|edsr_train|drf_infer|drf_train [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from pvsr.synthetic import cine_batch  # noqa: E402
from src.model.nets import RefineNet  # noqa: E402

KW = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], upscale_factor=4, num_stages=3,
          update_memory=True, num_updated_frames=6, refine_window_size=5, positional_encoding=True)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if mode == "infer":
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
        net = RefineNet(**KW).to(dev).eval()
        net.only_last_head = True
        eng = net.engine
        eng.use_graph = False
        inputs, pos = cine_batch(B, T=30, U=6, h=54, w=63, scale=4, seed=1234)
        pl = eng.plan_for(B, len(inputs), 54, 63, False, dev)
        with torch.no_grad():
            eng.stage_inputs(pl, [x.to(dev) for x in inputs], pos.to(dev))
            for _ in range(2):
                eng.run(pl)
                torch.cuda.synchronize()
    elif mode == "loader":
        # HBM-resident input pipeline: two fetches of 32 whole ACDCSR x4 cycles (pvsr_cine_gather x 2 each)
        from src.data.dataloader import DeviceDataloader
        from src.data.datasets import SyntheticCineDataset
        ds = SyntheticCineDataset(type='test', downscale_factor=4, num_sequences=32, num_phases=30, lr_size=(54, 63),
                                  num_frames=7, num_updated_frames=6)
        dl = DeviceDataloader(ds, batch_size=1)
        for _ in range(2):
            dl.fetch(list(range(32)))
            torch.cuda.synchronize()
    elif mode.startswith("drf"):
        # DRFNet x4, 64 features x 6 groups: 16 ACDCSR-shaped sequences, 3 frames of the recurrence
        from src.model.nets import DRFNet
        net = DRFNet(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=4).to(dev)
        net.engine.use_graph = False
        g = torch.Generator().manual_seed(1234)
        if mode == "drf_infer":
            xs = [torch.randn(16, 1, 54, 63, generator=g).to(dev) for _ in range(3)]
            net.eval()
            with torch.no_grad():
                for _ in range(2):
                    net(xs)
                    torch.cuda.synchronize()
        else:
            net.train()
            xs = [torch.randn(16, 1, 32, 32, generator=g).to(dev) for _ in range(3)]
            ts = [torch.randn(16, 1, 128, 128, generator=g).to(dev) for _ in range(3)]
            for _ in range(2):
                net.engine.loss_and_grads(xs, ts)
                torch.cuda.synchronize()
    elif mode.startswith("edsr"):
        # EDSR x4, 32 blocks x 256 features (configs/{train,test}/edsr_net/exp1_x4.yaml) - SURVEY 8 f3
        from src.model.nets import EDSRNet
        net = EDSRNet(in_channels=1, out_channels=1, num_resblocks=32, num_features=256, upscale_factor=4,
                      res_scale=0.1).to(dev)
        net.engine.use_graph = False
        g = torch.Generator().manual_seed(1234)
        if mode == "edsr_infer":
            n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
            x = torch.randn(n, 1, 54, 63, generator=g).to(dev)
            net.eval()
            with torch.no_grad():
                for _ in range(2):
                    net(x)
                    torch.cuda.synchronize()
        else:
            from pvsr.optim import FusedAdam
            net.train()
            opt = FusedAdam.for_net(net, lr=1e-4)
            x = torch.randn(16, 1, 32, 32, generator=g).to(dev)
            t = torch.randn(16, 1, 128, 128, generator=g).to(dev)
            for _ in range(2):
                net.engine.loss_and_grads(x, t)
                opt.step()
                torch.cuda.synchronize()
    else:
        N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
        from pvsr.optim import FusedAdam
        net = RefineNet(**KW).to(dev).train()
        opt = FusedAdam.for_net(net, lr=1e-4)
        eng = net.engine
        eng.use_graph = False
        inputs, pos, targets = cine_batch(N, T=7, U=6, h=32, w=32, scale=4, seed=4321, end_systole=3,
                                          with_targets=True)
        xs, ps, ts = [x.to(dev) for x in inputs], pos.to(dev), [t.to(dev) for t in targets]
        for _ in range(2):
            eng.loss_and_grads(xs, ps, ts)
            opt.step()
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
