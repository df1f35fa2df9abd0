"""ORACLE TOOLING - generates tests/golden/drfnet_*.npz by running the UNMODIFIED reference DRFNet
(/root/reference/src/model/nets/drf_net.py) on CPU in the build container:

    python oracle/make_golden_drf.py            (--sisr-only: only the DRFSISRNet cases -> tests/golden/drfsisr_*.npz)

Weights are not stored: the drop-in module reproduces the reference's construction order, so torch.manual_seed(0) +
construction gives the same parameters (per-tensor checksums are stored to prove it).  Stored: the T input frames, the
targets, the T outputs, the frame-averaged L1 loss (acdc_vsr_trainer.py:40-43,83-94), per-parameter gradient
norm / sum, every 1-D gradient (biases, PReLU slopes) and a strided sample of every weight gradient.
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STRIDE = 257

CASES = {
    # name: (kwargs, N, T, h, w)
    "x4_g3": (dict(num_features=64, num_groups=3, upscale_factor=4), 2, 3, 10, 9),
    "x2_g2": (dict(num_features=64, num_groups=2, upscale_factor=2), 2, 3, 9, 12),
    "x3_g2": (dict(num_features=64, num_groups=2, upscale_factor=3), 1, 2, 8, 7),
    "x8_g2": (dict(num_features=64, num_groups=2, upscale_factor=8), 1, 2, 6, 5),
    "x4_g1_f128": (dict(num_features=128, num_groups=1, upscale_factor=4), 1, 2, 8, 8),
}


def load_reference():
    sys.path.insert(0, REF)
    for n, p in [("src", REF + "/src"), ("src.model", REF + "/src/model"), ("src.model.nets", REF + "/src/model/nets")]:
        m = types.ModuleType(n)
        m.__path__ = [p]
        sys.modules[n] = m
    return importlib.import_module("src.model.nets.drf_net")


def make_case(ref, name, kw, N, T, h, w):
    base = dict(in_channels=1, out_channels=1)
    base.update(kw)
    torch.manual_seed(0)
    net = ref.DRFNet(**base)
    s = base["upscale_factor"]
    g = torch.Generator().manual_seed(2468)
    inputs = [torch.randn(N, 1, h, w, generator=g) for _ in range(T)]
    targets = [torch.randn(N, 1, h * s, w * s, generator=g) for _ in range(T)]
    outputs = net(inputs)
    loss = torch.stack([torch.nn.L1Loss()(o, t) for o, t in zip(outputs, targets)]).mean()
    loss.backward()
    rec = {"inputs": np.stack([x.numpy() for x in inputs]), "targets": np.stack([t.numpy() for t in targets]),
           "outputs": np.stack([o.detach().numpy() for o in outputs]), "loss": np.float64(loss.item())}
    meta = {"kwargs": base, "N": N, "T": T, "h": h, "w": w, "stride": STRIDE, "params": {}, "grads": {}}
    for k, p in net.named_parameters():
        meta["params"][k] = [list(p.shape), float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
        meta["grads"][k] = [float(p.grad.double().norm()), float(p.grad.double().sum())]
        rec["grad::" + k] = p.grad.numpy() if p.grad.dim() == 1 else p.grad.reshape(-1)[::STRIDE].numpy()
    rec["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, f"drfnet_{name}.npz"), **rec)
    print(name, "loss", loss.item(), "out sum", float(sum(o.sum() for o in outputs)),
          "params", sum(p.numel() for p in net.parameters()))


SISR_CASES = {
    # DRFSISRNet (drf_sisr_net.py): name: (kwargs, N, h, w)
    "x4_g2_s3": (dict(num_steps=3, num_features=64, num_groups=2, upscale_factor=4), 2, 9, 10),
    "x2_g1_s2": (dict(num_steps=2, num_features=64, num_groups=1, upscale_factor=2), 1, 8, 7),
}


def make_sisr_case(name, kw, N, h, w):
    """DRFSISRNet.forward (drf_sisr_net.py:38-49) + the SRFB trainer's loss (acdc_sisr_srfb_trainer.py:22-26)."""
    ref = importlib.import_module("src.model.nets.drf_sisr_net")
    base = dict(in_channels=1, out_channels=1)
    base.update(kw)
    torch.manual_seed(0)
    net = ref.DRFSISRNet(**base)
    s = base["upscale_factor"]
    g = torch.Generator().manual_seed(1357)
    x = torch.randn(N, 1, h, w, generator=g)
    target = torch.randn(N, 1, h * s, w * s, generator=g)
    outputs = net(x)
    loss = torch.stack([torch.nn.L1Loss()(o, target) for o in outputs]).mean()
    loss.backward()
    rec = {"input": x.numpy(), "target": target.numpy(), "outputs": np.stack([o.detach().numpy() for o in outputs]),
           "loss": np.float64(loss.item())}
    meta = {"kwargs": base, "N": N, "h": h, "w": w, "stride": STRIDE, "params": {}, "grads": {}}
    for k, p in net.named_parameters():
        meta["params"][k] = [list(p.shape), float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
        meta["grads"][k] = [float(p.grad.double().norm()), float(p.grad.double().sum())]
        rec["grad::" + k] = p.grad.numpy() if p.grad.dim() == 1 else p.grad.reshape(-1)[::STRIDE].numpy()
    rec["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, f"drfsisr_{name}.npz"), **rec)
    print("sisr", name, "loss", loss.item(), "steps", len(outputs))


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = load_reference()
    if "--sisr-only" not in sys.argv:
        for name, (kw, N, T, h, w) in CASES.items():
            make_case(ref, name, kw, N, T, h, w)
    for name, (kw, N, h, w) in SISR_CASES.items():
        make_sisr_case(name, kw, N, h, w)
