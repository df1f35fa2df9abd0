"""ORACLE TOOLING - generates tests/golden/edsr_*.npz by running the UNMODIFIED reference EDSRNet
(/root/reference/src/model/nets/edsr_net.py) on CPU in the build container:

    python oracle/make_golden_edsr.py

Weights are not stored: the drop-in module reproduces the reference's construction order, so torch.manual_seed(0) +
construction gives the same parameters (per-tensor checksums are stored to prove it).  Stored: input, target, output,
L1 loss, per-parameter gradient norm / sum, every bias gradient and a strided sample of every weight gradient.
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
    # name: (kwargs, N, h, w)
    "x4_f64_r2": (dict(num_resblocks=2, num_features=64, upscale_factor=4), 2, 12, 10),
    "x2_f128_r2": (dict(num_resblocks=2, num_features=128, upscale_factor=2), 2, 9, 11),
    "x3_f64_r1": (dict(num_resblocks=1, num_features=64, upscale_factor=3), 1, 8, 8),
    "x4_f256_r3": (dict(num_resblocks=3, num_features=256, upscale_factor=4, res_scale=0.1), 2, 8, 8),
    "x3_f256_r1": (dict(num_resblocks=1, num_features=256, upscale_factor=3), 1, 6, 7),
}


def load_reference():
    sys.path.insert(0, REF)
    for n, p in [("src", REF + "/src"), ("src.model", REF + "/src/model"), ("src.model.nets", REF + "/src/model/nets")]:
        m = types.ModuleType(n)
        m.__path__ = [p]
        sys.modules[n] = m
    return importlib.import_module("src.model.nets.edsr_net")


def make_case(ref, name, kw, N, h, w):
    base = dict(in_channels=1, out_channels=1, res_scale=0.1)
    base.update(kw)
    torch.manual_seed(0)
    net = ref.EDSRNet(**base)
    s = base["upscale_factor"]
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(N, 1, h, w, generator=g)
    target = torch.randn(N, 1, h * s, w * s, generator=g)
    out = net(x)
    loss = torch.nn.L1Loss()(out, target)
    loss.backward()
    rec = {"input": x.numpy(), "target": target.numpy(), "output": out.detach().numpy(), "loss": np.float64(loss.item())}
    meta = {"kwargs": base, "N": N, "h": h, "w": w, "stride": STRIDE, "params": {}, "grads": {}}
    for k, p in net.named_parameters():
        meta["params"][k] = [list(p.shape), float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
        meta["grads"][k] = [float(p.grad.double().norm()), float(p.grad.double().sum())]
        rec["grad::" + k] = p.grad.numpy() if p.grad.dim() == 1 else p.grad.reshape(-1)[::STRIDE].numpy()
    rec["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, f"edsr_{name}.npz"), **rec)
    print(name, "loss", loss.item(), "out sum", float(out.sum()), "params", sum(p.numel() for p in net.parameters()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = load_reference()
    for name, (kw, N, h, w) in CASES.items():
        make_case(ref, name, kw, N, h, w)
