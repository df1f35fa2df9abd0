#!/usr/bin/env python
"""Gradient error of one golden fixture against the CPU oracle, per parameter (debug aid): python profiles/grad_check.py x4_nopos"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
sys.path[:0] = [PKG, ROOT, os.path.join(ROOT, "tests"), os.path.join(PKG, "csrc")]
import torch  # noqa: E402
from helpers import build_net, load_golden, oracle_kwargs  # noqa: E402
from oracle import refinenet_oracle as O  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "x4_nopos"
z, meta = load_golden(name)
kw = meta["kwargs"]
net = build_net(kw)
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
inputs = [torch.from_numpy(x) for x in z["inputs"]]
pos = torch.from_numpy(z["pos"])
targets = [torch.from_numpy(t) for t in z["targets"]]
params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
out = O.refinenet_forward(params, inputs, pos, train=True, **oracle_kwargs(kw))
O.trainer_loss(out, targets, training=True).backward()
net = net.cuda().train()
o = net([x.cuda() for x in inputs], pos.cuda())
O.trainer_loss(o, [t.cuda() for t in targets], training=True).backward()
torch.cuda.synchronize()
rows = []
for k, p in net.named_parameters():
    r = params[k].grad
    if p.grad is None or r is None:
        continue
    g = p.grad.cpu()
    rel = float((g - r).norm() / r.norm())
    cos = float((g * r).sum() / (g.norm() * r.norm()))
    rows.append((rel, cos, k))
rows.sort(reverse=True)
print(" ".join(f"{k}={os.environ[k]}" for k in os.environ if k.startswith("PVSR_")) or "defaults")
for rel, cos, k in rows[:5]:
    print(f"  {rel:.4f} {cos:.6f} {k}")
