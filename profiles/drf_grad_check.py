#!/usr/bin/env python
"""Per-parameter gradient error of the DRFNet golden fixtures (debug aid): python profiles/drf_grad_check.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
sys.path[:0] = [PKG, ROOT, os.path.join(ROOT, "tests"), os.path.join(PKG, "csrc")]
import torch  # noqa: E402
import test_drf as TD  # noqa: E402

for name in (sys.argv[1:] or TD.CASES):
    z, meta = TD._load(name)
    net = TD._net(meta["kwargs"]).cuda().train()
    inputs = [torch.from_numpy(x).cuda() for x in z["inputs"]]
    targets = [torch.from_numpy(x).cuda() for x in z["targets"]]
    net.zero_grad()
    outs = net(inputs)
    loss = torch.stack([torch.nn.L1Loss()(o, t) for o, t in zip(outs, targets)]).mean()
    loss.backward()
    print(name, "loss", loss.item(), float(z["loss"]))
    for k, p in net.named_parameters():
        g = p.grad.detach().float().cpu()
        norm, _ = meta["grads"][k]
        want = torch.from_numpy(z["grad::" + k])
        got = g if g.dim() == 1 else g.reshape(-1)[::meta["stride"]]
        rel = ((got - want).norm() / want.norm()).item()
        flag = "  <<<" if (rel > 0.12 or abs(float(g.double().norm()) / norm - 1) > 0.12) else ""
        print(f"  {k:45s} n={want.numel():6d} norm {float(g.double().norm()):.4e} / {norm:.4e}  rel {rel:.4f}{flag}")
    from oracle import drf_oracle as O
    kw = meta["kwargs"]
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    terms = O.slope_gradient_terms(sd, [x.cpu() for x in inputs], [t.cpu() for t in targets], kw["num_groups"], kw["upscale_factor"])
    for k, (tot, ab) in terms.items():
        got = float(dict(net.named_parameters())[k].grad)
        print(f"  slope {k:42s} ref {tot:+.4e} got {got:+.4e} abs-sum {ab:.4e}  err/abs-sum {abs(got - tot) / ab:.5f}")
