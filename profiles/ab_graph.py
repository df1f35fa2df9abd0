"""A/B timing of the captured forward / backward graphs in ONE process (same box, same clocks):
python profiles/ab_graph.py [--env PVSR_PDL=0,1] - prints ms per replay of the training forward, training backward
and inference forward graphs for each setting (CUDA events, 20 replays after 5 warm-ups, L2 flush between replays)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from pvsr import lib as L  # noqa: E402
from pvsr.synthetic import cine_batch  # noqa: E402
from src.model.nets import RefineNet  # noqa: E402

KW = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], upscale_factor=4, num_stages=3,
          update_memory=True, num_updated_frames=6, refine_window_size=5, positional_encoding=True)


def timed(fn, flush, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def measure(dev, flush, do_infer=True):
    res = {}
    torch.manual_seed(0)
    net = RefineNet(**KW).to(dev).train()
    eng = net.engine
    inputs, pos, targets = cine_batch(16, T=7, U=6, h=32, w=32, scale=4, seed=4321, end_systole=3, with_targets=True)
    xs, ps, ts = [x.to(dev) for x in inputs], pos.to(dev), [t.to(dev) for t in targets]
    eng.loss_and_grads(xs, ps, ts)
    pl = eng.train_plan(xs)
    grads = {k: p.grad for k, p in net.named_parameters()}
    res["train_fwd_ms"] = timed(lambda: eng.run(pl), flush)
    res["train_bwd_ms"] = timed(lambda: eng.backward(pl, grads), flush)
    res["train_step_ms"] = timed(lambda: eng.loss_and_grads(xs, ps, ts), flush)
    if do_infer:
        net2 = RefineNet(**KW).to(dev).eval()
        net2.only_last_head = True
        e2 = net2.engine
        inputs, pos = cine_batch(32, T=30, U=6, h=54, w=63, scale=4, seed=1234)
        pl2 = e2.plan_for(32, len(inputs), 54, 63, False, dev)
        with torch.no_grad():
            e2.stage_inputs(pl2, [x.to(dev) for x in inputs], pos.to(dev))
            res["infer_fwd_ms"] = timed(lambda: e2.run(pl2), flush, n=8, warm=3)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--switch", default="pdl", choices=["pdl", "halo", "cta_pair"])
    ap.add_argument("--values", default="1,0,1,0")
    ap.add_argument("--no-infer", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = L.load()
    setter = {"pdl": lib.pvsr_set_pdl, "halo": lib.pvsr_set_halo_mode, "cta_pair": lib.pvsr_set_cta_pair}[a.switch]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for v in a.values.split(","):
        setter(int(v))
        r = measure(dev, flush, not a.no_infer)
        print(a.switch, v, {k: round(x, 3) for k, x in r.items()}, flush=True)


if __name__ == "__main__":
    main()
