"""ORACLE TOOLING - generates tests/golden/*.npz by running the UNMODIFIED reference model
(/root/reference/src/model/nets/refine_net.py) on CPU in the build container.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures together with this
script.  Loading recipe: namespace-stub packages skip src/__init__.py (which imports
nibabel etc.), and ConvLSTMCell.init_hidden's hard-coded `.cuda()` (refine_net.py:269-271) is replaced by
"same device as the conv weight" - the only two deviations from the reference source.

Weights are NOT stored (11.6 MB): the drop-in module reproduces the reference's initialisation order, so
`torch.manual_seed(seed)` + construction gives the same parameters; per-tensor checksums are stored to prove it.
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


def load_reference():
    sys.path.insert(0, REF)
    for n, p in [("src", REF + "/src"), ("src.model", REF + "/src/model"), ("src.model.nets", REF + "/src/model/nets")]:
        m = types.ModuleType(n)
        m.__path__ = [p]
        sys.modules[n] = m
    ref = importlib.import_module("src.model.nets.refine_net")

    def init_hidden(self, b, h, w):
        z = lambda: torch.zeros(b, self.hidden_dim, h, w, device=self.conv.weight.device, dtype=self.conv.weight.dtype)
        return (z(), z())

    ref.ConvLSTMCell.init_hidden = init_hidden
    return ref


def trainer_loss(outputs, targets):
    """acdc_vsr_refinenet_trainer.py:83-93 with nn.L1Loss."""
    l1 = torch.nn.L1Loss()
    loss = []
    for i, outs in enumerate(outputs):
        discount = np.power(0.5, (len(outputs) // 3 - i // 3 - 1))
        loss.append(torch.stack([l1(o, t) * discount for o, t in zip(outs, targets)]).mean())
    return torch.stack(loss).sum()


CASES = {
    # name: (net kwargs, N, T, h, w)
    "x4_pos": (dict(upscale_factor=4, positional_encoding=True), 2, 3, 8, 8),
    "x3_pos": (dict(upscale_factor=3, positional_encoding=True), 2, 2, 8, 8),
    "x2_pos": (dict(upscale_factor=2, positional_encoding=True), 2, 2, 8, 8),
    "x4_nopos": (dict(upscale_factor=4, positional_encoding=False), 1, 2, 8, 8),
    "x4_nomem": (dict(upscale_factor=4, positional_encoding=True, memory=False), 1, 2, 8, 8),
    "x4_rect": (dict(upscale_factor=4, positional_encoding=True), 1, 2, 7, 10),
    "x8_pos": (dict(upscale_factor=8, positional_encoding=True), 1, 1, 5, 6),
    "x4_2stage_2layer": (dict(upscale_factor=4, positional_encoding=True, num_stages=2, num_features=[64, 64]), 1, 2, 6, 9),
}
U = 3


def make_case(ref, name, kw, N, T, h, w):
    base = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
                num_updated_frames=U, refine_window_size=5)
    base.update(kw)
    torch.manual_seed(0)
    net = ref.RefineNet(**base)
    s = base["upscale_factor"]
    L = T + 2 * U
    g = torch.Generator().manual_seed(1234)
    inputs = [torch.randn(N, 1, h, w, generator=g) for _ in range(L)]
    pos = torch.randn(N, L, 1, generator=g)
    targets = [torch.randn(N, 1, h * s, w * s, generator=g) for _ in range(T)]

    net.eval()
    with torch.no_grad():
        out_eval = net(inputs, pos)
    net.train()
    out_train = net(inputs, pos)
    loss = trainer_loss(out_train, targets)
    loss.backward()

    rec = {
        "inputs": torch.stack(inputs).numpy(), "pos": pos.numpy(), "targets": torch.stack(targets).numpy(),
        "outputs_eval": torch.stack([torch.stack(o) for o in out_eval]).numpy(),
        "loss": np.float64(loss.item()),
    }
    meta = {"kwargs": base, "N": N, "T": T, "h": h, "w": w, "U": U, "params": {}, "grads": {}}
    for k, p in net.named_parameters():
        meta["params"][k] = [list(p.shape), float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
        if p.grad is None:
            meta["grads"][k] = None
        else:
            meta["grads"][k] = [float(p.grad.double().norm()), float(p.grad.double().sum())]
            if p.grad.numel() <= 2048:
                rec["grad::" + k] = p.grad.numpy()
    rec["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, f"refinenet_{name}.npz"), **rec)
    print(name, "loss", loss.item(), "out sum", float(out_eval[-1][0].sum()))


def known_answers(ref):
    """The KAT smoke values of SURVEY.md section 8c, regenerated (x4/x3/x2/x4-no-pos at N=2, L=19, 8x8)."""
    kat = {}
    for name, kw in [("x4", dict(upscale_factor=4, positional_encoding=True)),
                     ("x3", dict(upscale_factor=3, positional_encoding=True)),
                     ("x2", dict(upscale_factor=2, positional_encoding=True)),
                     ("x4_nopos", dict(upscale_factor=4, positional_encoding=False))]:
        torch.manual_seed(0)
        net = ref.RefineNet(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
                            num_updated_frames=6, refine_window_size=5, **kw)
        s = kw["upscale_factor"]
        g = torch.Generator().manual_seed(1234)
        inputs = [torch.randn(2, 1, 8, 8, generator=g) for _ in range(19)]
        pos = torch.randn(2, 19, 1, generator=g)
        targets = [torch.randn(2, 1, 8 * s, 8 * s, generator=g) for _ in range(7)]
        net.train()
        out = net(inputs, pos)
        loss = trainer_loss(out, targets)
        loss.backward()
        gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in net.parameters() if p.grad is not None)))
        kat[name] = {"out_sum": float(out[-1][0].sum()), "out_first4": [float(v) for v in out[-1][0][0, 0, 0, :4]],
                     "loss": float(loss), "grad_l2": gn,
                     "n_params": int(sum(p.numel() for p in net.parameters()))}
        print("KAT", name, kat[name])
    with open(os.path.join(OUT, "known_answers.json"), "w") as f:
        json.dump(kat, f, indent=1)


def metrics_golden():
    """Known answers of the reference's own metrics.py / utils.py on seeded images (inputs are regenerated from the
    seeds by the tests, only the expected values are stored)."""
    metrics = importlib.import_module("src.model.metrics")
    utils = importlib.import_module("src.utils")
    import pickle
    import tempfile
    out = []
    for seed, (n, h, w) in enumerate([(1, 216, 252), (2, 40, 37), (1, 11, 11), (3, 64, 48)]):
        g = torch.Generator().manual_seed(100 + seed)
        a = torch.randn(n, 1, h, w, generator=g)
        b = a + 0.1 * torch.randn(n, 1, h, w, generator=g)
        rec = {"seed": 100 + seed, "shape": [n, 1, h, w]}
        # cardiac bounding box (h0, hn, w0, wn) of the one patient these frames belong to (metrics.py:116-165 reads it
        # from a pickle keyed by patient name); at least the 11-pixel SSIM window in both directions
        box = [h // 5, max(h // 5 + 11, (4 * h) // 5), w // 4, max(w // 4 + 11, (3 * w) // 4)] if min(h, w) > 11 \
            else [0, h, 0, w]
        rec["cardiac_box"] = box
        with tempfile.TemporaryDirectory() as td:
            cp = os.path.join(td, "coordinates.pkl")
            with open(cp, "wb") as f:
                pickle.dump({"patient_g": tuple(box)}, f)
            cpsnr = metrics.CardiacPSNR(coordinates_path=cp, size_average=False)
            cssim = metrics.CardiacSSIM(coordinates_path=cp, size_average=False)
            cpsnr_avg, cssim_avg = metrics.CardiacPSNR(coordinates_path=cp), metrics.CardiacSSIM(coordinates_path=cp)
        for ds in ("acdc", "dsb15"):
            da, db = utils.denormalize(a, ds), utils.denormalize(b, ds)
            rec[ds] = {"denorm_sum": float(da.double().sum()), "psnr": float(metrics.PSNR()(da, db)),
                       "ssim": float(metrics.SSIM()(da, db)),
                       "psnr_per_sample": [float(v) for v in metrics.PSNR(size_average=False)(da, db)],
                       "ssim_per_sample": [float(v) for v in metrics.SSIM(size_average=False)(da, db)],
                       "cardiac_psnr": float(cpsnr_avg(da, db, "patient_g")),
                       "cardiac_ssim": float(cssim_avg(da, db, "patient_g")),
                       "cardiac_psnr_per_sample": [float(v) for v in cpsnr(da, db, "patient_g")],
                       "cardiac_ssim_per_sample": [float(v) for v in cssim(da, db, "patient_g")]}
        out.append(rec)
    with open(os.path.join(OUT, "metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("metrics golden written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = load_reference()
    if "--metrics" not in sys.argv:        # --metrics: only rewrite metrics.json
        for name, (kw, N, T, h, w) in CASES.items():
            make_case(ref, name, kw, N, T, h, w)
        known_answers(ref)
    metrics_golden()
