"""Stand-alone timing of the HBM-bound kernels against a device-to-device copy on the same box:
python profiles/membound_bench.py  -> one JSON object per kernel (algorithmic bytes / CUDA-event time, fraction of the
copy bandwidth measured in the same process).  L2 is flushed between repetitions."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from pvsr import ops  # noqa: E402


def timed(fn, flush, n=10, warm=3):
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


def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    res = []
    a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    ms = timed(lambda: b.copy_(a), flush)
    copy_gbs = 2 * a.numel() / ms / 1e6
    res.append({"kernel": "torch copy 1 GiB (read + write)", "ms": ms, "GB/s": copy_gbs})
    del a, b

    def add(name, ms, nbytes):
        res.append({"kernel": name, "ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1),
                    "frac_of_copy": round(nbytes / ms / 1e6 / copy_gbs, 3)})

    # inference: head last conv over 30 frames x 32 sequences at 216x252; in_conv over 42 x 32 frames at 54x63
    w = torch.randn(1, 64, 3, 3, device=dev, generator=g) * 0.05
    bias = torch.zeros(1, device=dev)
    x = torch.randn(960, 216, 252, 64, device=dev, generator=g).to(torch.bfloat16)
    add("head_conv_last fwd 960x216x252", timed(lambda: ops.head_conv_last(x, w, bias), flush), x.numel() * 2 + x.numel() // 64 * 4)
    del x
    xi = torch.randn(42 * 32, 54, 63, device=dev, generator=g)
    wi = torch.randn(64, 1, 3, 3, device=dev, generator=g)
    bi = torch.zeros(64, device=dev)
    sl = torch.tensor([0.2], device=dev)
    add("in_conv_prelu 1344x54x63", timed(lambda: ops.in_conv_prelu(xi, wi, bi, sl), flush), xi.numel() * (4 + 128))
    # training: head last conv and its adjoints over one stage (3 lists x 7 frames x 16) at 128x128
    xt = torch.randn(336, 128, 128, 64, device=dev, generator=g).to(torch.bfloat16)
    dout = torch.randn(336, 128, 128, device=dev, generator=g)
    add("head_conv_last fwd 336x128x128", timed(lambda: ops.head_conv_last(xt, w, bias), flush), xt.numel() * 2 + dout.numel() * 4)
    add("head_conv_last bwd data+weight 336x128x128", timed(lambda: ops.head_conv_last_bwd(xt, w, dout), flush),
        2 * xt.numel() * 2 + 2 * dout.numel() * 4)
    for r in res:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
