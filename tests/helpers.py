"""Shared test helpers: golden-fixture loading and construction of the drop-in RefineNet."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[len("refinenet_"):-4] for p in glob.glob(os.path.join(GOLDEN, "refinenet_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"refinenet_{name}.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def build_net(kwargs, seed=0):
    """The drop-in module initialised exactly like the reference under the same seed."""
    from src.model.nets import RefineNet
    torch.manual_seed(seed)
    return RefineNet(**kwargs)


def oracle_kwargs(kw):
    return dict(num_stages=kw["num_stages"], num_updated_frames=kw["num_updated_frames"],
                refine_window_size=kw["refine_window_size"], upscale_factor=kw["upscale_factor"],
                positional_encoding=kw.get("positional_encoding", False), memory=kw.get("memory", True),
                num_layers=len(kw["num_features"]))
