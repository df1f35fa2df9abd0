"""ORACLE - TEST INFRASTRUCTURE ONLY.  Loader for the UNMODIFIED reference model staged under oracle/_ref/ by
oracle/make_ref.py (never read from /root/reference at run time: that path does not exist on the GPU box).

Two deviations from running the reference as-is, both from SURVEY.md section 8c's loading recipe:
  * `src/__init__.py` is bypassed (it imports nibabel / SimpleITK / box, none installed): the two model files are
    executed under temporary stub packages, and the product's own `src` package is put back afterwards;
  * `ConvLSTMCell.init_hidden` allocates its zero state on the conv weight's device instead of the hard-coded
    `.cuda()` of refine_net.py:269-271, so the model runs on the host CPU.
Used by bench.py's CPU legs (`cpu_baseline.kind = "reference"`) and by tests that cross-check the port
(oracle/refinenet_oracle.py) against the live reference where it is available.
"""
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "_ref")
_cache = {}


def available():
    return os.path.exists(os.path.join(ROOT, "src", "model", "nets", "refine_net.py"))


def manifest():
    try:
        with open(os.path.join(ROOT, "MANIFEST.json")) as f:
            return json.load(f)
    except Exception:
        return None


def _exec(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference's `refine_net` module (and `metrics`, `utils`) or raises
    FileNotFoundError when oracle/_ref/ was not staged."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise FileNotFoundError("oracle/_ref is empty: run `python oracle/make_ref.py` where /root/reference exists")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    try:
        for n, p in (("src", "src"), ("src.model", "src/model"), ("src.model.nets", "src/model/nets")):
            m = types.ModuleType(n)
            m.__path__ = [os.path.join(ROOT, p)]
            sys.modules[n] = m
        _exec("src.model.nets.base_net", os.path.join(ROOT, "src/model/nets/base_net.py"))
        refine = _exec("src.model.nets.refine_net", os.path.join(ROOT, "src/model/nets/refine_net.py"))
        metrics = _exec("src.model.metrics", os.path.join(ROOT, "src/model/metrics.py"))
        utils = _exec("src.utils", os.path.join(ROOT, "src/utils.py"))
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)

    def init_hidden(self, b, h, w):
        z = lambda: torch.zeros(b, self.hidden_dim, h, w, device=self.conv.weight.device, dtype=self.conv.weight.dtype)
        return (z(), z())

    refine.ConvLSTMCell.init_hidden = init_hidden
    ns = types.SimpleNamespace(refine_net=refine, metrics=metrics, utils=utils, RefineNet=refine.RefineNet)
    _cache["ns"] = ns
    return ns


def build_net(state_dict=None, seed=0, **kwargs):
    """The reference RefineNet on the CPU (eval mode); `state_dict` (26 reference keys) or seeded default init."""
    ns = load()
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
              num_updated_frames=6, refine_window_size=5, upscale_factor=4, positional_encoding=True)
    kw.update(kwargs)
    torch.manual_seed(seed)
    net = ns.RefineNet(**kw)
    if state_dict is not None:
        net.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()}, strict=True)
    return net.eval()
