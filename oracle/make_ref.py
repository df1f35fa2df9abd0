"""ORACLE TOOLING - recipe that stages the UNMODIFIED reference model under oracle/_ref/ (git-ignored, shipped to
the GPU box with the snapshot like the built .so files).

    python oracle/make_ref.py            # needs /root/reference (the build container); no-op elsewhere

The reference is pure Python and not a package (no setup.py / pyproject.toml), so "building" it means placing the
two files its hot path consists of -

    src/model/nets/refine_net.py   (RefineNet and its blocks)
    src/model/nets/base_net.py     (its nn.Module base class)
    src/model/metrics.py, src/utils.py (PSNR / SSIM / Cardiac* / denormalize - torch + numpy + pickle only)

- where `oracle/ref_model.py` can import them without the reference's `src/__init__.py` (which pulls nibabel,
SimpleITK, python-box ...: none installed here).  Nothing is edited; oracle/_ref/MANIFEST.json records the sha256 of
every staged file so a run can state which reference bytes it timed.  Reference sources never enter the git history
(`oracle/_ref/` is in .gitignore); only `tests/`, `__graft_entry__` and bench.py's CPU legs use the staged copy.
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["src/model/nets/refine_net.py", "src/model/nets/base_net.py", "src/model/metrics.py", "src/utils.py"]


def make_ref(verbose=False):
    """Returns the manifest dict, or None when /root/reference is absent and nothing was staged before."""
    manifest_path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isdir(REF):
        if os.path.exists(manifest_path):
            with open(manifest_path) as f:
                return json.load(f)
        return None
    manifest = {"source": REF, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(json.dumps(manifest, indent=1))
    return manifest


if __name__ == "__main__":
    m = make_ref(verbose=True)
    sys.exit(0 if m else 1)
