"""Recipe for oracle/_ref: the reference's own hot-path modules, UNMODIFIED, where bench.py's CPU baseline legs can run them.

The reference is pure Python; its tree (/root/reference) exists only in the build container. To time the reference's real
step loop (performance_benchmark.py:106-133) on the GPU box's host cores in the same run as the CUDA path, this copies the
seven files the loop executes -- environments/{__init__,base,chemical_reactor,power_grid,robot_assembly}.py and
core/{__init__,types}.py -- byte for byte into oracle/_ref/neorl_industrial/ (git-ignored: never in history; NOT
gpurun-ignored: it travels to the box like the built .so files) together with a manifest of their SHA-256 sums.
`__graft_entry__.build()` calls build(); on the GPU box (no /root/reference) it is a no-op and the copy that travelled is used.

    python oracle/make_ref.py            # (re)make oracle/_ref and print the manifest
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("NIG_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = ["environments/__init__.py", "environments/base.py", "environments/chemical_reactor.py",
         "environments/power_grid.py", "environments/robot_assembly.py", "core/__init__.py", "core/types.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def manifest():
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as f:
            return json.load(f)
    except Exception:
        return None


def available() -> bool:
    m = manifest()
    return bool(m) and all(os.path.isfile(os.path.join(DEST, "neorl_industrial", f)) for f in FILES)


def verify() -> bool:
    """The copy is byte-identical to what the manifest recorded (and to the reference tree, when that is present)."""
    m = manifest()
    if not m:
        return False
    for f in FILES:
        p = os.path.join(DEST, "neorl_industrial", f)
        if not os.path.isfile(p) or _sha(p) != m["sha256"][f]:
            return False
        src = os.path.join(REF_ROOT, "src", "neorl_industrial", f)
        if os.path.isfile(src) and _sha(src) != m["sha256"][f]:
            return False
    return True


def build(force: bool = False) -> bool:
    src_pkg = os.path.join(REF_ROOT, "src", "neorl_industrial")
    if not os.path.isdir(src_pkg):
        return available()                       # GPU box: use what travelled
    if available() and verify() and not force:
        return True
    shutil.rmtree(DEST, ignore_errors=True)
    sums = {}
    for f in FILES:
        dst = os.path.join(DEST, "neorl_industrial", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_pkg, f), dst)
        sums[f] = _sha(dst)
    version = None
    try:
        with open(os.path.join(src_pkg, "_version.py")) as f:
            for line in f:
                if line.strip().startswith("__version__ = version ="):
                    version = line.split("=")[-1].strip().strip("'\"")
    except OSError:
        pass
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src_pkg, "reference_version": version, "sha256": sums,
                   "note": "byte-for-byte copies made by oracle/make_ref.py; git-ignored"}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print(json.dumps({"available": ok, "verified": verify(), "manifest": manifest()}, indent=1))
