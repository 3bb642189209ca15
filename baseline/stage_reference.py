#!/usr/bin/env python
"""Stage the UNMODIFIED reference (plus the four one-line Python-3 spellings it needs to import at all) under
baseline/_ref/ so that `bench.py --impl reference` can time the reference's own modules on the GPU box's host cores.

    python baseline/stage_reference.py            # needs /root/reference (build container only)

baseline/_ref/ is git-ignored (never part of the history) but travels with `gpurun`, like a built .so.  The reference is
pure Python on PyTorch (requirements.txt:7), there is nothing to compile or pip-install: the package is a plain copy of
its *.py files.  Patches (SURVEY.md section 8c - syntax / import errors on Python 3.12, none changes a result):
  loss.py                `print prob1`            -> `print(prob1)`
  models/dilated_fcn.py  `cuda(async=True)`       -> `cuda(non_blocking=True)`;  `import drn` -> `from models import drn`
  models/drn.py          `gen.next()`             -> `next(gen)`
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
PATCHES = {
    "loss.py": [("print prob1", "print(prob1)")],
    "models/dilated_fcn.py": [("cuda(async=True)", "cuda(non_blocking=True)"), ("\nimport drn\n", "\nfrom models import drn\n")],
    "models/drn.py": [("gen.next()", "next(gen)")],
}


def stage(force=False):
    if not os.path.isdir(REF):
        return None
    if os.path.isdir(DST) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    keep = ("models", "loss.py", "util.py", "transform.py", "argmyparse.py", "datasets.py", "requirements.txt")
    os.makedirs(DST)
    for name in keep:
        src = os.path.join(REF, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DST, name), ignore=shutil.ignore_patterns("*.pyc", "__pycache__"))
        elif os.path.exists(src):
            shutil.copy(src, os.path.join(DST, name))
    for rel, pairs in PATCHES.items():
        p = os.path.join(DST, rel)
        s = open(p).read()
        for a, b in pairs:
            assert a in s, (rel, a)
            s = s.replace(a, b)
        open(p, "w").write(s)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write("%s with the Python-3 spellings listed in baseline/stage_reference.py\n" % REF)
    return DST


def import_reference():
    """put baseline/_ref at the FRONT of sys.path and return its modules; None when it was never staged.
    The reference's factories default to pretrained=True (model-zoo download): forced off, there is no network."""
    if not os.path.isdir(DST):
        return None
    for mod in ("models", "models.drn", "models.dilated_fcn", "models.fusion", "models.model_util", "loss", "util"):
        sys.modules.pop(mod, None)
    sys.path.insert(0, DST)
    import models.drn as drn
    for name in ("drn_d_22", "drn_d_38"):
        orig = getattr(drn, name)
        if not getattr(orig, "_mcd_nopre", False):
            wrapped = (lambda f: lambda pretrained=False, **kw: f(pretrained=False, **kw))(orig)
            wrapped._mcd_nopre = True
            setattr(drn, name, wrapped)
    import loss
    import models.model_util as model_util
    import util
    return dict(loss=loss, model_util=model_util, util=util, drn=drn)


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
