"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference modules of the hot path, staged so that
they travel to the GPU box (``oracle/_ref/`` is git-ignored, not gpurun-ignored — like the built
``.so``; the repository history never holds reference sources).

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under ``adaptive_b200/`` imports it.  Users:
``bench.py --impl reference`` (the CPU arm, ``cpu_baseline.kind = "reference"``), the
``gpu_eager_baseline`` leg of ``bench.py`` (the same modules on ``cuda:0``: the bar BASELINE.md
section 1 names) and ``tests/`` (cross-checks of the oracle against the live modules).

The reference is Python: "building" it means copying the three files of the path
(``code_src/models/{adaptive_attention,baseline_attention,model_utils}.py``) next to two empty
package markers of our own — the reference's ``code_src/models/__init__.py`` imports its
optimizer factory and an unfinished experiment, neither on the path.  ``stage()`` is called by
``__graft_entry__.build()`` whenever ``/root/reference`` exists (this container); on the GPU box
the already staged files are used as they are.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("adaptive_attention.py", "baseline_attention.py", "model_utils.py")


def stage(verbose: bool = False) -> bool:
    """Copy the path's three modules from the reference checkout into ``oracle/_ref``.
    Returns True when ``oracle/_ref`` is usable afterwards (freshly staged or staged before)."""
    src_dir = os.path.join(REF_ROOT, "code_src", "models")
    dst_dir = os.path.join(DST, "code_src", "models")
    if not os.path.isdir(src_dir):
        return available()
    os.makedirs(dst_dir, exist_ok=True)
    for pkg in (os.path.join(DST, "code_src"), dst_dir):
        with open(os.path.join(pkg, "__init__.py"), "w") as f:
            f.write("# package marker written by oracle/build_ref.py (not a reference file)\n")
    digests = []
    for name in FILES:
        shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
        with open(os.path.join(dst_dir, name), "rb") as f:
            digests.append("%s  %s" % (hashlib.sha256(f.read()).hexdigest(), name))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write("staged from %s by oracle/build_ref.py (unmodified copies)\n" % src_dir + "\n".join(digests) + "\n")
    if verbose:
        print("oracle/_ref staged:\n  " + "\n  ".join(digests))
    return True


def available() -> bool:
    d = os.path.join(DST, "code_src", "models")
    return all(os.path.exists(os.path.join(d, n)) for n in FILES)


def import_reference():
    """-> (adaptive_attention, baseline_attention) modules of the reference, imported from ``oracle/_ref``.
    Raises ImportError when the staging has not happened (run ``__graft_entry__.build()`` where /root/reference exists)."""
    if not available():
        raise ImportError("oracle/_ref is empty: run `python -c 'import __graft_entry__ as g; g.build()'` in the container that "
                          "has %s" % REF_ROOT)
    if DST not in sys.path:
        sys.path.insert(0, DST)
    warnings.filterwarnings("ignore", message=".*nn.functional.(tanh|sigmoid) is deprecated.*")
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    warnings.filterwarnings("ignore", category=UserWarning, module="torch.nn.functional")
    from code_src.models import adaptive_attention, baseline_attention  # noqa: E402

    return adaptive_attention, baseline_attention


def make_encoder2decoder(ref_mod, cf, identity_trunk: bool = True):
    """``Encoder2Decoder(cf)`` of the reference without the ResNet-152 download (SURVEY 8c workaround 1): torchvision's
    constructor is asked for random weights, and (``identity_trunk``) the trunk is replaced by ``nn.Identity`` so that
    ``images`` are ``[B,2048,7,7]`` feature maps — BASELINE config 1."""
    import torch.nn as nn
    import torchvision.models as tvm

    orig = tvm.resnet152
    tvm.resnet152 = lambda pretrained=True, **kw: orig(weights=None)
    try:
        model = ref_mod.Encoder2Decoder(cf)
    finally:
        tvm.resnet152 = orig
    if identity_trunk:
        model.encoder.resnet_conv = nn.Identity()
    return model


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("oracle/_ref available:", ok)
