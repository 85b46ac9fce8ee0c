"""Stage the reference's model package for the CPU arm of bench.py -- TEST / MEASUREMENT INFRASTRUCTURE.

``python -m oracle.stage_ref`` (also run by ``__graft_entry__.build()``): where ``/root/reference`` exists (the
build container) its ``core/models`` package -- the hot path's own source files, nothing else -- is copied verbatim
into the git-ignored ``baseline/_ref/core/models`` (the location the bench contract reserves for the reference).  The directory is NOT gpurun-ignored, so it travels to the GPU box
with the working tree, where ``bench.py --impl reference`` (and the ``cpu_baseline`` leg) imports the UNMODIFIED
reference ``TBNModel`` from it through ``oracle/ref_shim.py`` (``TBN_REFERENCE_ROOT``) and times it on the host cores.
Nothing under ``baseline/_ref`` enters the git history and the product package never imports it.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
REFERENCE = "/root/reference"


def stage(reference: str = REFERENCE) -> bool:
    src = os.path.join(reference, "core", "models")
    if not os.path.isfile(os.path.join(src, "model.py")):
        return os.path.isfile(os.path.join(STAGE, "core", "models", "model.py"))  # prebuilt copy (GPU box) or nothing
    dst = os.path.join(STAGE, "core", "models")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    with open(os.path.join(STAGE, "core", "__init__.py"), "w"):
        pass
    with open(os.path.join(STAGE, "README"), "w") as f:
        f.write("verbatim copy of /root/reference/core/models (made by oracle/stage_ref.py); git-ignored, not product source\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not available")
