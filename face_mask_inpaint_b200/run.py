"""python -m face_mask_inpaint_b200.run <reference_script.py> [args...]

Installs the sm_100a drop-ins (patch.install) and then runs one of the reference's entry scripts unchanged with
runpy, from the reference checkout that contains the script (SURVEY.md §8b 'How scripts stay unchanged').

  FMI_OFFLINE=1      no network: torchvision `pretrained=True` loss networks get seeded random weights (offline.py)
  torchrun ... -m face_mask_inpaint_b200.run train_reference_fill.py ...
                     one process per GPU: every optimizer the script creates gets the bucketed NCCL gradient all-reduce
                     (dist.enable_for_scripts), parameters are broadcast from rank 0, only rank 0 writes checkpoints
At exit the number of sm_100a kernel launches of the process is written to stderr."""
from __future__ import annotations

import atexit
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    root = os.path.dirname(script)
    os.environ.setdefault("WANDB_MODE", "disabled")
    from . import _lib, patch
    patch.install(root)
    if os.environ.get("FMI_OFFLINE") == "1":
        from .offline import stub_pretrained
        stub_pretrained()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from . import dist
        dist.enable_for_scripts()
    lib = _lib.load()
    atexit.register(lambda: sys.stderr.write(f"[fmi_b200] {lib.fmi_kernel_launch_count()} sm_100a kernel launches\n"))
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
