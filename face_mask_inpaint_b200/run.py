"""python -m face_mask_inpaint_b200.run <reference_script.py> [args...]

Installs the sm_100a drop-ins (patch.install) and then runs one of the reference's entry scripts unchanged with
runpy, from the reference checkout that contains the script (SURVEY.md §8b 'How scripts stay unchanged')."""
from __future__ import annotations

import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    root = os.path.dirname(script)
    os.environ.setdefault("WANDB_MODE", "disabled")
    from . import patch
    patch.install(root)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
