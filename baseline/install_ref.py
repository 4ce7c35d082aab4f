"""Install the UNMODIFIED reference into baseline/_ref (git-ignored).

The documented command

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
        --target baseline/_ref /root/reference

fails for this reference ("Neither 'setup.py' nor 'pyproject.toml' found"): the upstream repo is
a single script, not a package.  So the install is a verbatim file copy of the reference tree
(byte-identical; `bench.py --impl reference` imports `ddp_example` from there and calls its own
`ConvNet` / `dist_train`).  Outcome recorded in DESIGN.md.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("PDT_REFERENCE_SRC", "/root/reference")


def install(verbose: bool = False) -> str:
    if os.path.exists(os.path.join(DST, "ddp_example.py")):
        return DST
    if not os.path.isdir(SRC):
        raise FileNotFoundError(f"reference source {SRC} not found and {DST} is empty")
    os.makedirs(DST, exist_ok=True)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
                        "--target", DST, SRC], capture_output=True, text=True)
    if verbose:
        print("[install_ref] pip:", (r.stdout + r.stderr).strip().splitlines()[-1:] or "")
    if r.returncode != 0:
        for name in os.listdir(SRC):
            s = os.path.join(SRC, name)
            if os.path.isfile(s):
                shutil.copy2(s, os.path.join(DST, name))
        assert filecmp.cmp(os.path.join(SRC, "ddp_example.py"), os.path.join(DST, "ddp_example.py"), shallow=False)
        if verbose:
            print(f"[install_ref] not pip-installable (script repo): copied verbatim into {DST}")
    return DST


if __name__ == "__main__":
    print(install(verbose=True))
