"""Install the UNMODIFIED reference (pure Python, MIT) into git-ignored ``baseline/_ref`` so that the stock arms of
``bench.py`` can drive the reference's own modules on the GPU box (``/root/reference`` does not exist there; the
working-tree copy under ``baseline/_ref`` travels with the snapshot, it is never committed).

    python baseline/install_reference.py [--force]

Recipe (the task's one offline install): ``pip install --no-index --no-build-isolation --no-deps --target
baseline/_ref <copy of /root/reference>``; the copy lives under /tmp because the build writes ``*.egg-info`` into
the source tree and ``/root/reference`` is read-only.  The reference's ``pyproject.toml`` lists only the top-level
package (``packages = ["cinema"]``: upstream is used through an editable install), so the wheel lacks the
``cinema.mae`` sub-package; the recipe completes it the way ``pip install -e`` would expose it, by placing the
sub-package's files from the same source tree next to the installed ones.  ``--no-deps``: timm / omegaconf / monai are not in the
offline wheelhouse; ``baseline/stock.py`` supplies the same ~100 lines of stub modules that
``tests/golden/make_golden.py`` uses.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference")
TARGET = ROOT / "baseline" / "_ref"
SUBPACKAGES = ("mae",)


def installed() -> bool:
    return (TARGET / "cinema" / "mae" / "mae.py").exists()


def install(force: bool = False) -> bool:
    """Returns True when ``baseline/_ref`` holds the reference afterwards."""
    if installed() and not force:
        return True
    if not REF_SRC.exists():
        return False
    if TARGET.exists():
        shutil.rmtree(TARGET)
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference install failed:\n{r.stdout}\n{r.stderr}")
        for sub in SUBPACKAGES:  # what an editable install exposes and the wheel leaves out (see module docstring)
            shutil.copytree(src / "cinema" / sub, TARGET / "cinema" / sub, dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*_test.py"))
    return installed()


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print(f"baseline/_ref: {'installed' if ok else 'unavailable (no /root/reference here)'}")
