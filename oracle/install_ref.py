"""Installs the UNMODIFIED reference (elvis-sik/muscle_synergies) into baseline/_ref/ so that bench.py --impl reference
can time the reference's own CPU implementation on the GPU box, where /root/reference does not exist.

    python oracle/install_ref.py            # here, in the build container

baseline/_ref/ is git-ignored (no reference source enters the history) and not gpurun-ignored (it travels with the
snapshot).  What is installed: the `muscle_synergies` package by pip from a scratch copy of the checkout (the build
writes into its source tree; /root/reference is read-only), plus project/segment.py - the windowing script the
package itself does not ship.  Dependencies are not resolved (--no-deps): the pinned versions (pandas < 2,
scikit-learn <= 0.24) are not installable here; matplotlib / seaborn, imported at module top by the reference,
are stood in for by oracle/refstub.py.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.environ.get("MS_REFERENCE_SOURCE", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "muscle_synergies", "__init__.py")) and os.path.isfile(
        os.path.join(TARGET, "segment.py"))


def install(force: bool = False) -> str:
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "src", "muscle_synergies")):
        raise RuntimeError(f"reference checkout not found under {SOURCE}")
    with tempfile.TemporaryDirectory(prefix="ms_ref_") as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, copy, symlinks=True)
        shutil.rmtree(TARGET, ignore_errors=True)
        os.makedirs(TARGET, exist_ok=True)
        subprocess.check_call(
            [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
             "--find-links", "/opt/wheelhouse", "--target", TARGET, copy])
        shutil.copy(os.path.join(copy, "project", "segment.py"), os.path.join(TARGET, "segment.py"))
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
