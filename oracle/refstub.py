"""Import the UNMODIFIED reference package from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY.  The reference (elvis-sik/muscle_synergies) imports
matplotlib/seaborn at module top (src/muscle_synergies/analysis.py:18,23,
src/muscle_synergies/vicon_data/user_data.py:31, project/segment.py:8-9); neither is
installed here, so four empty stand-in modules are registered before the import.
Nothing in the loader / segmenter / NMF-wrapper code paths touches them.

/root/reference does not exist on the GPU box: `oracle/make_golden.py` and the `reference`-marked CPU
tests use this module here and skip when it is absent; `bench.py --impl reference` (and the cpu_baseline
leg) use the pip-installed copy under baseline/_ref/ (oracle/install_ref.py), which travels.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MS_REFERENCE_ROOT", "/root/reference")
# the pip-installed copy that travels to the GPU box (oracle/install_ref.py): package and segment.py side by side
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _paths():
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "muscle_synergies")):
        return [os.path.join(REFERENCE_ROOT, "src"), os.path.join(REFERENCE_ROOT, "project")]
    if os.path.isfile(os.path.join(INSTALLED_ROOT, "muscle_synergies", "__init__.py")):
        return [INSTALLED_ROOT]
    return []


def reference_available() -> bool:
    return bool(_paths())


def _install_stubs():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_ms_stub", False):
        return  # a real matplotlib is importable; leave it alone
    try:
        import matplotlib  # noqa: F401
        import seaborn  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl._ms_stub = True
    plt = types.ModuleType("matplotlib.pyplot")
    patches = types.ModuleType("matplotlib.patches")
    sns = types.ModuleType("seaborn")

    class _Style:
        @staticmethod
        def use(*_a, **_k):
            return None

    plt.style = _Style()
    plt.Figure = object
    plt.Axes = object
    patches.Rectangle = object
    mpl.pyplot = plt
    mpl.patches = patches
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("matplotlib.patches", patches)
    sys.modules.setdefault("seaborn", sns)


def import_reference():
    """Returns (muscle_synergies, segment) modules of the reference, unmodified."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT} or {INSTALLED_ROOT}")
    _install_stubs()
    for p in _paths():
        if p not in sys.path:
            sys.path.insert(0, p)
    import muscle_synergies  # type: ignore
    import segment  # type: ignore

    return muscle_synergies, segment
