"""Import the UNMODIFIED reference package from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY.  The reference (elvis-sik/muscle_synergies) imports
matplotlib/seaborn at module top (src/muscle_synergies/analysis.py:18,23,
src/muscle_synergies/vicon_data/user_data.py:31, project/segment.py:8-9); neither is
installed here, so four empty stand-in modules are registered before the import.
Nothing in the loader / segmenter / NMF-wrapper code paths touches them.

/root/reference does not exist on the GPU box: only `oracle/make_golden.py` and the
`reference`-marked CPU tests use this module, and they skip when it is absent.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "muscle_synergies"))


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
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    for p in (os.path.join(REFERENCE_ROOT, "src"), os.path.join(REFERENCE_ROOT, "project")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import muscle_synergies  # type: ignore
    import segment  # type: ignore

    return muscle_synergies, segment
