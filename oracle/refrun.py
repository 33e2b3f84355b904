"""Run the UNMODIFIED reference package under stub dependencies (oracle; test infrastructure only).

``load()`` imports ``/root/reference/dl_segmentation_utils`` exactly as it lies on disk, with ``oracle/refstubs`` standing
in for tensorflow / rasterio / descarteslabs / geopandas / osgeo (none installable here; see refstubs/README.md), and
returns the package module — or ``None`` where ``/root/reference`` does not exist (the GPU box).  It is how the
fixtures under ``tests/golden/ref_*`` were produced (``tests/golden/gen_golden_reference.py``) and how the CPU tests
compare ``oracle/`` with the reference live.  The only patch applied is ``numpy.int = int`` (removed from NumPy 1.24;
the reference uses it at ``_img_to_tf_mp.py:167`` and ``_img_to_tf_threaded.py:236``).
"""
import importlib
import importlib.util
import os
import sys

REFERENCE_DIR = os.environ.get("B2_REFERENCE_DIR", "/root/reference")
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstubs")
_NAME = "_reference_dl_segmentation_utils"
_STUBBED = ("tensorflow", "rasterio", "descarteslabs", "geopandas", "osgeo")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "dl_segmentation_utils", "__init__.py"))


def load():
    if not available():
        return None
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int                                            # the one patch: NumPy >= 1.24 dropped the alias
    for m in _STUBBED:                                          # a real install must never be shadowed silently
        if m in sys.modules and not getattr(sys.modules[m], "__version__", "").endswith("refstub"):
            raise RuntimeError("%s is already imported from %s" % (m, getattr(sys.modules[m], "__file__", "?")))
    sys.path.insert(0, STUBS)
    try:
        pkg_dir = os.path.join(REFERENCE_DIR, "dl_segmentation_utils")
        spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[_NAME] = mod
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(_NAME, None)
        raise
    finally:
        sys.path.remove(STUBS)
    return mod


def submodule(name):
    """e.g. submodule('_img_to_tf_mp') -> the reference's module object."""
    load()
    return importlib.import_module(_NAME + "." + name)


def stub(name):
    """The stub module the reference was given, e.g. stub('descarteslabs')."""
    load()
    return sys.modules[name]
