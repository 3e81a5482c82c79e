"""MEASUREMENT / TEST INFRASTRUCTURE (not product code): locating and importing the UNMODIFIED reference.

The reference tree is found at ``baseline/_ref/`` (git-ignored copy staged by tools/stage_reference.py; it travels to
the GPU box) or at ``/root/reference`` (build container only). Used by tests/ and by ``bench.py``'s reference /
baseline legs (tools/bench_legs.py) — never by the product package. It is a loader only: nothing of ``oracle/`` is
involved when the reference's own modules are timed.
"""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")

# plotting / image / timezone libraries utils/trainer.py and utils/utils.py import at module level; none of them is
# installed in this image and none is needed by train_one_epoch / validate
STUBS = ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.measure", "pytz", "seaborn")


def ref_root():
    for p in CANDIDATES:
        if os.path.isfile(os.path.join(p, "models", "model.py")):
            return p
    return None


def load_ref_module(rel, name=None):
    """Imports one reference file (e.g. 'models/model.py') under a private module name, leaving ``sys.modules['models']``
    alone so that the drop-in package and the reference can live in one process."""
    root = ref_root()
    if root is None:
        raise FileNotFoundError("reference tree not found (baseline/_ref or /root/reference)")
    name = name or "b2s_reference_" + rel.replace("/", "_").replace(".py", "")
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def stub_missing_modules():
    """Registers empty stand-ins for the plotting libraries that are not installed. Returns the list of stubbed names."""
    done = []
    for name in STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
            done.append(name)
    if "matplotlib" in done or "matplotlib.pyplot" in done:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "skimage" in done or "skimage.measure" in done:
        sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    return done
