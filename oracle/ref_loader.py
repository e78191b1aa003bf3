"""TEST INFRASTRUCTURE ONLY - loader for the UNMODIFIED reference layers (in-container only).

Imports the reference's own Python layers (`src/layers/*.py`) straight from /root/reference and
satisfies their `from src.libs.cutils import ...` with oracle/_ref/cutils.so, which
oracle/build_ref.sh compiles from the reference's `src/libs/cutils.pyx` where it lies.

/root/reference does not exist on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py
may call this.  It is used by tests/golden/make_golden.py (to mint the committed golden vectors)
and by CPU tests that pin the oracle port against the real reference when it is present.
"""
import importlib.machinery
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("AEC_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUTILS_SO = os.path.join(_HERE, "_ref", "cutils.so")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "src", "layers", "conv2d.py")) and os.path.isfile(REF_CUTILS_SO)


def load_reference_cutils():
    """Returns the compiled reference `cutils` module (im2col_event, min_argmax)."""
    name = "src.libs.cutils"
    if name in sys.modules:
        return sys.modules[name]
    if not os.path.isfile(REF_CUTILS_SO):
        raise RuntimeError("oracle/_ref/cutils.so missing - run oracle/build_ref.sh (needs /root/reference)")
    loader = importlib.machinery.ExtensionFileLoader(name, REF_CUTILS_SO)
    spec = importlib.util.spec_from_file_location(name, REF_CUTILS_SO, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def load_reference_layers():
    """Returns (IntegrationLayer, Conv2DLayer, MaxPoolLayer, conv2d_dense, im2col_dense) of the reference."""
    if not reference_available():
        raise RuntimeError("reference not available (no %s or no oracle/_ref/cutils.so)" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cutils = load_reference_cutils()
    import src.libs  # noqa: F401  (regular package under /root/reference)
    setattr(sys.modules["src.libs"], "cutils", cutils)
    from src.layers.integration import IntegrationLayer
    from src.layers.conv2d import Conv2DLayer, conv2d
    from src.layers.maxpool import MaxPoolLayer
    from src.layers.functional import im2col
    return IntegrationLayer, Conv2DLayer, MaxPoolLayer, conv2d, im2col
