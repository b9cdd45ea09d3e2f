"""
TEST INFRASTRUCTURE -- loads the *unmodified* reference numeric backends by
file path from /root/reference (read-only, only present in the build
container, never on the GPU box).

``import syncopy`` itself is impossible here (its __init__ pulls dask, h5py,
matplotlib, ...), but the pure NumPy/SciPy backends on the hot path import only
a handful of light-weight ``syncopy.shared`` modules.  We register empty stub
packages so that no ``__init__.py`` is ever executed, then exec the individual
files under their real dotted names.

Used by oracle/make_golden.py (to generate tests/golden/*.npz) and by the
``needs_reference`` CPU tests (oracle == live reference).  Nothing in the
``-m gpu`` tests, smoke() or bench.py touches this module.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SPYB_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "syncopy")

# (dotted module name, path relative to REF_PKG) -- order matters
_LOAD_ORDER = [
    ("syncopy.shared.log", "shared/log.py"),
    ("syncopy.shared.errors", "shared/errors.py"),
    ("syncopy.shared.filetypes", "shared/filetypes.py"),
    ("syncopy.shared.parsers", "shared/parsers.py"),
    ("syncopy.shared.tools", "shared/tools.py"),
    ("syncopy.shared.const_def", "shared/const_def.py"),
    ("syncopy.specest._norm_spec", "specest/_norm_spec.py"),
    ("syncopy.specest.mtmfft", "specest/mtmfft.py"),
    ("syncopy.specest.stft", "specest/stft.py"),
    ("syncopy.specest.mtmconvol", "specest/mtmconvol.py"),
    ("syncopy.specest.superlet", "specest/superlet.py"),
    ("syncopy.specest.wavelets.wavelets", "specest/wavelets/wavelets.py"),
    ("syncopy.specest.wavelets.transform", "specest/wavelets/transform.py"),
    ("syncopy.specest.wavelet", "specest/wavelet.py"),
    ("syncopy.connectivity.csd", "connectivity/csd.py"),
    ("syncopy.connectivity.wilson_sf", "connectivity/wilson_sf.py"),
    ("syncopy.connectivity.granger", "connectivity/granger.py"),
    ("syncopy.preproc.firws", "preproc/firws.py"),
    ("syncopy.preproc.resampling", "preproc/resampling.py"),
]

_loaded = None


def available():
    return os.path.isfile(os.path.join(REF_PKG, "specest", "mtmfft.py"))


def _stub(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load():
    """Return a namespace with the reference backend modules (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if "syncopy" in sys.modules and not getattr(sys.modules["syncopy"], "_spyb_stub", False):
        raise RuntimeError("a real `syncopy` is already imported; refusing to stub over it")

    top = _stub("syncopy", REF_PKG)
    top._spyb_stub = True
    top.__tbcount__ = 5       # shared/errors.py reads this
    top.__logdir__ = None
    top.__version__ = "2023.09-bypath"
    for pkg in ("specest", "connectivity", "shared", "preproc"):
        _stub(f"syncopy.{pkg}", os.path.join(REF_PKG, pkg))
    _stub("syncopy.specest.wavelets", os.path.join(REF_PKG, "specest", "wavelets"))

    ns = types.SimpleNamespace()
    for dotted, rel in _LOAD_ORDER:
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF_PKG, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = mod
        if dotted == "syncopy.preproc.resampling":
            # `from syncopy.preproc import firws` inside resampling.py
            sys.modules["syncopy.preproc"].firws = sys.modules["syncopy.preproc.firws"]
        if dotted == "syncopy.specest.wavelet":
            # `from syncopy.specest.wavelets import cwt` inside wavelet.py
            sys.modules["syncopy.specest.wavelets"].cwt = sys.modules[
                "syncopy.specest.wavelets.transform"].cwt
        spec.loader.exec_module(mod)
        setattr(ns, dotted.split(".")[-1], mod)
    ns.wavelets_mod = sys.modules["syncopy.specest.wavelets.wavelets"]
    _loaded = ns
    return ns


def extract_function(rel_path, name, env):
    """
    Compile ONE function of a reference module that cannot be imported as a whole (its module pulls h5py / dask at
    import time) straight from the file under /root/reference and return it, bound to the globals in `env`.
    Decorators are dropped (`process_io` only adds the HDF5 plumbing of the parallel runtime).  Nothing is copied
    into this repository: the source is read and executed where it lies.
    """
    import ast
    path = os.path.join(REF_PKG, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ns = dict(env)
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(f"{name} not found in {path}")
