"""TEST INFRASTRUCTURE ONLY -- loads the *real* Kover reference for validation.

The reference (``/root/reference``, Python 2.7 + Cython) cannot be imported
unmodified on Python 3.12 / numpy 2.  This loader reads the
hot-path source files where they lie, applies the purely mechanical Py2->Py3
substitutions listed below IN MEMORY (no reference source is ever written into
this repository) and executes them as modules of a synthetic package named
``kover_ref``.  The only artefact written to disk is the compiled popcount
extension (``oracle/_ref/popcount*.so``), built from the reference's own
``popcount.pyx`` with Cython + gcc; ``oracle/_ref/`` is git-ignored.

It is available only where ``/root/reference`` exists (this container).  The GPU
box has no reference tree: nothing on the ``-m gpu`` path may import this file.
It is used by ``tests/golden/make_golden.py`` (to mint the committed known-answer
vectors) and by ``tests/test_oracle_vs_reference.py`` (skipped when absent).

Mechanical patch list (each is a syntax/stdlib rename, never a semantic change):
  xrange -> range; dict.iter{items,keys,values} -> {items,keys,values};
  np.bool/np.float/np.infty -> bool/float/np.inf; integer ``/`` used as floor
  division in build_row_mask (rules.py:218,222), a debug format (scm.py:84) and ``n_kmers =
  shape[1] / 2`` (experiment_scm.py:286,455) -> ``//``; ``d.keys()[0]`` -> ``list(d.keys())[0]`` (cart.py:141, tree.py:108);
  ``scipy.misc.comb`` -> ``scipy.special.comb``; ``np.var(d.values())`` -> ``np.var(list(d.values()))``
  (experiment_cart.py:477); ``np.vstack(<generator>)`` -> ``np.vstack([...])``
  (experiment_cart.py:139); tabs expanded in models.py / metrics.py;
  ``import h5py`` satisfied by a stub (h5py is not installed in this image).
"""
import importlib.machinery
import importlib.util
import os
import re
import subprocess
import sys
import sysconfig
import tempfile
import types

REFERENCE_ROOT = os.environ.get("KOVER_REFERENCE_ROOT", "/root/reference")
_CORE = os.path.join(REFERENCE_ROOT, "core", "kover")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_OUT = os.path.join(_HERE, "_ref")
_PKG = "kover_ref"


def reference_available():
    return os.path.isfile(os.path.join(_CORE, "learning", "common", "rules.py"))


_SUBS = [
    (r"\bxrange\b", "range"),
    (r"\.iteritems\(\)", ".items()"),
    (r"\.iterkeys\(\)", ".keys()"),
    (r"\.itervalues\(\)", ".values()"),
    (r"\bnp\.bool\b", "bool"),
    (r"\bnp\.float\b", "float"),
    (r"\bnp\.infty\b", "np.inf"),
    (r"idx / mask_n_bits", "idx // mask_n_bits"),
    (r"mask_n_bits / 8", "mask_n_bits // 8"),
    (r"len\(rule_blacklist\) / 2", "len(rule_blacklist) // 2"),
    (r"rule_classifications\.shape\[1\] / 2", "rule_classifications.shape[1] // 2"),      # experiment_scm.py:286, 455
    # dict.values() is a view on Python 3 (experiment_cart.py:477)
    (r'np\.var\((\w+)\["class_importance"\]\.values\(\)\)', r'np.var(list(\1["class_importance"].values()))'),
    (r"left_n_examples_by_class\.keys\(\)\[0\]", "list(left_n_examples_by_class.keys())[0]"),
    (r"self\.breiman_info\.p_j_given_t\.keys\(\)\[np\.argmax\(self\.breiman_info\.p_j_given_t\.values\(\)\)\]",
     "list(self.breiman_info.p_j_given_t.keys())[np.argmax(list(self.breiman_info.p_j_given_t.values()))]"),
    (r"from scipy\.misc import comb", "from scipy.special import comb"),
    # numpy >= 1.16 refuses a generator as the argument of vstack (experiment_cart.py:139)
    (r"np\.vstack\(\((_unpack_binary_bytes_from_ints\(kmer_matrix\[:, idx\]\) for idx in kmer_idx_by_rule)\)\)",
     r"np.vstack([\1])"),
    (r"^import h5py as h$", "import types as _t; h = _t.SimpleNamespace(h5f=_t.SimpleNamespace(ACC_RDONLY=0, ACC_RDWR=1))  # h5py stub"),
]


def _patched_source(path):
    with open(path, "r") as f:
        src = f.read()
    # Python 2 reads a tab as "up to the next multiple of 8"; models.py mixes tabs consistently with 4
    src = src.expandtabs(4) if path.endswith("models.py") else src
    src = src.expandtabs(8) if path.endswith("metrics.py") else src
    for pat, rep in _SUBS:
        src = re.sub(pat, rep, src, flags=re.M)
    return src


def build_reference_popcount(force=False):
    """Cythonize + compile the reference's popcount.pyx into oracle/_ref/ (binary only)."""
    os.makedirs(_REF_OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    so_path = os.path.join(_REF_OUT, "popcount" + suffix)
    if os.path.isfile(so_path) and not force:
        return so_path
    if not reference_available():
        raise RuntimeError("reference tree not present; cannot build oracle/_ref")
    import numpy as np
    pyx = os.path.join(_CORE, "learning", "common", "popcount.pyx")
    with tempfile.TemporaryDirectory() as tmp:
        c_path = os.path.join(tmp, "popcount.c")
        subprocess.check_call([sys.executable, "-m", "cython", "-2", pyx, "-o", c_path],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        inc = sysconfig.get_paths()["include"]
        # same optimisation level the reference asks for (core/setup.py:30-31: -march=native)
        subprocess.check_call(["gcc", "-O2", "-march=native", "-shared", "-fPIC", "-w",
                               "-I", inc, "-I", np.get_include(), c_path, "-o", so_path])
    return so_path


def load_reference_popcount():
    """Import the compiled reference popcount extension (works without /root/reference
    once oracle/_ref/ holds the binary)."""
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    so_path = os.path.join(_REF_OUT, "popcount" + suffix)
    if not os.path.isfile(so_path):
        so_path = build_reference_popcount()
    name = _PKG + ".learning.common.popcount"
    if name in sys.modules:
        return sys.modules[name]
    loader = importlib.machinery.ExtensionFileLoader("popcount", so_path)
    spec = importlib.util.spec_from_loader("popcount", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def _mk_pkg(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = []
    m.__package__ = name
    sys.modules[name] = m
    return m


def _exec_module(name, path):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__package__ = name.rsplit(".", 1)[0]
    m.__file__ = path
    sys.modules[name] = m
    code = compile(_patched_source(path), path, "exec")
    exec(code, m.__dict__)
    return m


def load_reference():
    """Returns a namespace with the reference's hot-path modules:
    .utils .rules .models .tree .scm .cart .popcount"""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (_PKG, _PKG + ".learning", _PKG + ".learning.common", _PKG + ".learning.learners"):
        _mk_pkg(p)
    ns = types.SimpleNamespace()
    ns.popcount = load_reference_popcount()
    ns.utils = _exec_module(_PKG + ".utils", os.path.join(_CORE, "utils.py"))
    ns.rules = _exec_module(_PKG + ".learning.common.rules", os.path.join(_CORE, "learning/common/rules.py"))
    ns.models = _exec_module(_PKG + ".learning.common.models", os.path.join(_CORE, "learning/common/models.py"))
    ns.tree = _exec_module(_PKG + ".learning.common.tree", os.path.join(_CORE, "learning/common/tree.py"))
    ns.scm = _exec_module(_PKG + ".learning.learners.scm", os.path.join(_CORE, "learning/learners/scm.py"))
    ns.cart = _exec_module(_PKG + ".learning.learners.cart", os.path.join(_CORE, "learning/learners/cart.py"))
    # experiment drivers: only their pure functions (_bound, _predictions, metrics) are usable without HDF5
    _mk_pkg(_PKG + ".dataset")
    _mk_pkg(_PKG + ".learning.experiments")
    if _PKG + ".dataset.ds" not in sys.modules:
        ds = types.ModuleType(_PKG + ".dataset.ds")
        ds.KoverDataset = type("KoverDataset", (), {})        # stub: needs h5py, never instantiated here
        sys.modules[_PKG + ".dataset.ds"] = ds
    ns.metrics = _exec_module(_PKG + ".learning.experiments.metrics",
                              os.path.join(_CORE, "learning/experiments/metrics.py"))
    ns.experiment_scm = _exec_module(_PKG + ".learning.experiments.experiment_scm",
                                     os.path.join(_CORE, "learning/experiments/experiment_scm.py"))
    ns.experiment_cart = _exec_module(_PKG + ".learning.experiments.experiment_cart",
                                      os.path.join(_CORE, "learning/experiments/experiment_cart.py"))
    return ns


class ChunkedArray(object):
    """A numpy-backed stand-in for the h5py ``kmer_matrix`` dataset: shape/dtype/chunks and
    numpy-style ``__getitem__`` (the only h5py surface rules.py touches: rules.py:113-116,
    123-128, 164, 253)."""

    def __init__(self, array, chunks=None):
        self._a = array
        self.shape = array.shape
        self.dtype = array.dtype
        self.chunks = chunks

    def __getitem__(self, key):
        out = self._a[key]
        # h5py returns fresh arrays; the reference pop-counts them in place (rules.py:259)
        return out.copy()


if __name__ == "__main__":
    ref = load_reference()
    print("reference loaded:", [k for k in vars(ref)])
