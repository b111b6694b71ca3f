"""Load the UNMODIFIED reference package (nkran/guido) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists so that golden fixtures can be
generated from, and the C/numpy restatement in this directory validated
against, the reference's own code.  It only works where ``/root/reference``
is mounted (the build container); nothing on the product path or in the
``-m gpu`` tests / ``bench.py`` / ``smoke()`` may import it.

The reference's third-party imports (pyfaidx, pyranges, scikit-allel, mcdm,
azimuth) are absent from this image, so tiny stub modules are registered in
``sys.modules`` first.  The stubs carry no algorithmic content: ``Sequence``
is a plain record with ``name/seq/start/end`` and ``__len__`` (what
guido/locus.py:82-96 reads), everything else is an empty namespace.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GUIDO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "guido"))


class _Sequence:
    """Stand-in for pyfaidx.Sequence (fields used by guido/locus.py:82-96)."""

    def __init__(self, name="", seq="", start=None, end=None, comp=False):
        self.name = name
        self.seq = seq
        self.start = start
        self.end = end
        self.comp = comp

    def __len__(self):
        return len(self.seq)

    def __str__(self):
        return self.seq


def _install_stubs():
    if "pyfaidx" not in sys.modules:
        m = types.ModuleType("pyfaidx")
        m.Sequence = _Sequence
        m.Fasta = type("Fasta", (), {})
        m.Faidx = type("Faidx", (), {})
        sys.modules["pyfaidx"] = m
    for name in ("allel", "mcdm", "pyranges"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "azimuth" not in sys.modules:
        az = types.ModuleType("azimuth")
        mc = types.ModuleType("azimuth.model_comparison")
        az.model_comparison = mc
        sys.modules["azimuth"] = az
        sys.modules["azimuth.model_comparison"] = mc


_cached = None


def load_reference():
    """Return the reference ``guido`` package (imported from /root/reference)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _cached = importlib.import_module("guido")
    return _cached
