"""The C-ABI shared library on a box without a GPU: it loads, exports every symbol the header
declares, does its host-only work (image build / save / load / fetch), and FAILS LOUDLY -- no CPU
fallback -- for every call that needs a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from guido_b200 import _native
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "guido_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(guido_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/guido_b200.h but not exported"
    assert set(_native.EXPORTED_SYMBOLS) == set(names)
    assert _native.lib().guido_version() == 2


def test_hit_record_layout():
    assert _native.OT_HIT_DTYPE.itemsize == 16
    assert ctypes.sizeof(_native.ImageInfo) == 56 and ctypes.sizeof(_native.Stats) == 56


def test_image_roundtrip(tmp_path):
    contigs = util.synthetic_contigs(2, [1000, 64, 63, 65, 1, 777], n_runs=2, run_len=(1, 30), iupac=4)
    img = _native.Image.from_seqs(contigs)
    assert [(c[0], c[1]) for c in img.contigs] == [(n, len(s)) for n, s in contigs]
    for i, (_, s) in enumerate(contigs):
        assert img.fetch(i, 0, len(s)) == s.upper()
    offs = [c[2] for c in img.contigs]
    assert all(o % 64 == 0 for o in offs) and offs[0] == 64
    assert all(offs[i + 1] >= offs[i] + len(contigs[i][1]) + 64 for i in range(len(offs) - 1))
    img.save(tmp_path / "t.g2b")
    back = _native.Image.load(tmp_path / "t.g2b")
    assert back.contigs == img.contigs and back.info.n_runs == img.info.n_runs
    assert back.fetch(0, 10, 500) == contigs[0][1].upper()[10:510]
    with open(tmp_path / "t.fa", "w") as fh:
        for n, s in contigs:
            fh.write(f">{n} extra words\r\n")
            for i in range(0, len(s), 50):
                fh.write(s[i:i + 50] + "\r\n")
    fa = _native.Image.from_fasta(tmp_path / "t.fa")
    assert fa.contigs == img.contigs
    assert all(fa.fetch(i, 0, c[1]) == img.fetch(i, 0, c[1]) for i, c in enumerate(img.contigs))
    w = img.fetch_windows(np.array([offs[0] - 3, offs[0], offs[1] + 60]), 8)
    assert w[0].tobytes().decode() == "NNN" + contigs[0][1].upper()[:5]
    assert w[1].tobytes().decode() == contigs[0][1].upper()[:8]
    assert w[2].tobytes().decode() == contigs[1][1].upper()[60:64] + "NNNN"
    with pytest.raises(OSError):
        _native.Image.load(tmp_path / "missing.g2b")
    with pytest.raises(ValueError):
        img.fetch(0, 990, 20)


def test_synth_is_deterministic_and_uniform():
    a = _native.Image.synth([200000, 5000], seed=42, n_frac=0.01, run_min=20, run_max=200)
    b = _native.Image.synth([200000, 5000], seed=42, n_frac=0.01, run_min=20, run_max=200)
    c = _native.Image.synth([200000, 5000], seed=43, n_frac=0.01, run_min=20, run_max=200)
    s = a.fetch(0, 0, 200000)
    assert s == b.fetch(0, 0, 200000) and s != c.fetch(0, 0, 200000)
    counts = {k: s.count(k) for k in "ACGTN"}
    assert 1500 < counts["N"] < 3000
    assert all(abs(counts[k] - 49500) < 1500 for k in "ACGT")


def test_encode_guides():
    enc = _native.encode_guides(["ACGT" * 5, "T" * 20])
    assert enc.tolist() == [[int("1100" * 5, 2), int("1010" * 5, 2)], [0xFFFFF, 0xFFFFF]]
    with pytest.raises(ValueError):
        _native.encode_guides(["ACGTN" * 4])


@pytest.mark.skipif(_native.device_count() > 0, reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    img = _native.Image.from_seqs([("c", "ACGT" * 100)])
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        img.upload()
    with pytest.raises(RuntimeError, match="not uploaded|no CPU fallback"):
        img.pam_scan_genome("NGG")
    with pytest.raises(RuntimeError, match="not uploaded|no CPU fallback"):
        img.offtarget_search(["ACGT" * 5])
    with pytest.raises(RuntimeError):
        _native.pam_scan_seq("ACGTGGACGT", "NGG")
    with pytest.raises(RuntimeError):
        _native.mmej_batch(["ACGTACGTAC"], [5])


def test_argument_errors_do_not_need_a_device():
    with pytest.raises(KeyError):
        _native.pam_scan_seq("ACGT", "NGX")
    with pytest.raises(ValueError):
        _native.pam_scan_seq("ACGT", "NNNNNNNNN")
    with pytest.raises(ValueError):
        _native.mmej_batch(["A" * 600], [300])
