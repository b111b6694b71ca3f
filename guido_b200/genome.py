"""``Genome`` (drop-in for guido.genome, reference guido/genome.py).

Same constructor, attributes, ``.guido`` pickle and ``load_genome_from_file`` as the reference.
``build()`` differs in what it produces for the search index: instead of shelling out to
``bowtie-build`` (genome.py:156-171) it packs the FASTA into the 2-bit, N-masked genome image
(``<dir>/<genome_name>.g2b``, see csrc/image.cpp) that the kernels scan from HBM.  The path is
stored in ``bowtie_index`` -- the attribute ``find_off_targets`` gates on (locus.py:373,
guides.py:210) -- and additionally in ``device_image_file``.
"""
import gzip
import pickle
import struct
import subprocess
import zlib
from pathlib import Path

from . import _native
from .helpers import is_tool
from .sequence import FastaFile, build_fai

IMAGE_SUFFIX = ".g2b"


def check_file(filename, supported_ext):
    """True when ``filename`` exists and carries one of the supported extensions."""
    filename = Path(filename)
    return bool(filename.exists() and filename.suffix in supported_ext)


def _write_bgzf(path, data):
    """BGZF (blocked gzip, what bgzip writes): <=64 KiB members with a 'BC' extra field + EOF block."""
    with open(path, "wb") as out:
        for i in range(0, len(data), 0xff00):
            chunk = data[i:i + 0xff00]
            comp = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = comp.compress(chunk) + comp.flush()
            bsize = len(body) + 25
            out.write(struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
            out.write(body)
            out.write(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        out.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def _sorted_annotation_lines(path):
    """`sort -k1,1 -k4,4n` (genome.py:141): by sequence name, then numeric start; stable."""
    def key(line):
        cols = line.split("\t")
        try:
            start = float(cols[3]) if len(cols) > 3 else 0.0
        except ValueError:
            start = 0.0
        return (cols[0], start)
    with open(path) as fh:
        lines = [ln if ln.endswith("\n") else ln + "\n" for ln in fh]
    return sorted(lines, key=key)


class Genome:
    def __init__(self, genome_name, genome_file_abspath=None, annotation_file_abspath=None,
                 bowtie_index_abspath=None):
        """A reference genome: name, FASTA path, optional GTF/GFF3 annotation and optional path to an
        existing search index (here: a ``.g2b`` genome image)."""
        self.genome_name = genome_name
        self._bowtie_ignore = False
        self.bowtie_index = None

        if genome_file_abspath:
            if not check_file(genome_file_abspath, [".fa", ".fasta", ".fna"]):
                raise ValueError(
                    f"Genome file {genome_file_abspath} does noth exist or is not in the right format.")
            self.genome_file_abspath = Path(genome_file_abspath).resolve()

        if bowtie_index_abspath:
            self.bowtie_index = Path(bowtie_index_abspath)
            self._bowtie_ignore = True

        if annotation_file_abspath:
            if not check_file(annotation_file_abspath, [".gtf", ".gff3"]):
                raise ValueError(
                    f"Annotation file file {genome_file_abspath} does noth exist or is not in the right format.")
            self.annotation_file_abspath = Path(annotation_file_abspath).resolve()
        else:
            self.annotation_file_abspath = None

    @property
    def is_built(self):
        """Whether the FASTA index, the annotation index and the genome image exist."""
        if not (self.genome_file_abspath and self.fai_file):
            return False
        ready = [True, not (self.annotation_file_abspath and not self.tbi_file)]
        if not self._bowtie_ignore:
            image = self.genome_file_abspath.parent / f"{self.genome_name}{IMAGE_SUFFIX}"
            ready.append(image.exists())
        return all(ready)

    def build(self, n_threads=1, save_pickle=True, bowtie_ignore=False, bowtie_path=""):
        """Index the annotation and the FASTA, pack the genome image, save ``<name>.guido``.

        ``n_threads`` and ``bowtie_path`` are accepted for compatibility and ignored;
        ``bowtie_ignore=True`` skips the genome image (use an existing one via
        ``bowtie_index_abspath``).
        """
        self._bowtie_ignore = bowtie_ignore
        if self.annotation_file_abspath:
            print("Indexing genome annotation.")
            sorted_gz = self.annotation_file_abspath.with_suffix(str(self.annotation_file_abspath.suffix) + ".gz")
            self.annotation_sorted_gz_file = sorted_gz
            _write_bgzf(sorted_gz, "".join(_sorted_annotation_lines(self.annotation_file_abspath)).encode())
            if is_tool("tabix"):
                subprocess.run(f"tabix {sorted_gz}", stderr=subprocess.PIPE, shell=True)
            self.tbi_file = sorted_gz.with_suffix(".tbi")

        if self.genome_file_abspath.exists():
            build_fai(self.genome_file_abspath)
            self.fai_file = self.genome_file_abspath.with_suffix(str(self.genome_file_abspath.suffix) + ".fai")

            if not self._bowtie_ignore:
                print("Building genome image")
                image_path = self.genome_file_abspath.parent / f"{self.genome_name}{IMAGE_SUFFIX}"
                img = _native.Image.from_fasta(self.genome_file_abspath)
                img.save(image_path)
                img.close()
                self.bowtie_index = Path(image_path)
                self.device_image_file = Path(image_path)
                print(f"Done: {self.bowtie_index}")

            if save_pickle:
                guido_file = Path(f"{self.genome_file_abspath.parent / self.genome_name}.guido")
                with open(guido_file, "wb") as fh:
                    pickle.dump(self.__dict__, fh, protocol=pickle.HIGHEST_PROTOCOL)
                print(f"{self.genome_name} genome data can now be used by Guido: {guido_file}")

    @property
    def sequence(self):
        if self.is_built:
            return FastaFile(str(self.genome_file_abspath))

    def device_image(self, device=None):
        """The uploaded genome image of this genome (loaded on first use)."""
        from .off_targets import get_image
        if not self.bowtie_index:
            raise ValueError("Bowtie index is not built for the genome / locus.")
        return get_image(self.bowtie_index, device=device)


def load_genome_from_file(guido_file):
    """Restore a :class:`Genome` from the ``.guido`` file written by :meth:`Genome.build`."""
    if not Path(guido_file).exists():
        raise ValueError(f"Provided path to .guido ({guido_file}) file does not exist.")
    with open(guido_file, "rb") as fh:
        state = pickle.load(fh)
    G = Genome(genome_name=state["genome_name"])
    for key, val in state.items():
        setattr(G, key, val)
    return G
