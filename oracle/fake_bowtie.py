"""A stand-in for the ``bowtie`` subprocess -- TEST INFRASTRUCTURE ONLY.

The reference shells out to bowtie 1.3.1 (guido/off_targets.py:151-167), which is absent from
this image.  ``patched(ref_pkg, genomes)`` swaps ``subprocess.Popen`` inside the reference's
``guido.off_targets`` module for a class that

* parses the captured command line (``-n``, ``-l``, ``-e``, ``-x``, ``-f``) exactly as the
  reference built it,
* reads the reads from the temp FASTA the reference wrote,
* runs the CPU restatement of bowtie's ``-n`` rule (``oracle.bowtie_n``) over the registered
  genome for that index path, and
* returns bowtie-format TSV from ``communicate()``.

The reference's REAL ``run_bowtie`` / ``Locus.find_off_targets`` then do all parsing,
filtering, record building, CFD and layers -- so everything downstream of the alignment rule
is checked by reference code, and the restatement only decides which windows are reported.
"""
import contextlib

from . import oracle as O


class _FakePopen:
    genomes = {}
    last_argv = None
    last_reads = None

    def __init__(self, argv, **kwargs):
        type(self).last_argv = list(argv)
        self.argv = list(argv)

    def _opt(self, flag):
        return self.argv[self.argv.index(flag) + 1]

    def communicate(self):
        seed_mms = int(self._opt("-n"))
        seed_len = int(self._opt("-l"))
        e = int(self._opt("-e"))
        index = self._opt("-x")
        fasta = self._opt("-f")
        assert "-a" in self.argv and "-y" in self.argv
        names, reads = [], []
        with open(fasta) as fh:
            for line in fh:
                line = line.rstrip("\n")
                if line.startswith(">"):
                    names.append(line[1:])
                elif line:
                    reads.append(line)
        type(self).last_reads = list(zip(names, reads))
        contigs = type(self).genomes[str(index)]
        raw = O.bowtie_n(contigs, reads, seed_mms, seed_len, e)
        return O.bowtie_tsv(raw, names, reads), ""


@contextlib.contextmanager
def patched(ref_pkg, genomes):
    """genomes: {index_path(str): [(contig_name, sequence), ...]}"""
    mod = ref_pkg.off_targets
    saved = mod.subprocess.Popen

    class _SubprocessProxy:
        PIPE = mod.subprocess.PIPE
        Popen = _FakePopen

    _FakePopen.genomes = {str(k): v for k, v in genomes.items()}
    real = mod.subprocess
    mod.subprocess = _SubprocessProxy
    try:
        yield _FakePopen
    finally:
        mod.subprocess = real
        assert real.Popen is saved


class FakeGenome:
    """Duck-typed Genome: Locus.find_off_targets only reads ``.bowtie_index`` (locus.py:366-373)."""

    def __init__(self, index_path):
        self.bowtie_index = index_path
