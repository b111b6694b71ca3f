"""Plain sequence record + minimal FASTA index (.fai) reader/writer.

The reference relies on pyfaidx (``Faidx`` to write ``<fasta>.fai`` in Genome.build,
guido/genome.py:151-154; ``Fasta(...).get_seq(chrom, start, end)`` -- 1-based inclusive, case
preserved -- in guido/locus.py:773,853).  pyfaidx is not part of the hot path and is absent from
the target image, so the two operations are provided here with the same conventions.
"""
import os


class Sequence:
    """name / seq / start / end record (the fields of pyfaidx.Sequence that Locus reads)."""

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

    def __repr__(self):
        return f">{self.name}:{self.start}-{self.end}\n{self.seq}"

    def __getitem__(self, item):
        return self.seq[item]


def is_sequence_like(obj):
    return all(hasattr(obj, a) for a in ("seq", "start", "end", "name")) and not isinstance(obj, str)


def build_fai(fasta_path):
    """Write ``<fasta>.fai`` (name, length, offset, line bases, line width) and return its path."""
    fasta_path = os.fspath(fasta_path)
    entries = []
    with open(fasta_path, "rb") as fh:
        name, length, offset, lb, lw, pos = None, 0, 0, 0, 0, 0
        for line in fh:
            n = len(line)
            if line.startswith(b">"):
                if name is not None:
                    entries.append((name, length, offset, lb, lw))
                name = line[1:].split()[0].decode() if line[1:].split() else ""
                length, offset, lb, lw = 0, pos + n, 0, 0
            elif name is not None:
                bases = len(line.rstrip(b"\r\n"))
                if lb == 0 and bases:
                    lb, lw = bases, n
                length += bases
            pos += n
        if name is not None:
            entries.append((name, length, offset, lb, lw))
    fai = fasta_path + ".fai"
    with open(fai, "w") as out:
        for e in entries:
            out.write("\t".join(str(x) for x in e) + "\n")
    return fai


class FastaFile:
    """Random access to a FASTA through its .fai (built on demand)."""

    def __init__(self, fasta_path, one_based_attributes=True):
        self.path = os.fspath(fasta_path)
        fai = self.path + ".fai"
        if not os.path.exists(fai):
            build_fai(self.path)
        self.index = {}
        with open(fai) as fh:
            for line in fh:
                name, length, offset, lb, lw = line.rstrip("\n").split("\t")[:5]
                self.index[name] = (int(length), int(offset), int(lb), int(lw))

    def keys(self):
        return list(self.index)

    def __contains__(self, name):
        return name in self.index

    def get_seq(self, name, start, end):
        """Bases start..end (1-based, inclusive) of sequence ``name``, case preserved."""
        if name not in self.index:
            raise KeyError(f"{name} not in {self.path}.")
        length, offset, lb, lw = self.index[name]
        s0 = max(int(start) - 1, 0)
        e0 = min(int(end), length)
        if e0 <= s0 or lb == 0:
            return Sequence(name=name, seq="", start=start, end=end)
        first = offset + (s0 // lb) * lw + s0 % lb
        last = offset + ((e0 - 1) // lb) * lw + (e0 - 1) % lb + 1
        with open(self.path, "rb") as fh:
            fh.seek(first)
            raw = fh.read(last - first)
        seq = raw.replace(b"\n", b"").replace(b"\r", b"").decode()
        return Sequence(name=name, seq=seq, start=start, end=end)
