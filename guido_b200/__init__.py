"""guido_b200 -- B200-native hot path of nkran/guido behind guido's own Python API.

``import guido_b200 as guido`` is the intended switch: ``Genome`` / ``genome.build()``,
``load_genome_from_file``, ``Locus``, ``locus_from_coordinates`` / ``locus_from_gene`` /
``locus_from_sequence``, ``Guide``, ``rev_comp`` and ``is_tool`` keep the reference's signatures
and record formats (reference guido/__init__.py:2-5).  The PAM scan, the off-target search and the
MMEJ enumeration run as hand-written sm_100a kernels behind the C-ABI in ``include/guido_b200.h``;
there is no CPU fallback.
"""
# flake8: noqa
from .genome import Genome, load_genome_from_file
from .guides import Guide
from .helpers import is_tool, rev_comp
from .locus import Locus, locus_from_coordinates, locus_from_gene, locus_from_sequence

__version__ = "0.1.0"
