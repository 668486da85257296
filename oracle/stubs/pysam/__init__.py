"""Stand-in for pysam (absent in this image).  TEST INFRASTRUCTURE ONLY.

Two jobs:

* ``import mixemt`` must succeed (``mixemt/__init__.py:19-24`` eagerly imports
  ``assemble``, which imports pysam) so that the reference's modules can be the
  parity oracle.
* The UNMODIFIED CLI (``bin/mixemt``) must run end to end on synthetic
  alignments, with and without ``mixemt_b200.install()``, so that "identical
  reported haplogroup calls and read-to-haplotype assignments" can be checked on
  the reference's own report code (tests/test_cli_gpu.py).  For that this module
  implements the small slice of the pysam API the reference touches:

    FastaFile(fn).references / .fetch(name)            bin/mixemt:157-161
    AlignmentFile(fn, 'rb').count() / .fetch()         bin/mixemt:53-56, :294
    AlignmentFile(fn, 'wb', template=...).write/close  assemble.py:382-390
    AlignedSegment.query_name / mapping_quality / query_sequence /
        query_qualities / is_reverse / get_aligned_pairs(matches_only)
                                                       preprocess.py:119-134, observe.py:65-84

  A "BAM" here is an ``.npz`` written by :func:`write_fake_bam`: ungapped
  alignments (name, 0-based start, sequence, strand, mapping quality).  Written
  "BAM"s are text files with one read name per line.

Nothing here is called by the product package.
"""
import numpy as np


class AlignedSegment(object):
    __slots__ = ("query_name", "reference_start", "query_sequence", "query_qualities",
                 "is_reverse", "mapping_quality")

    def __init__(self, name=None, start=0, seq="", reverse=False, mapq=60, quals=None):
        self.query_name = name
        self.reference_start = start
        self.query_sequence = seq
        self.query_qualities = quals
        self.is_reverse = reverse
        self.mapping_quality = mapq

    def get_aligned_pairs(self, matches_only=False):
        start = self.reference_start
        return [(i, start + i) for i in range(len(self.query_sequence))]


def write_fake_bam(path, names, starts, seqs, reverse=None, mapq=None):
    """Store ungapped alignments; ``seqs`` is a list of str."""
    n = len(names)
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    with open(path, "wb") as handle:
        np.savez(handle,
                 names=np.array("\n".join(names)),
                 starts=np.asarray(starts, dtype=np.int64),
                 seq=np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8),
                 off=off,
                 reverse=np.zeros(n, bool) if reverse is None else np.asarray(reverse, bool),
                 mapq=np.full(n, 60, np.int64) if mapq is None else np.asarray(mapq, np.int64))


class AlignmentFile(object):
    def __init__(self, filename, mode="rb", template=None):
        self.filename = filename
        self.mode = mode
        self._written = []
        self._alns = None
        if "r" in mode:
            with np.load(filename) as z:
                names = str(z["names"]).split("\n") if z["starts"].size else []
                starts = z["starts"].tolist()
                seq = z["seq"].tobytes().decode("ascii")
                off = z["off"].tolist()
                rev = z["reverse"].tolist()
                mapq = z["mapq"].tolist()
            self._alns = [AlignedSegment(names[i], starts[i], seq[off[i]:off[i + 1]], rev[i], mapq[i])
                          for i in range(len(starts))]

    def count(self):
        return len(self._alns)

    def fetch(self):
        return iter(self._alns)

    def write(self, aln):
        self._written.append(aln.query_name)

    def close(self):
        if "w" in self.mode:
            with open(self.filename, "w") as handle:
                for name in self._written:
                    handle.write("%s\n" % name)


class FastaFile(object):
    def __init__(self, filename):
        self._seqs = {}
        self.references = []
        name = None
        with open(filename) as handle:
            for line in handle:
                line = line.strip()
                if line.startswith(">"):
                    name = line[1:].split()[0]
                    self.references.append(name)
                    self._seqs[name] = []
                elif name is not None and line:
                    self._seqs[name].append(line)

    def fetch(self, reference):
        return "".join(self._seqs[reference])
