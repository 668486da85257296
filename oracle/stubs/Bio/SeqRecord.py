"""Stand-in for Bio.SeqRecord (test infrastructure; see ../pysam)."""


class SeqRecord(object):
    def __init__(self, seq, id="<unknown id>", description="<unknown description>"):
        self.seq = seq
        self.id = id
        self.description = description
