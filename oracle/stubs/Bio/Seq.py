"""Stand-in for Bio.Seq (test infrastructure; see ../pysam)."""


class Seq(str):
    pass
