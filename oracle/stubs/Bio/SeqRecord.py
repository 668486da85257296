"""Stand-in module; never called by the hot path."""


class SeqRecord(object):
    pass
