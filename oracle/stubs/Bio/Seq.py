"""Stand-in module; never called by the hot path."""


class Seq(object):
    pass
