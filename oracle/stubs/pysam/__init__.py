"""Import-time stand-in for pysam (absent in this image).

Test infrastructure only: lets `import mixemt` succeed so that the reference's
preprocess/em/phylotree modules (which never call pysam) can be used as the
parity oracle. Nothing here is ever called on the hot path.
"""
