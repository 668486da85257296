"""Import-time stand-in for Biopython (absent in this image); see ../pysam."""
