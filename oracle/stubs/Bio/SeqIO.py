"""Stand-in for Bio.SeqIO (test infrastructure; see ../pysam): FASTA output only
(assemble.py:427)."""


def write(records, handle, fmt):
    if fmt != "fasta":
        raise ValueError("stand-in SeqIO writes FASTA only")
    n = 0
    for rec in records:
        title = ("%s %s" % (rec.id, rec.description)).strip()
        handle.write(">%s\n" % title)
        seq = str(rec.seq)
        for i in range(0, len(seq), 60):
            handle.write(seq[i:i + 60] + "\n")
        n += 1
    return n
