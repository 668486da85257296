"""Per-source-line share of executed warp instructions and stall samples.

    ncu -i X.ncu-rep --page source --csv > sass.csv          # per-instruction metrics
    cuobjdump -xelf all lib.so; nvdisasm --print-line-info file.cubin > sass_lines.txt
    python scripts/ncu_lines.py sass.csv sass_lines.txt <kernel name substring> [source file]

The ncu CSV has no line numbers in --csv mode; instructions are joined by
their offset inside the function with nvdisasm's //## File/line annotations.
"""
import collections
import csv
import re
import sys


def load_lines(path, func):
    """offset -> (line, inlined-at line chain text) for one function."""
    out = {}
    cur = None
    active = False
    for raw in open(path, errors="replace"):
        if raw.startswith(".text."):
            active = func in raw
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', raw)
        if m:
            cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)), "inlined" in m.group(3))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", raw)
        if m and cur:
            out[int(m.group(1), 16)] = cur
    return out


def main():
    sass_csv, lines_txt, func = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else None
    lines = load_lines(lines_txt, func)
    rows = list(csv.reader(open(sass_csv)))
    for i, r in enumerate(rows):
        if "Instructions Executed" in r:
            break
    hdr = rows[i]
    idx = {h: j for j, h in enumerate(hdr)}
    data = [r for r in rows[i + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot_i = tot_s = 0.0
    for r in data:
        off = int(r[0], 16) - base
        key = lines.get(off, ("?", 0, False))[:2]
        ins = float(r[idx["Instructions Executed"]] or 0)
        smp = float(r[idx["# Samples"]] or 0)
        agg[key]["inst"] += ins
        agg[key]["samp"] += smp
        agg[key]["thr"] += float(r[idx["Thread Instructions Executed"]] or 0)
        for s in stalls:
            agg[key][s] += float(r[idx[s]] or 0)
        tot_i += ins
        tot_s += smp
    text = {}
    if src:
        for n, l in enumerate(open(src), 1):
            text[n] = l.rstrip()
    print("total warp instructions %.4g, samples %d" % (tot_i, tot_s))
    for key in sorted(agg):
        a = agg[key]
        if a["inst"] / tot_i < 0.004 and a["samp"] / tot_s < 0.004:
            continue
        top = sorted(((a[s], s) for s in stalls), reverse=True)[:2]
        print("%-10s:%4d %5.1f%% inst %5.1f%% samp lanes %4.1f %-26s| %s" % (
            key[0][:10], key[1], 100 * a["inst"] / tot_i, 100 * a["samp"] / tot_s,
            a["thr"] / max(a["inst"], 1), ",".join("%s:%d" % (s[6:], v) for v, s in top if v),
            text.get(key[1], "")[:90].strip() if key[0].startswith(src.rsplit("/", 1)[-1][:10]) else ""))


if __name__ == "__main__":
    main()
