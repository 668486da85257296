"""Index arithmetic of the chunk-dictionary records (em_pack_pairs_kernel / em_pass_coded_pairs_kernel
in csrc/em.cu) restated in numpy: pack a row, then read it back the way a pass-kernel thread does
(8-byte code word, byte k, table offset code << 4) and compare with the row.  Checks the layout
formulas, not the CUDA code itself."""
import numpy as np


def pack_row(row, threads):
    ld = len(row)
    n_chunks = ld // 2
    bits = row.view(np.uint64).reshape(n_chunks, 2)
    uniq, inv = np.unique(bits, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    if len(uniq) > 256:
        return None
    rec = np.zeros(threads * 8 + 256 * 16, dtype=np.uint8)
    for c in range(n_chunks):
        rec[(c % threads) * 8 + c // threads] = inv[c]          # out[(c % T) * 8 + c / T]
    tab = rec[threads * 8:].view(np.uint64).reshape(256, 2)
    tab[:len(uniq)] = uniq
    return rec


def read_back(rec, ld, threads, nc):
    n_chunks = ld // 2
    out = np.full(ld, np.nan)
    tab_bytes = rec[threads * 8:]
    for tid in range(threads):
        cw = rec[tid * 8: tid * 8 + 8].view(np.uint32)          # ld.shared.v2.u32
        for k in range(nc):
            word = int(cw[0] if k < 4 else cw[1])
            off = ((word >> (8 * (k & 3))) & 0xFF) << 4
            v = tab_bytes[off: off + 16].view(np.float64)       # ld.shared.v2.f64
            c = tid + k * threads
            if c < n_chunks:
                out[2 * c], out[2 * c + 1] = v[0], v[1]
    return out


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    for ld in (5408, 1024, 8192, 4976, 6144):
        for threads in (512, 384):
            nc = -(-(ld // 2) // threads)
            if nc > 8:
                continue
            vals = rs.rand(40)
            row = vals[np.repeat(rs.randint(0, 40, size=ld // 8 + 1), 8)[:ld]].copy()
            row[rs.randint(0, ld, size=60)] = rs.rand(60)        # a few odd cells
            rec = pack_row(row, threads)
            assert rec is not None
            back = read_back(rec, ld, threads, nc)
            assert np.array_equal(back, row), (ld, threads)
            print("ld %d threads %d nc %d: %d-byte record reads back exactly" % (ld, threads, nc, len(rec)))
