"""Fuzz of the host-side C++ of the library (csrc/parse.cpp, csrc/reduce.cpp) under
AddressSanitizer + UBSan (development aid, no GPU; not collected by pytest):

    mkdir -p /tmp/asan && cd /tmp/asan
    printf '#include <stdarg.h>\nnamespace mxb { void set_error(const char *f, ...) {} }\n' > stub.cpp
    g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -fPIC -shared \
        -I$REPO/include stub.cpp $REPO/mixemt_b200/csrc/parse.cpp $REPO/mixemt_b200/csrc/reduce.cpp \
        -o libhost_asan.so
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
        python $REPO/tests/analysis/host_asan_fuzz.py

Random well-formed and malformed signature buffers of exact size (any overrun is seen) through
mxb_sig_count / mxb_sig_parse, random fragment sets through mxb_reduce_reads / export.
"""
import ctypes, random, sys
import numpy as np
lib = ctypes.CDLL('/tmp/asan/libhost_asan.so')
P = ctypes.c_void_p
def ptr(a): return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else ctypes.c_void_p(a.ctypes.data if a is not None else None)
rnd = random.Random(1)
pos2idx = np.full(17000, -1, dtype=np.int32); pos2idx[::4] = np.arange(len(pos2idx[::4]), dtype=np.int32)
sym2code = np.full(256, 255, dtype=np.uint8)
for i, c in enumerate(b"ACGT"): sym2code[c] = i
alphabet = "0123456789:,ACGTN _+-\n\txé"
def rand_sig():
    mode = rnd.random()
    if mode < 0.6:
        k = rnd.randint(1, 12)
        return ",".join("%d:%s" % (rnd.randint(0, 4200) * 4, rnd.choice("ACGT")) for _ in range(k))
    n = rnd.randint(0, 30)
    return "".join(rnd.choice(alphabet) for _ in range(n))
for it in range(3000):
    reads = [rand_sig() for _ in range(rnd.randint(0, 40))]
    enc = [r.encode("utf-8") for r in reads]
    buf = b"".join(enc)
    # exact-size heap buffer so that any overrun is seen by ASan
    arr = np.frombuffer(buf, dtype=np.uint8).copy() if buf else np.zeros(0, dtype=np.uint8)
    offsets = np.zeros(len(reads) + 1, dtype=np.int64); np.cumsum([len(e) for e in enc], out=offsets[1:])
    n = len(reads)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    rc = lib.mxb_sig_count(P(arr.ctypes.data), P(offsets.ctypes.data), ctypes.c_int64(n), P(row_ptr.ctypes.data))
    assert rc == 0
    total = int(row_ptr[-1])
    pos_idx = np.empty(total, dtype=np.int32); base = np.empty(total, dtype=np.uint8)
    br, bp = ctypes.c_int64(-1), ctypes.c_int64(0)
    rc = lib.mxb_sig_parse(P(arr.ctypes.data), P(offsets.ctypes.data), ctypes.c_int64(n), P(pos2idx.ctypes.data), ctypes.c_int64(len(pos2idx)),
                           P(sym2code.ctypes.data), P(row_ptr.ctypes.data), P(pos_idx.ctypes.data), P(base.ctypes.data), ctypes.byref(br), ctypes.byref(bp))
    assert rc in (0, 2, 3, 4, 5, 6), rc
# reduce
for it in range(600):
    nf = rnd.randint(0, 50)
    lens = [rnd.randint(0, 6) for _ in range(nf)]
    fp = np.zeros(nf + 1, dtype=np.int64); np.cumsum(lens, out=fp[1:])
    tot = int(fp[-1])
    pos = np.array([rnd.choice([-5, 0, 7, 16568, 2**31 - 1, 123456]) for _ in range(tot)], dtype=np.int32)
    # ascending inside a fragment not required for memory safety
    base = np.array([rnd.choice(b"ACGTN") for _ in range(tot)], dtype=np.uint8)
    h = ctypes.c_void_p()
    rc = lib.mxb_reduce_reads(P(fp.ctypes.data), P(pos.ctypes.data) if tot else None, P(base.ctypes.data) if tot else None, ctypes.c_int64(nf), ctypes.byref(h))
    assert rc == 0, rc
    ns, no, nc = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    lib.mxb_sigset_sizes(h, ctypes.byref(ns), ctypes.byref(no), ctypes.byref(nc))
    ns, no, nc = ns.value, no.value, nc.value
    outs = [np.zeros(ns + 1, np.int64), np.empty(no, np.int32), np.empty(no, np.uint8), np.empty(ns, np.int64), np.empty(ns, np.int64),
            np.empty(nf, np.int64), np.empty(nf, np.int64), np.empty(nc, np.uint8), np.zeros(ns + 1, np.int64)]
    lib.mxb_sigset_export(h, *[P(o.ctypes.data) for o in outs])
    assert outs[3].sum() == nf
    lib.mxb_sigset_destroy(h)
print("fuzz ok")
