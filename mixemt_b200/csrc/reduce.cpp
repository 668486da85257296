// Host-side fragment -> signature reduction on binary observations.
//
// Replaces the string round trip of preprocess.read_signature / reduce_reads and
// the `sorted(read_sigs)` / weights of build_em_input (reference
// mixemt/preprocess.py:142-148, :163-174, :218-220): fragments are lists of
// (0-based position, base) pairs, ascending by position like `sorted(obs_by_pos)`;
// equal lists collapse into one signature with a multiplicity.  Rows come out in
// the reference's order -- a plain string sort of "pos:base,pos:base,..."
// ('100:C,200:T' < '10:A' < '9:A') -- so the matrix built from them
// has the reference's row order.  Everything is threaded (OpenMP); at 1M
// fragments this takes tens of milliseconds where the Python loop takes ~30 s.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <numeric>
#include <parallel/algorithm>
#include <vector>

#include "mixemt_b200.h"

namespace mxb {
void set_error(const char *fmt, ...);
}

struct mxb_sigset {
    int64_t n_frag = 0;
    std::vector<int64_t> row_ptr;      // [n_sig + 1]
    std::vector<int32_t> pos;          // [n_obs] 0-based reference positions
    std::vector<uint8_t> base;         // [n_obs] ASCII
    std::vector<int64_t> weights;      // [n_sig]
    std::vector<int64_t> first_frag;   // [n_sig] lowest fragment index carrying the signature
    std::vector<int64_t> sig_of_frag;  // [n_frag]
    std::vector<int64_t> frag_order;   // [n_frag] fragments grouped by row, ascending inside a row
    std::vector<char> strings;         // signature strings, concatenated
    std::vector<int64_t> str_off;      // [n_sig + 1]
};

namespace {

inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

struct FragView {
    const int64_t *ptr;
    const int32_t *pos;
    const uint8_t *base;
    // <0, 0, >0 : content order (length, then positions, then bases) -- any total order works
    int compare(int64_t a, int64_t b) const {
        const int64_t la = ptr[a + 1] - ptr[a], lb = ptr[b + 1] - ptr[b];
        if (la != lb) return la < lb ? -1 : 1;
        int c = memcmp(pos + ptr[a], pos + ptr[b], (size_t)la * sizeof(int32_t));
        if (c) return c;
        return memcmp(base + ptr[a], base + ptr[b], (size_t)la);
    }
};

// "%d:%s" joined by ',' (preprocess.py:142-148); returns the number of chars.
inline size_t render(const int32_t *pos, const uint8_t *base, int64_t k, char *out) {
    char *p = out;
    for (int64_t i = 0; i < k; ++i) {
        if (i) *p++ = ',';
        char tmp[16];
        int n = 0;
        int64_t v = pos[i];
        if (v < 0) { *p++ = '-'; v = -v; }
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) *p++ = tmp[--n];
        *p++ = ':';
        *p++ = (char)base[i];
    }
    return (size_t)(p - out);
}

}  // namespace

extern "C" {

int mxb_reduce_reads(const int64_t *frag_ptr, const int32_t *pos, const uint8_t *base,
                     int64_t n_frag, mxb_sigset **out) {
    if (!out || n_frag < 0 || (n_frag > 0 && !frag_ptr)) {
        mxb::set_error("mxb_reduce_reads: bad argument");
        return MXB_ERR_ARG;
    }
    *out = nullptr;
    if (n_frag > 0 && frag_ptr[n_frag] > frag_ptr[0] && (!pos || !base)) {
        mxb::set_error("mxb_reduce_reads: NULL observation arrays");
        return MXB_ERR_ARG;
    }
    try {
        mxb_sigset *ss = new mxb_sigset();
        ss->n_frag = n_frag;
        const FragView fv{frag_ptr, pos, base};
        // 1. hash every fragment, sort fragment ids by (hash, content, id)
        std::vector<uint64_t> hash((size_t)n_frag);
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < n_frag; ++f) {
            const int64_t a = frag_ptr[f], b = frag_ptr[f + 1];
            uint64_t h = mix64((uint64_t)(b - a) + 0x9e3779b97f4a7c15ULL);
            for (int64_t k = a; k < b; ++k)
                h = mix64(h ^ (((uint64_t)(uint32_t)pos[k] << 8) | base[k]));
            hash[(size_t)f] = h;
        }
        std::vector<int64_t> order((size_t)n_frag);
        std::iota(order.begin(), order.end(), (int64_t)0);
        __gnu_parallel::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
            if (hash[(size_t)a] != hash[(size_t)b]) return hash[(size_t)a] < hash[(size_t)b];
            const int c = fv.compare(a, b);
            if (c) return c < 0;
            return a < b;
        });
        // 2. group boundaries: representatives are the lowest fragment id of each group
        std::vector<int64_t> starts;
        for (int64_t i = 0; i < n_frag; ++i) {
            if (i == 0 || hash[(size_t)order[i]] != hash[(size_t)order[i - 1]] ||
                fv.compare(order[i], order[i - 1]) != 0)
                starts.push_back(i);
        }
        const int64_t n_sig = (int64_t)starts.size();
        starts.push_back(n_frag);
        // 3. render the signature strings of the representatives
        std::vector<int64_t> raw_off((size_t)n_sig + 1, 0);
        for (int64_t u = 0; u < n_sig; ++u) {
            const int64_t f = order[starts[u]];
            // worst case per observation: ',' + 11 chars of position + ':' + base
            raw_off[(size_t)u + 1] = raw_off[(size_t)u] + 14 * (frag_ptr[f + 1] - frag_ptr[f]) + 1;
        }
        std::vector<char> raw((size_t)raw_off[(size_t)n_sig]);
        std::vector<int64_t> raw_len((size_t)n_sig);
#pragma omp parallel for schedule(static)
        for (int64_t u = 0; u < n_sig; ++u) {
            const int64_t f = order[starts[u]];
            raw_len[(size_t)u] = (int64_t)render(pos + frag_ptr[f], base + frag_ptr[f],
                                                frag_ptr[f + 1] - frag_ptr[f],
                                                raw.data() + raw_off[(size_t)u]);
        }
        // 4. the reference's row order: plain string comparison (bytes; ASCII only here)
        std::vector<int64_t> rank((size_t)n_sig);
        std::iota(rank.begin(), rank.end(), (int64_t)0);
        __gnu_parallel::sort(rank.begin(), rank.end(), [&](int64_t a, int64_t b) {
            const int64_t la = raw_len[(size_t)a], lb = raw_len[(size_t)b];
            const int c = memcmp(raw.data() + raw_off[(size_t)a], raw.data() + raw_off[(size_t)b],
                                 (size_t)std::min(la, lb));
            if (c) return c < 0;
            if (la != lb) return la < lb;
            return a < b;
        });
        // 5. emit rows in that order
        ss->row_ptr.assign((size_t)n_sig + 1, 0);
        ss->str_off.assign((size_t)n_sig + 1, 0);
        ss->weights.resize((size_t)n_sig);
        ss->first_frag.resize((size_t)n_sig);
        for (int64_t r = 0; r < n_sig; ++r) {
            const int64_t u = rank[(size_t)r];
            const int64_t f = order[starts[u]];
            ss->row_ptr[(size_t)r + 1] = ss->row_ptr[(size_t)r] + (frag_ptr[f + 1] - frag_ptr[f]);
            ss->str_off[(size_t)r + 1] = ss->str_off[(size_t)r] + raw_len[(size_t)u];
            ss->weights[(size_t)r] = starts[u + 1] - starts[u];
            ss->first_frag[(size_t)r] = f;
        }
        ss->pos.resize((size_t)ss->row_ptr[(size_t)n_sig]);
        ss->base.resize((size_t)ss->row_ptr[(size_t)n_sig]);
        ss->strings.resize((size_t)ss->str_off[(size_t)n_sig]);
        ss->sig_of_frag.resize((size_t)n_frag);
        ss->frag_order.resize((size_t)n_frag);
        std::vector<int64_t> frag_off((size_t)n_sig + 1, 0);
        for (int64_t r = 0; r < n_sig; ++r)
            frag_off[(size_t)r + 1] = frag_off[(size_t)r] + ss->weights[(size_t)r];
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t r = 0; r < n_sig; ++r) {
            const int64_t u = rank[(size_t)r];
            const int64_t f = order[starts[u]];
            const int64_t k = frag_ptr[f + 1] - frag_ptr[f];
            if (k) {
                memcpy(ss->pos.data() + ss->row_ptr[(size_t)r], pos + frag_ptr[f],
                       (size_t)k * sizeof(int32_t));
                memcpy(ss->base.data() + ss->row_ptr[(size_t)r], base + frag_ptr[f], (size_t)k);
            }
            if (raw_len[(size_t)u])
                memcpy(ss->strings.data() + ss->str_off[(size_t)r], raw.data() + raw_off[(size_t)u],
                       (size_t)raw_len[(size_t)u]);
            for (int64_t i = starts[u]; i < starts[u + 1]; ++i) {
                ss->sig_of_frag[(size_t)order[i]] = r;
                ss->frag_order[(size_t)(frag_off[(size_t)r] + (i - starts[u]))] = order[i];
            }
        }
        *out = ss;
        return MXB_OK;
    } catch (const std::bad_alloc &) {
        mxb::set_error("mxb_reduce_reads: out of host memory");
        return MXB_ERR_NOMEM;
    }
}

int mxb_sigset_sizes(const mxb_sigset *ss, int64_t *n_sig, int64_t *n_obs, int64_t *n_chars) {
    if (!ss) { mxb::set_error("mxb_sigset_sizes: NULL handle"); return MXB_ERR_ARG; }
    if (n_sig) *n_sig = (int64_t)ss->weights.size();
    if (n_obs) *n_obs = (int64_t)ss->pos.size();
    if (n_chars) *n_chars = (int64_t)ss->strings.size();
    return MXB_OK;
}

int mxb_sigset_export(const mxb_sigset *ss, int64_t *row_ptr, int32_t *pos, uint8_t *base,
                      int64_t *weights, int64_t *first_frag, int64_t *sig_of_frag,
                      int64_t *frag_order, char *strings, int64_t *str_offsets) {
    if (!ss) { mxb::set_error("mxb_sigset_export: NULL handle"); return MXB_ERR_ARG; }
#define COPY(dst, vec) do { if (dst && !(vec).empty()) memcpy(dst, (vec).data(), (vec).size() * sizeof((vec)[0])); } while (0)
    COPY(row_ptr, ss->row_ptr);
    COPY(pos, ss->pos);
    COPY(base, ss->base);
    COPY(weights, ss->weights);
    COPY(first_frag, ss->first_frag);
    COPY(sig_of_frag, ss->sig_of_frag);
    COPY(frag_order, ss->frag_order);
    COPY(strings, ss->strings);
    COPY(str_offsets, ss->str_off);
#undef COPY
    return MXB_OK;
}

int mxb_sigset_destroy(mxb_sigset *ss) {
    delete ss;
    return MXB_OK;
}

}  // extern "C"
