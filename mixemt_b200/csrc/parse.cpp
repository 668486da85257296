// Host-side signature marshalling: "pos:base,pos:base,..." strings -> CSR.
//
// Replaces the per-row Python of preprocess.pos_obs_from_sig
// (reference mixemt/preprocess.py:151-160) with one threaded pass.  The error
// contract mirrors what the reference's double loop (preprocess.py:188-191)
// would raise first: rows are visited in order; inside a row the whole
// signature is parsed before any lookup (ValueError from a malformed field
// wins over KeyError from an unknown position of the same row).
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "mixemt_b200.h"

namespace mxb {
void set_error(const char *fmt, ...);
}

namespace {

inline bool is_space(unsigned char c) {
    return c == ' ' || (c >= '\t' && c <= '\r');
}

// Python int(str) for base 10: optional surrounding whitespace, optional sign,
// digits with optional single underscores between digits.
// Returns false when int() would raise ValueError.  Saturates at +-2^62.
inline bool parse_py_int(const char *s, const char *e, int64_t *out) {
    while (s < e && is_space((unsigned char)*s)) ++s;
    while (e > s && is_space((unsigned char)e[-1])) --e;
    if (s == e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') { neg = (*s == '-'); ++s; }
    if (s == e) return false;
    int64_t v = 0;
    bool prev_digit = false;
    for (; s < e; ++s) {
        unsigned char c = (unsigned char)*s;
        if (c >= '0' && c <= '9') {
            if (v < ((int64_t)1 << 58)) v = v * 10 + (c - '0');
            else v = (int64_t)1 << 62;
            prev_digit = true;
        } else if (c == '_' && prev_digit && s + 1 < e && s[1] >= '0' && s[1] <= '9') {
            prev_digit = false;
        } else {
            return false;
        }
    }
    *out = neg ? -v : v;
    return true;
}

}  // namespace

extern "C" {

int mxb_sig_count(const char *buf, const int64_t *offsets, int64_t n_rows,
                  int64_t *row_ptr) {
    if (n_rows < 0 || !offsets || !row_ptr || (!buf && n_rows > 0 && offsets[n_rows] > 0)) {
        mxb::set_error("mxb_sig_count: bad argument");
        return MXB_ERR_ARG;
    }
    row_ptr[0] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; ++r) {
        const char *s = buf + offsets[r];
        const char *e = buf + offsets[r + 1];
        int64_t k = 1;
        for (const char *p = s; p < e; ++p) k += (*p == ',');
        row_ptr[r + 1] = k;
    }
    for (int64_t r = 0; r < n_rows; ++r) row_ptr[r + 1] += row_ptr[r];
    return MXB_OK;
}

int mxb_sig_parse(const char *buf, const int64_t *offsets, int64_t n_rows,
                  const int32_t *pos2idx, int64_t pos2idx_len,
                  const uint8_t *sym2code, const int64_t *row_ptr,
                  int32_t *pos_idx, uint8_t *base_code,
                  int64_t *bad_row, int64_t *bad_pos) {
    if (n_rows < 0 || !offsets || !row_ptr || !sym2code || pos2idx_len < 0 ||
        (pos2idx_len > 0 && !pos2idx) || !bad_row || !bad_pos) {
        mxb::set_error("mxb_sig_parse: bad argument");
        return MXB_ERR_ARG;
    }
    int64_t first_bad = INT64_MAX;  // row index of the first failing row
    int first_kind = MXB_OK;
    int64_t first_pos = 0;
#pragma omp parallel
    {
        int64_t my_bad = INT64_MAX, my_pos = 0;
        int my_kind = MXB_OK;
#pragma omp for schedule(static) nowait
        for (int64_t r = 0; r < n_rows; ++r) {
            if (r > my_bad) continue;
            const char *s = buf + offsets[r];
            const char *e = buf + offsets[r + 1];
            int64_t k = row_ptr[r];
            const int64_t k_end = row_ptr[r + 1];
            bool malformed = false;
            bool have_key = false;
            int64_t key_pos = 0;
            const char *f = s;
            while (true) {
                const char *fe = (const char *)memchr(f, ',', (size_t)(e - f));
                if (!fe) fe = e;
                // field [f, fe): exactly one ':' required (var.split(':') -> 2 items)
                const char *colon = (const char *)memchr(f, ':', (size_t)(fe - f));
                if (!colon || memchr(colon + 1, ':', (size_t)(fe - colon - 1))) {
                    malformed = true;
                    break;
                }
                int64_t pos;
                if (!parse_py_int(f, colon, &pos)) { malformed = true; break; }
                if (k >= k_end) { malformed = true; break; }  // row_ptr inconsistent
                int32_t idx = (pos >= 0 && pos < pos2idx_len) ? pos2idx[pos] : -1;
                if (idx < 0 && !have_key) { have_key = true; key_pos = pos; }
                pos_idx[k] = idx < 0 ? 0 : idx;
                base_code[k] = (fe - colon - 1 == 1) ? sym2code[(unsigned char)colon[1]] : 255;
                ++k;
                if (fe == e) break;
                f = fe + 1;
            }
            if (malformed) { my_bad = r; my_kind = MXB_ERR_VALUE; my_pos = 0; }
            else if (have_key) { my_bad = r; my_kind = MXB_ERR_KEY; my_pos = key_pos; }
        }
#pragma omp critical
        {
            if (my_bad < first_bad) { first_bad = my_bad; first_kind = my_kind; first_pos = my_pos; }
        }
    }
    if (first_kind != MXB_OK) {
        *bad_row = first_bad;
        *bad_pos = first_pos;
        if (first_kind == MXB_ERR_VALUE)
            mxb::set_error("malformed signature in row %lld", (long long)first_bad);
        else
            mxb::set_error("position %lld (row %lld) is not a known variant site",
                           (long long)first_pos, (long long)first_bad);
        return first_kind;
    }
    *bad_row = -1;
    *bad_pos = 0;
    return MXB_OK;
}

}  // extern "C"
