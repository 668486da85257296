// Host stand-ins for the CUDA constructs the EM kernels use (csrc/em_kernels.cuh compiled with
// MXB_CPU_EMUL).  TEST INFRASTRUCTURE ONLY -- an interleaving model, not a product path:
// every thread of a CTA is a cooperative fiber (ucontext), scheduled in a seeded random order
// and switched at every synchronising call (__syncthreads, shuffles, mbarrier waits); CTAs run
// one after another.  mbarriers, 1-D bulk copies with transaction counts, warp shuffles and
// block barriers are modelled just far enough to execute the kernels' index arithmetic and
// synchronisation protocol and to detect deadlocks and ring-slot hazards:
//   * a bulk copy lands at a random later scheduling step; in `eager` mode its bytes are
//     written when it is *issued* (a reader that still needed the old contents sees the new
//     ones: write-after-read hazard), otherwise when it *completes* (a reader that did not
//     wait for the barrier sees stale bytes: read-after-write hazard);
//   * a round of the scheduler in which no fiber makes progress and no copy is pending is a
//     deadlock and aborts with the positions of the fibers.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <random>
#include <unordered_map>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __cluster_dims__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }

inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emul {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    long shfl_gen = 0;     // shuffles this lane has entered
};
struct MBar { int count = 0, pending = 0, phase = 0; long tx = 0; };
struct Copy { void *dst; const void *src; uint32_t bytes; uint32_t bar; };
struct Warp { unsigned long long slot[2][32]; long arrivals = 0; };

inline std::vector<Fiber> fibers;
inline std::vector<Warp> warps;
inline int cur = -1;
inline ucontext_t sched_ctx;
inline std::function<void()> body;
inline long progress = 0;          // bumped on every state change another fiber may wait for
inline int bar_count = 0;
inline long bar_gen = 0;
inline std::unordered_map<uint32_t, MBar> mbars;
inline std::vector<Copy> copies;
inline bool eager_copies = false;
inline std::mt19937 rng;
inline long switches = 0;

constexpr uint32_t kSmemTag = 0x1000;      // shared-window address of smem_raw[0]
constexpr size_t kSmemBytes = 232448;

inline void yield() {
    ++switches;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
inline void trampoline() {
    body();
    fibers[cur].done = true;
    ++progress;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
inline void mbar_check(MBar &b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1;
        b.pending = b.count;
        ++progress;
    }
}
inline void land(size_t i) {
    Copy c = copies[i];
    copies.erase(copies.begin() + (long)i);
    if (!eager_copies) memcpy(c.dst, c.src, c.bytes);
    MBar &b = mbars.at(c.bar);
    b.tx -= c.bytes;
    mbar_check(b);
    ++progress;
}

// Runs body() once per thread of every CTA of the grid.
inline void launch(dim3 grid, dim3 block, std::function<void()> fn, unsigned seed, bool eager) {
    gridDim = grid;
    blockDim = block;
    body = fn;
    eager_copies = eager;
    rng.seed(seed);
    const int n = (int)block.x;
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        for (unsigned by = 0; by < grid.y; ++by) {
            blockIdx = {bx, by, 0};
            fibers.assign((size_t)n, Fiber());
            warps.assign((size_t)(n + 31) / 32, Warp());
            mbars.clear();
            copies.clear();
            bar_count = 0;
            for (int t = 0; t < n; ++t) {
                Fiber &f = fibers[(size_t)t];
                f.stack.resize(96 * 1024);
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = f.stack.data();
                f.ctx.uc_stack.ss_size = f.stack.size();
                f.ctx.uc_link = nullptr;
                makecontext(&f.ctx, (void (*)())trampoline, 0);
            }
            std::vector<int> order((size_t)n);
            for (int t = 0; t < n; ++t) order[(size_t)t] = t;
            int live = n;
            while (live > 0) {
                const long before = progress;
                std::shuffle(order.begin(), order.end(), rng);
                for (int t : order) {
                    if (fibers[(size_t)t].done) continue;
                    if (!copies.empty() && rng() % 4 == 0) land(rng() % copies.size());
                    cur = t;
                    threadIdx = {(unsigned)t, 0, 0};
                    swapcontext(&sched_ctx, &fibers[(size_t)t].ctx);
                    if (fibers[(size_t)t].done) --live;
                }
                if (!copies.empty()) land(rng() % copies.size());
                else if (progress == before && live > 0) {
                    fprintf(stderr, "emul: DEADLOCK in CTA (%u,%u): %d threads alive, barrier count %d\n",
                            bx, by, live, bar_count);
                    exit(3);
                }
            }
            if (!copies.empty()) {
                fprintf(stderr, "emul: CTA (%u,%u) exited with %zu bulk copies in flight\n", bx, by,
                        copies.size());
                exit(4);
            }
        }
    }
}

template <class T>
inline T shfl_from(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const int tid = (int)threadIdx.x, lane = tid & 31;
    Warp &w = warps[(size_t)tid >> 5];
    const long g = fibers[(size_t)tid].shfl_gen++;
    const int b = (int)(g & 1);
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[b][lane] = raw;
    ++w.arrivals;
    ++progress;
    while (w.arrivals < 32 * (g + 1)) yield();
    T out;
    memcpy(&out, &w.slot[b][src_lane & 31], sizeof(T));
    return out;
}

}  // namespace emul

// ---- CUDA intrinsics ---------------------------------------------------------------------
inline void __syncthreads() {
    const long g = emul::bar_gen;
    if (++emul::bar_count == (int)blockDim.x) {
        emul::bar_count = 0;
        ++emul::bar_gen;
        ++emul::progress;
    } else {
        while (emul::bar_gen == g) emul::yield();
    }
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask) {
    return emul::shfl_from(v, ((int)threadIdx.x & 31) ^ mask);
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emul::shfl_from(v, src); }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (emul::shfl_from(pred ? 1 : 0, l) ? 1u : 0u) << l;
    return m;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <class T> inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; ++emul::progress; return old; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
    unsigned long long old = *p;
    if (old == cmp) *p = v;
    return old;
}
inline int atomicOr(int *p, int v) { int old = *p; *p = old | v; return old; }

// ---- the helpers of em_kernels.cuh that are inline PTX on the device -------------------------
namespace mxb {
extern unsigned char smem_raw[];
inline uint32_t smem_u32(const void *p) {
    return (uint32_t)((const unsigned char *)p - smem_raw) + emul::kSmemTag;
}
inline void *smem_ptr(uint32_t a) { return smem_raw + (a - emul::kSmemTag); }
inline void mbar_init(uint64_t *bar, int count) {
    emul::MBar b;
    b.count = b.pending = count;
    emul::mbars[smem_u32(bar)] = b;
}
inline void mbar_init_fence() {}
inline void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    emul::MBar &b = emul::mbars.at(bar);
    b.tx += bytes;
    b.pending -= 1;
    if (b.pending < 0) { fprintf(stderr, "emul: too many arrivals on an mbarrier\n"); exit(5); }
    emul::mbar_check(b);
}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { mbar_expect_tx_u32(smem_u32(bar), bytes); }
inline void mbar_arrive_u32(uint32_t bar) {
    emul::MBar &b = emul::mbars.at(bar);
    b.pending -= 1;
    if (b.pending < 0) { fprintf(stderr, "emul: too many arrivals on an mbarrier\n"); exit(5); }
    emul::mbar_check(b);
}
inline void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    while ((uint32_t)emul::mbars.at(bar).phase == parity) emul::yield();
}
inline void mbar_wait(uint64_t *bar, uint32_t parity) { mbar_wait_u32(smem_u32(bar), parity); }
inline void bulk_load_u32(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    if (bytes % 16 != 0 || dst % 16 != 0 || (uintptr_t)src % 16 != 0) {
        fprintf(stderr, "emul: bulk copy of %u bytes is not 16-byte aligned\n", bytes);
        exit(6);
    }
    if (emul::eager_copies) memcpy(smem_ptr(dst), src, bytes);
    emul::copies.push_back({smem_ptr(dst), src, bytes, bar});
}
inline void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    bulk_load_u32(smem_u32(dst), src, bytes, smem_u32(bar));
}
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline uint2 lds_v2_u32(uint32_t a) { uint2 v; memcpy(&v, smem_ptr(a), 8); return v; }
inline double2 lds_v2_f64(uint32_t a) { double2 v; memcpy(&v, smem_ptr(a), 16); return v; }
}  // namespace mxb
