// orlg_device.cuh -- device-side building blocks of the batched step path (sm_100a).
//
// Spectrum state is bit-packed: one 128-bit word group per (core, link, env); bit s = slot s
// free (the reference's available_slots[link, s] == 1, rmsa_env.py:337-339); bits >= S are 0.
// Every RLE / np.any / np.sum of the reference collapses to popc / ffs / funnel shifts here
// (identities verified against the reference in SURVEY.md section 7 step 6).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orlg {

constexpr int NW = 4;              // 32-bit words per (core, link): S <= 128
constexpr int MAX_SLOTS = 32 * NW;
constexpr uint32_t CAND_NONE = 0xFFu;

// ---------------------------------------------------------------- 128-bit slot masks
struct Bits {
    uint32_t w[NW];
};

__device__ __forceinline__ Bits bits_from(uint4 v) {
    Bits b;
    b.w[0] = v.x; b.w[1] = v.y; b.w[2] = v.z; b.w[3] = v.w;
    return b;
}
__device__ __forceinline__ uint4 bits_to(const Bits &b) { return make_uint4(b.w[0], b.w[1], b.w[2], b.w[3]); }
__device__ __forceinline__ Bits bits_ones() {
    Bits b;
#pragma unroll
    for (int i = 0; i < NW; i++) b.w[i] = 0xFFFFFFFFu;
    return b;
}
__device__ __forceinline__ Bits bits_and(const Bits &a, const Bits &b) {
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = a.w[i] & b.w[i];
    return r;
}
__device__ __forceinline__ Bits bits_or(const Bits &a, const Bits &b) {
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = a.w[i] | b.w[i];
    return r;
}
__device__ __forceinline__ Bits bits_andnot(const Bits &a, const Bits &b) {   // a & ~b
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = a.w[i] & ~b.w[i];
    return r;
}
__device__ __forceinline__ bool bits_contains(const Bits &a, const Bits &m) {   // (a & m) == m
    uint32_t miss = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) miss |= (~a.w[i]) & m.w[i];
    return miss == 0;
}
// logical shift right by 0 < s < 32
__device__ __forceinline__ Bits bits_shr_small(const Bits &a, int s) {
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = __funnelshift_r(a.w[i], i + 1 < NW ? a.w[i + 1] : 0u, s);
    return r;
}
__device__ __forceinline__ Bits bits_shl1(const Bits &a) {
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = __funnelshift_l(i > 0 ? a.w[i - 1] : 0u, a.w[i], 1);
    return r;
}
__device__ __forceinline__ int bits_popc(const Bits &a) {
    int c = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) c += __popc(a.w[i]);
    return c;
}
// index of the lowest set bit, or -1
__device__ __forceinline__ int bits_ffs(const Bits &a) {
    int r = -1;
#pragma unroll
    for (int i = NW - 1; i >= 0; i--)
        if (a.w[i]) r = 32 * i + __ffs(a.w[i]) - 1;
    return r;
}
// index of the highest set bit, or -1
__device__ __forceinline__ int bits_fls(const Bits &a) {
    int r = -1;
#pragma unroll
    for (int i = 0; i < NW; i++)
        if (a.w[i]) r = 32 * i + 31 - __clz(a.w[i]);
    return r;
}
__device__ __forceinline__ Bits bits_clear_lowest(const Bits &a) {
    Bits r = a;
    bool done = false;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        if (!done && r.w[i]) { r.w[i] &= r.w[i] - 1; done = true; }
    }
    return r;
}
// bits [lo, hi) set, 0 <= lo <= hi <= 128
__device__ __forceinline__ Bits bits_range(int lo, int hi) {
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        int a = min(max(lo - 32 * i, 0), 32), b = min(max(hi - 32 * i, 0), 32);
        uint32_t ma = a >= 32 ? 0u : (0xFFFFFFFFu << a);     // bits >= a
        uint32_t mb = b >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu << b);   // bits < b
        r.w[i] = ma & mb;
    }
    return r;
}
// B[i] = 1 iff A[i .. i+n-1] are all 1 (n >= 1): shift-AND doubling
__device__ __forceinline__ Bits bits_runs_ge(const Bits &a, int n) {
    Bits b = a;
    int len = 1;
    while (len < n) {
        int s = min(len, n - len);
        if (s >= 32) {            // only reachable for n > 32 (never with the reference's bit rates)
            for (int q = 0; q < s; q++) b = bits_and(b, bits_shr_small(b, 1));
        } else {
            b = bits_and(b, bits_shr_small(b, s));
        }
        len += s;
    }
    return b;
}
// length of the run of ones starting at bit `start` (A[start] must be 1; bit 128 counts as 0)
__device__ __forceinline__ int bits_run_length(const Bits &a, int start) {
    Bits z;
    Bits below = bits_range(0, start);
#pragma unroll
    for (int i = 0; i < NW; i++) z.w[i] = ~a.w[i] & ~below.w[i];
    int pos = bits_ffs(z);
    return (pos < 0 ? MAX_SLOTS : pos) - start;
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based traffic)
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// -ln((r + 0.5) / 2^32) with every rounding fixed (DESIGN.md "Traffic"): IEEE add/mul/div/fma only,
// so the CPU oracle reproduces it bit-for-bit.
__device__ __forceinline__ double neg_log_u32(uint32_t r) {
    double x = __dadd_rn((double)r, 0.5);
    long long bits = __double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    double z = __dmul_rn(s, s);
    double p = 1.0 / 23.0;
    p = __fma_rn(p, z, 1.0 / 21.0);
    p = __fma_rn(p, z, 1.0 / 19.0);
    p = __fma_rn(p, z, 1.0 / 17.0);
    p = __fma_rn(p, z, 1.0 / 15.0);
    p = __fma_rn(p, z, 1.0 / 13.0);
    p = __fma_rn(p, z, 1.0 / 11.0);
    p = __fma_rn(p, z, 1.0 / 9.0);
    p = __fma_rn(p, z, 1.0 / 7.0);
    p = __fma_rn(p, z, 1.0 / 5.0);
    p = __fma_rn(p, z, 1.0 / 3.0);
    p = __fma_rn(p, z, 1.0);
    double lnm = __dmul_rn(__dmul_rn(2.0, s), p);
    double lnx = __fma_rn((double)e, 0.6931471805599453, lnm);
    return __dadd_rn(22.180709777918249, -lnx);
}

// ---------------------------------------------------------------- per-env release-time heap (HD-ary, HBM)
// Replaces heapq in optical_network_env.py:143-154 / rmsa_env.py:591-597 (pop order by time is
// identical for any heap arity because release times are distinct).  Two parallel arrays per env:
// release times (f64) and payloads (u64).  Slots 0..HD-2 are unused so that the HD children of slot s
// start at HD*(s-HD+2): one aligned 8*HD-byte group, i.e. a level of a sift-down is ONE independent
// fetch.  With HD = 16, up to 272 live services need at most 2 levels.  Payload moves are deferred to
// the end of a pop so that they cost one round trip in total instead of one per level.
// (An unsorted table with a scan was measured 2x slower: profiles/r1_notes.md.)
#ifndef ORLG_HEAP_ARITY
#define ORLG_HEAP_ARITY 16
#endif
constexpr unsigned HD = ORLG_HEAP_ARITY;
constexpr unsigned HEAP_ROOT = HD - 1;
constexpr int HEAP_MAX_DEPTH = HD >= 16 ? 3 : 4;      // 16-ary: 4368 entries, 8-ary: 4680 entries
#define ORLG_INF __longlong_as_double(0x7ff0000000000000LL)

__device__ __forceinline__ unsigned heap_first_child(unsigned s) { return HD * (s - HD + 2); }
__device__ __forceinline__ unsigned heap_parent(unsigned c) { return c / HD + HD - 2; }

__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ void heap_push(double *ht, unsigned long long *hp, unsigned &n, double t,
                                          unsigned long long payload) {
    unsigned i = n + HEAP_ROOT;
    n++;
    while (i > HEAP_ROOT) {
        unsigned p = heap_parent(i);
        double pt = ht[p];
        if (pt <= t) break;
        ht[i] = pt;
        hp[i] = hp[p];
        i = p;
    }
    ht[i] = t;
    hp[i] = payload;
}

// pops the root; returns its payload and the new minimum time (+inf when empty)
__device__ __forceinline__ unsigned long long heap_pop(double *ht, unsigned long long *hp, unsigned &n, double &new_min) {
    const unsigned long long top = hp[HEAP_ROOT];
    n--;
    if (n == 0) {
        new_min = ORLG_INF;
        return top;
    }
    const unsigned end = n + HEAP_ROOT;      // valid slots [HEAP_ROOT, end); the old last element sits at `end`
    const double lt = ht[end];
    const unsigned long long lp = hp[end];
    unsigned i = HEAP_ROOT;
    unsigned mv_src[HEAP_MAX_DEPTH];
    int nm = 0;
    bool go = true;
    new_min = lt;
#pragma unroll
    for (int lev = 0; lev < HEAP_MAX_DEPTH; lev++) {
        mv_src[lev] = 0;
        const unsigned c0 = heap_first_child(i);
        if (go && c0 < end) {
            const double2 *g = reinterpret_cast<const double2 *>(ht + c0);
            double bt = ORLG_INF;
            unsigned bi = c0;
#pragma unroll
            for (unsigned q = 0; q < HD / 2; q++) {
                const double2 v = g[q];
                if (c0 + 2 * q < end && v.x < bt) { bt = v.x; bi = c0 + 2 * q; }
                if (c0 + 2 * q + 1 < end && v.y < bt) { bt = v.y; bi = c0 + 2 * q + 1; }
            }
            if (lt <= bt) {
                go = false;
            } else {
                ht[i] = bt;
                if (lev == 0) new_min = bt;
                mv_src[lev] = bi;          // payload of slot bi moves up into the slot visited at this level
                nm = lev + 1;
                i = bi;
            }
        } else {
            go = false;
        }
    }
    ht[i] = lt;
    // deferred payload moves: level 0 writes the root, level k writes mv_src[k-1]
    unsigned long long pv[HEAP_MAX_DEPTH];
#pragma unroll
    for (int lev = 0; lev < HEAP_MAX_DEPTH; lev++)
        if (lev < nm) pv[lev] = hp[mv_src[lev]];
#pragma unroll
    for (int lev = 0; lev < HEAP_MAX_DEPTH; lev++)
        if (lev < nm) hp[lev == 0 ? HEAP_ROOT : mv_src[lev - 1]] = pv[lev];
    hp[i] = lp;
    return top;
}

// payload: path row (20 bits) | start (9) | slots (8) | core (5) | service id (22)
__device__ __forceinline__ unsigned long long pack_service(int row, int start, int n, int core, int sid) {
    return (unsigned long long)(uint32_t)row | ((unsigned long long)(uint32_t)start << 20) |
           ((unsigned long long)(uint32_t)n << 29) | ((unsigned long long)(uint32_t)core << 37) |
           ((unsigned long long)((uint32_t)sid & 0x3FFFFFu) << 42);
}
__device__ __forceinline__ int svc_row(unsigned long long p) { return (int)(p & 0xFFFFFu); }
__device__ __forceinline__ int svc_start(unsigned long long p) { return (int)((p >> 20) & 0x1FFu); }
__device__ __forceinline__ int svc_slots(unsigned long long p) { return (int)((p >> 29) & 0xFFu); }
__device__ __forceinline__ int svc_core(unsigned long long p) { return (int)((p >> 37) & 0x1Fu); }
__device__ __forceinline__ int svc_id(unsigned long long p) { return (int)((p >> 42) & 0x3FFFFFu); }

}  // namespace orlg
