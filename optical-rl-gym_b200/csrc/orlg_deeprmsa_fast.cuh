// orlg_deeprmsa_fast.cuh -- latency-oriented DeepRMSA step kernel (the headline path).
//
// Same semantics as step_kernel<ORLG_DEEPRMSA> (orlg_kernels.cuh), reorganised so that one thread's
// step is ONE round trip to HBM plus the release-heap chain:
//   * all E link masks of the env (E x 16 B, coalesced across the warp) and the scalar block are
//     requested at kernel entry and stay in registers; allocation, releases and the candidate-path
//     AND-reduction are register-only, branch-free sweeps over the E links; only dirty links are
//     written back;
//   * the read-only topology tables (pair -> paths, path -> link bitmap / spectral efficiency,
//     slots-per-bit-rate, node CDF, f32 normalisation tables) are staged once per CTA in shared memory;
//   * the 5 candidate paths' block features are straight-line code (no data-dependent loops for j = 1),
//     so their instruction streams interleave;
//   * the observation tile of the CTA is staged in shared memory and written with full-line stores.
#pragma once
#include "orlg_kernels.cuh"

namespace orlg {

constexpr int FAST_THREADS = 64;
constexpr int FAST_MIN_BLOCKS = 7;     // 448 threads/SM x 148 SMs >= 65536 envs in one wave

__device__ __forceinline__ Bits bits_runs_ge_flat(const Bits &a, int n) {
    // shift-AND doubling without a data-dependent loop for n <= 16 (larger n: generic path)
    if (n > 16) return bits_runs_ge(a, n);
    Bits b = a;
    int len = 1;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        int s = max(min(len, n - len), 0);
        Bits sh = bits_shr_small(b, s);       // s == 0: funnel shift by 0 returns b
        b = bits_and(b, sh);
        len += s;
    }
    return b;
}

// mask of bits >= start (0 <= start <= 128)
__device__ __forceinline__ Bits bits_from_pos(int start) {
    Bits r;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        int a = min(max(start - 32 * i, 0), 32);
        r.w[i] = a >= 32 ? 0u : (0xFFFFFFFFu << a);
    }
    return r;
}

__device__ __forceinline__ int bits_run_length_flat(const Bits &a, int start) {
    Bits f = bits_from_pos(start);
    Bits z;
#pragma unroll
    for (int i = 0; i < NW; i++) z.w[i] = ~a.w[i] & f.w[i];
    int pos = bits_ffs(z);
    return (pos < 0 ? MAX_SLOTS : pos) - start;
}

template <int ET, int KM, int JT, bool OBS64>
__global__ void __launch_bounds__(FAST_THREADS, FAST_MIN_BLOCKS)
deeprmsa_fast_kernel(const Params p, const StepIO io, const int mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int env = blockIdx.x * FAST_THREADS + tid;
    const bool live = env < p.n;
    const int e = live ? env : p.n - 1;
    const int J = JT == 1 ? 1 : p.J;

    // ---------------- stage 1: every load whose address depends only on the env id
    uint4 M[ET];
    if (mode != MODE_FULL_RESET) {
        const uint4 *mr = p.masks + e;
#pragma unroll
        for (int l = 0; l < ET; l++) M[l] = mr[(size_t)l * p.n];
    } else {
        uint4 full = bits_to(bits_range(0, p.S));
#pragma unroll
        for (int l = 0; l < ET; l++) M[l] = full;
    }
    double now = p.now[e];
    double hold = p.cur_hold[e];
    uint2 rq = p.cur_req[e];
    long long cnt[8];
#pragma unroll
    for (int q = 0; q < 8; q++) cnt[q] = p.counters[(size_t)q * p.n + e];
    unsigned ridx = p.req_index[e];
    unsigned nheap = p.nheap[e];
    double hmin = p.heap_min[e];
    unsigned err = p.errors[e];
    const int act = (mode == MODE_STEP) ? io.actions[e] : -1;
    unsigned long long candw = 0;
    if (p.cand_stride == 8) candw = *reinterpret_cast<const unsigned long long *>(p.cand + (size_t)e * 8);
    double *ht = p.heap_time + (size_t)e * p.heap_cap;
    unsigned long long *hp = p.heap_pay + (size_t)e * p.heap_cap;

    // ---------------- topology tables -> shared memory (one copy per CTA)
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < p.tab_vec; i += FAST_THREADS) dst[i] = p.tab_blob[i];
    }
    const unsigned short *s_pair_first = reinterpret_cast<const unsigned short *>(smem + p.off_pair_first);
    const unsigned char *s_pair_count = smem + p.off_pair_count;
    const unsigned *s_path_lm = reinterpret_cast<const unsigned *>(smem + p.off_path_lm);
    const unsigned char *s_path_se = smem + p.off_path_se;
    const unsigned char *s_nslots = smem + p.off_nslots;              // [se][128]
    const unsigned *s_node_thr = reinterpret_cast<const unsigned *>(smem + p.off_node_thr);
    const float *s_pos = reinterpret_cast<const float *>(smem + p.off_pos);     // (2v - S) / S
    const float *s_nsl = reinterpret_cast<const float *>(smem + p.off_nsl);     // (2n - 11) / 7, n < 32
    unsigned char *stage = smem + p.tab_vec * 16;
    __syncthreads();

    int src = rq.x & 0xff, dst = (rq.x >> 8) & 0xff, br = (int)(rq.x >> 16), sid = (int)rq.y;
    bool accepted = false, done = false;
    int d_row = -1, d_start = -1, d_n = -1;
    unsigned dirty = 0;

    if (mode == MODE_FULL_RESET) {
        now = 0.0; nheap = 0; hmin = ORLG_INF; ridx = 0; err = 0;
        dirty = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < 8; q++) cnt[q] = 0;
    }

    if (live) {
        if (mode == MODE_STEP) {
            // ============ Phase A (deeprmsa_env.py:48-58 -> rmsa_env.py:163-209) ============
            if (hmin <= now + 4.0 * p.mean_iat) {       // a release is likely this step: warm the heap's first level
                prefetch_l2(ht + HD);
                prefetch_l2(hp + HEAP_ROOT);
            }
            const int pair = src * p.N + dst;
            const int first = s_pair_first[pair];
            const int npaths = s_pair_count[pair];
            if (act >= 0 && act < p.k * J) {
                const int route = JT == 1 ? act : act / J;
                if (route < npaths) {
                    unsigned st = p.cand_stride == 8 ? (unsigned)((candw >> (8 * act)) & 0xffu)
                                                     : (unsigned)p.cand[(size_t)e * p.cand_stride + act];
                    if (st != CAND_NONE) {
                        if (nheap + HEAP_ROOT + 1 > (unsigned)p.heap_cap) {
                            err |= ORLG_ERR_HEAP_OVERFLOW;
                        } else {
                            const int row = first + route;
                            const int se = s_path_se[row];
                            const int n = br < 128 ? s_nslots[se * 128 + br] : p.nslots[se * (p.br_max + 1) + br];
                            const unsigned lm = s_path_lm[row];
                            const Bits rm = bits_range((int)st, (int)st + n);
#pragma unroll
                            for (int l = 0; l < ET; l++) {           // _provision_path: clear [start, start+n) on the path's links
                                const unsigned sel = 0u - ((lm >> l) & 1u);
                                M[l].x &= ~(rm.w[0] & sel); M[l].y &= ~(rm.w[1] & sel);
                                M[l].z &= ~(rm.w[2] & sel); M[l].w &= ~(rm.w[3] & sel);
                            }
                            dirty |= lm;
                            const double rel = __dadd_rn(now, hold);
                            heap_push(ht, hp, nheap, rel, pack_service(row, (int)st, n, 0, sid));
                            hmin = fmin(hmin, rel);
                            cnt[1] += 1; cnt[3] += 1; cnt[5] += br; cnt[7] += br;
                            accepted = true;
                            d_row = row; d_start = (int)st; d_n = n;
                        }
                    }
                } else {
                    err |= ORLG_ERR_NO_SUCH_PATH;
                }
            }
            if (io.reward) io.reward[env] = accepted ? 1.0f : -1.0f;
            if (io.decision) {
                int *d = io.decision + (size_t)env * 6;
                d[0] = accepted; d[1] = d_row; d[2] = d_start; d[3] = d_n; d[4] = accepted ? 0 : -1; d[5] = -1;
            }
            if (io.info) {
#pragma unroll
                for (int q = 0; q < 8; q++) io.info[(size_t)env * 8 + q] = cnt[q];
            }
        }

        if (mode == MODE_STEP || mode == MODE_FULL_RESET) {
            // ============ Phase B: _next_service (rmsa_env.py:545-597) ============
            double arrival, holding;
            int nsrc, ndst, nbr;
            if (p.traffic == ORLG_TRAFFIC_PHILOX) {
                philox_request(p, s_node_thr, e, ridx, now, arrival, holding, nsrc, ndst, nbr);
            } else if ((long long)ridx < p.trace_len) {
                const orlg_request r = p.trace[(size_t)e * p.trace_len + ridx];
                arrival = r.arrival; holding = r.holding; nsrc = r.src; ndst = r.dst;
                nbr = min(max(r.bit_rate, 0), p.br_max);
            } else {
                err |= ORLG_ERR_TRACE_EXHAUSTED;
                arrival = now; holding = 0.0; nsrc = 0; ndst = 1; nbr = p.br_lo;
            }
            ridx++;
            now = arrival; hold = holding; src = nsrc; dst = ndst; br = nbr;
            sid = (int)cnt[2];
            cnt[0] += 1; cnt[2] += 1; cnt[4] += br; cnt[6] += br;
            while (nheap > 0 && hmin <= now) {            // release loop, on the register-resident masks
                const unsigned long long pl = heap_pop(ht, hp, nheap, hmin);
                const unsigned lm = s_path_lm[svc_row(pl)];
                const int rs = svc_start(pl);
                const Bits rm = bits_range(rs, rs + svc_slots(pl));
#pragma unroll
                for (int l = 0; l < ET; l++) {
                    const unsigned sel = 0u - ((lm >> l) & 1u);
                    M[l].x |= rm.w[0] & sel; M[l].y |= rm.w[1] & sel;
                    M[l].z |= rm.w[2] & sel; M[l].w |= rm.w[3] & sel;
                }
                dirty |= lm;
            }
            done = (cnt[2] == (long long)p.episode_length);
        }

        if (mode == MODE_EPISODE_RESET || (mode == MODE_STEP && done && p.auto_reset)) {
            cnt[2] = 1; cnt[3] = 0; cnt[6] = br; cnt[7] = 0;      // rmsa_env.py:285-330
        }

        // ============ Phase C: observation of the pending request (deeprmsa_env.py:60-121) ============
        const int pair = src * p.N + dst;
        const int first = s_pair_first[pair];
        const int npaths = min((int)s_pair_count[pair], KM);
        unsigned lms[KM];
        int ns[KM];
        Bits A[KM];
#pragma unroll
        for (int q = 0; q < KM; q++) {
            A[q] = bits_ones();
            lms[q] = 0; ns[q] = 1;
            if (q < npaths) {
                lms[q] = s_path_lm[first + q];
                const int se = s_path_se[first + q];
                ns[q] = br < 128 ? s_nslots[se * 128 + br] : p.nslots[se * (p.br_max + 1) + br];
            }
        }
#pragma unroll
        for (int l = 0; l < ET; l++) {
#pragma unroll
            for (int q = 0; q < KM; q++) {
                const unsigned keep = ((lms[q] >> l) & 1u) - 1u;
                A[q].w[0] &= M[l].x | keep; A[q].w[1] &= M[l].y | keep;
                A[q].w[2] &= M[l].z | keep; A[q].w[3] &= M[l].w | keep;
            }
        }
        {   // write back the links this step touched
            uint4 *mw = p.masks + env;
#pragma unroll
            for (int l = 0; l < ET; l++)
                if ((dirty >> l) & 1u) mw[(size_t)l * p.n] = M[l];
        }

        const int W = 2 * J + 3;
        const bool want_obs = io.obs != nullptr;
        float *so32 = reinterpret_cast<float *>(stage) + (size_t)tid * p.obs_dim;
        double *so64 = reinterpret_cast<double *>(stage) + (size_t)tid * p.obs_dim;
        if (want_obs) {
            const int lo = min(src, dst), hi = max(src, dst);
            if (OBS64) {
                for (int q = 0; q < p.obs_dim; q++) so64[q] = (q > 2 * p.N) ? -1.0 : 0.0;
                so64[0] = __ddiv_rn((double)br, 100.0);
                so64[1 + lo] = 1.0; so64[1 + p.N + hi] = 1.0;
            } else {
                for (int q = 0; q < p.obs_dim; q++) so32[q] = (q > 2 * p.N) ? -1.0f : 0.0f;
                so32[0] = __fdiv_rn((float)br, 100.0f);
                so32[1 + lo] = 1.0f; so32[1 + p.N + hi] = 1.0f;
            }
        }
        unsigned long long cand_out = 0xFFFFFFFFFFFFFFFFULL;
#pragma unroll
        for (int q = 0; q < KM; q++) {
            if (q < npaths) {
                const int n = ns[q];
                const Bits B = bits_runs_ge_flat(A[q], n);
                Bits starts = bits_andnot(B, bits_shl1(B));
                const int total = bits_popc(A[q]);
                const int runs = bits_popc(bits_andnot(A[q], bits_shl1(A[q])));
                const int ob = 1 + 2 * p.N + q * W;
                if (JT == 1) {
                    const int st = bits_ffs(starts);
                    const int len = bits_run_length_flat(A[q], max(st, 0));
                    if (st >= 0) cand_out = (cand_out & ~(0xFFULL << (8 * q))) | ((unsigned long long)st << (8 * q));
                    if (want_obs && st >= 0) {
                        if (OBS64) {
                            so64[ob] = __ddiv_rn(__dmul_rn(2.0, __dadd_rn((double)st, -__dmul_rn(0.5, (double)p.S))), (double)p.S);
                            so64[ob + 1] = __ddiv_rn(__dadd_rn((double)len, -8.0), 8.0);
                        } else {
                            so32[ob] = s_pos[st];
                            so32[ob + 1] = (float)(len - 8) * 0.125f;
                        }
                    }
                    if (io.obs_int) {
                        int *oi = io.obs_int + ((size_t)env * p.k + q) * W;
                        oi[0] = st; oi[1] = st >= 0 ? len : -1;
                    }
                } else {
                    for (int b = 0; b < J; b++) {
                        const int st = bits_ffs(starts);
                        if (p.cand_stride == 8) {
                            if (st >= 0) cand_out = (cand_out & ~(0xFFULL << (8 * (q * J + b)))) | ((unsigned long long)st << (8 * (q * J + b)));
                        } else {
                            p.cand[(size_t)env * p.cand_stride + q * J + b] = (unsigned char)(st < 0 ? CAND_NONE : st);
                        }
                        int len = -1;
                        if (st >= 0) {
                            starts = bits_clear_lowest(starts);
                            len = bits_run_length_flat(A[q], st);
                            if (want_obs) {
                                if (OBS64) {
                                    so64[ob + 2 * b] = __ddiv_rn(__dmul_rn(2.0, __dadd_rn((double)st, -__dmul_rn(0.5, (double)p.S))), (double)p.S);
                                    so64[ob + 2 * b + 1] = __ddiv_rn(__dadd_rn((double)len, -8.0), 8.0);
                                } else {
                                    so32[ob + 2 * b] = s_pos[st];
                                    so32[ob + 2 * b + 1] = (float)(len - 8) * 0.125f;
                                }
                            }
                        }
                        if (io.obs_int) {
                            int *oi = io.obs_int + ((size_t)env * p.k + q) * W;
                            oi[2 * b] = st; oi[2 * b + 1] = len;
                        }
                    }
                }
                if (want_obs) {
                    if (OBS64) {
                        so64[ob + 2 * J] = __ddiv_rn(__dadd_rn((double)n, -5.5), 3.5);
                        so64[ob + 2 * J + 1] = __ddiv_rn(__dmul_rn(2.0, __dadd_rn((double)total, -__dmul_rn(0.5, (double)p.S))), (double)p.S);
                        if (runs > 0) so64[ob + 2 * J + 2] = __ddiv_rn(__dadd_rn(__ddiv_rn((double)total, (double)runs), -4.0), 4.0);
                    } else {
                        so32[ob + 2 * J] = n < 32 ? s_nsl[n] : __fdiv_rn((float)(2 * n - 11), 7.0f);
                        so32[ob + 2 * J + 1] = s_pos[total];
                        if (runs > 0) so32[ob + 2 * J + 2] = __fdiv_rn((float)(total - 4 * runs), (float)(4 * runs));
                    }
                }
                if (io.obs_int) {
                    int *oi = io.obs_int + ((size_t)env * p.k + q) * W;
                    oi[2 * J] = n; oi[2 * J + 1] = total; oi[2 * J + 2] = runs;
                }
            }
        }
        if (p.cand_stride == 8) {
            *reinterpret_cast<unsigned long long *>(p.cand + (size_t)env * 8) = cand_out;
        } else {
            for (int q = npaths * J; q < p.k * J; q++) p.cand[(size_t)env * p.cand_stride + q] = (unsigned char)CAND_NONE;
        }
        if (io.obs_int)
            for (int q = npaths * W; q < p.k * W; q++) io.obs_int[(size_t)env * p.k * W + q] = -1;
    }

    if (io.obs != nullptr) {
        __syncthreads();
        const size_t tile0 = (size_t)blockIdx.x * FAST_THREADS * p.obs_dim;
        const int rows = min(FAST_THREADS, p.n - blockIdx.x * FAST_THREADS);
        const int total_el = rows * p.obs_dim;
        if (OBS64) {
            const double *s = reinterpret_cast<const double *>(stage);
            double *g = reinterpret_cast<double *>(io.obs) + tile0;
            for (int q = tid; q < total_el; q += FAST_THREADS) g[q] = s[q];
        } else {
            const float *s = reinterpret_cast<const float *>(stage);
            float *g = reinterpret_cast<float *>(io.obs) + tile0;
            for (int q = tid; q < total_el; q += FAST_THREADS) g[q] = s[q];
        }
    }

    if (live && mode != MODE_OBSERVE) {
        p.now[env] = now;
        p.cur_hold[env] = hold;
        p.cur_req[env] = make_uint2((unsigned)src | ((unsigned)dst << 8) | ((unsigned)br << 16), (unsigned)sid);
#pragma unroll
        for (int q = 0; q < 8; q++) p.counters[(size_t)q * p.n + env] = cnt[q];
        p.req_index[env] = ridx;
        p.nheap[env] = nheap;
        p.heap_min[env] = hmin;
        p.errors[env] = err;
        if (mode == MODE_STEP && io.done) io.done[env] = done ? 1 : 0;
    }
}

}  // namespace orlg
