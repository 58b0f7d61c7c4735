/*
 * orlg_oracle.c -- CPU restatement of Optical RL-Gym's step hot path.  TEST INFRASTRUCTURE.
 *
 * This file is the ORACLE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it.  The product (liborlg.so, the
 * optical_rl_gym_b200 package) never links, imports or executes it.
 *
 * It restates, literally and on plain per-slot int arrays (NOT the product's
 * bit-packed layout, so that it is an independent check), the reference files
 * (paths relative to /root/reference):
 *   optical_rl_gym/envs/optical_network_env.py:76-94,143-173,181-210
 *   optical_rl_gym/envs/rmsa_env.py:163-359,364-437,545-697,747-803
 *   optical_rl_gym/envs/deeprmsa_env.py:48-155
 *   optical_rl_gym/envs/rwa_env.py:101-208,258-349,385-502
 *   optical_rl_gym/envs/rmcsa_env.py:209-384,386-483,488-558,690-794,882-911
 * Parity pinning: tests/test_oracle_golden.py replays tests/golden/<case>.npz (recorded from
 * the live reference by tests/golden/make_golden.py) through this file bit-for-bit.
 *
 * Request sources: (a) a recorded trace (arrival, holding, src, dst, bit_rate), which
 * is how parity with the reference (MT19937 traffic) is established; (b) the product's
 * counter-based Philox traffic (DESIGN.md "Traffic"), restated here independently of
 * the CUDA code so the two can be compared bit-for-bit.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { KIND_RWA = 0, KIND_RMSA = 1, KIND_DEEPRMSA = 2, KIND_RMCSA = 3 };

typedef struct {
    int32_t kind;
    int32_t num_nodes, num_links, k_paths, num_paths;
    int32_t num_slots;          /* num_spectrum_resources */
    int32_t num_cores;          /* num_spatial_resources (1 unless RMCSA) */
    int32_t num_mods;
    int32_t j;                  /* DeepRMSA blocks per path */
    int32_t episode_length;
    int32_t allow_rejection;
    int32_t bit_rate_lo, bit_rate_hi;   /* continuous mode randint bounds */
    int32_t num_bit_rates;              /* >0: discrete mode (Philox traffic only) */
    int32_t stats;                      /* 1: maintain the float statistics of info (row f1) */
    double channel_width;
    double mean_holding, mean_iat;
    double worst_xt;                    /* RMCSA, before the +4 dB margin */
} ocfg_t;

typedef struct {
    const int32_t *pair_first, *pair_count;   /* [N*N] */
    const int32_t *path_hops, *path_se, *path_mod, *path_link_ptr, *path_links;
    const double *path_length;
    const int32_t *mod_se;
    const double *mod_osnr, *mod_xt;
    const double *node_prob;                  /* [N] node_request_probabilities */
    const int32_t *bit_rates;                 /* [num_bit_rates] */
    const double *bit_rate_prob;
    const int32_t *link_order;                /* [E] link indices in topology.edges() iteration order */
} otab_t;

typedef struct {
    int32_t id, src, dst, bit_rate, path_row, initial_slot, number_slots, core, mod, accepted;
    double arrival, holding;
} service_t;

typedef struct { double t; service_t s; } event_t;

typedef struct {
    ocfg_t c;
    otab_t t;
    /* owned copies of the tables */
    void *own[16];
    int nown;
    int8_t *avail;      /* [C][E][S] 1 = free  (graph["available_slots"/"available_wavelengths"]) */
    int32_t *alloc;     /* [C][E][S] service id or -1 */
    event_t *heap; int nheap, capheap;
    double now;
    service_t cur;
    int new_service;
    int64_t processed, accepted, ep_processed, ep_accepted;
    int64_t br_req, br_prov, ep_br_req, ep_br_prov;
    int64_t req_index;          /* requests generated since the last full reset */
    /* traffic */
    int traffic;                /* 0 = trace, 1 = philox */
    const double *tr_arr, *tr_hold; const int32_t *tr_src, *tr_dst, *tr_br; int64_t tr_len;
    uint64_t seed; uint32_t env_id;
    uint32_t src_thr[256];      /* integer CDF thresholds for Philox src/dst draws */
    uint32_t br_thr[64];
    int error;
    /* float statistics of row f1 (rmsa_env.py:439-543, 699-744) */
    double *link_util, *link_comp, *link_last;   /* [E] time-averaged utilisation / compactness, last update time */
    double *link_frag;                            /* [E] time-averaged external fragmentation (rmsa_env.py:487-524) */
    double g_thr, g_comp, g_last;                 /* topology.graph["throughput" / "compactness" / "last_update"] (rmsa_env.py:439-462) */
    long run_br;                                  /* sum of bit_rate over graph["running_services"] */
    long sum_nh;                                  /* sum over running services of number_slots * hops */
    /* discrete bit-rate selection (rmsa_env.py:88-110): bit_rate_requested_histogram / bit_rate_provisioned_histogram,
       and what step() derives from them for `info` (rmsa_env.py:217-227, 268-273) */
    int64_t hist_req[64], hist_prov[64];
    double br_blocking[65];                       /* per bit rate (order of bit_rates), then fairness */
    /* RWA actions_output (rwa_env.py:52-58, 103): only its marginals reach `info` (rwa_env.py:148-151) */
    int64_t act_rows[32], act_cols[520], act_total;
} oenv_t;

static int br_slot(const oenv_t *e, int bit_rate) {
    for (int i = 0; i < e->c.num_bit_rates; i++)
        if (e->t.bit_rates[i] == bit_rate) return i;
    return -1;
}

/* ------------------------------------------------------------------ tables */
static void *dup_mem(oenv_t *e, const void *p, size_t n) {
    void *q = malloc(n ? n : 1);
    if (p && n) memcpy(q, p, n);
    e->own[e->nown++] = q;
    return q;
}

/* ------------------------------------------------------------------ Philox traffic */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* -ln((r+0.5)/2^32): only IEEE-754 +,-,*,/ and fma, in a fixed order (DESIGN.md "Traffic") */
static double neg_log_u32(uint32_t r) {
    double x = (double)r + 0.5;
    uint64_t bits; memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m; memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    p = fma(p, z, 1.0);
    double lnm = (2.0 * s) * p;
    double lnx = fma((double)e, 0.6931471805599453, lnm);
    return 22.180709777918249 - lnx;
}

/* thr[i] = round(cdf_i * 2^32) clamped, last forced to 2^32-1 sentinel semantics below */
static void build_thresholds(const double *p, int n, uint32_t *thr) {
    double tot = 0.0, acc = 0.0;
    for (int i = 0; i < n; i++) tot += p[i];
    for (int i = 0; i < n; i++) {
        acc += p[i];
        double v = floor(acc / tot * 4294967296.0 + 0.5);
        if (v > 4294967295.0) v = 4294967295.0;
        thr[i] = (uint32_t)v;
    }
    thr[n - 1] = 4294967295u;
}

static int pick_thr(const uint32_t *thr, int n, uint32_t r) {
    /* first i with r <= thr[i] when r is compared as "r < thr" except for the sentinel */
    for (int i = 0; i < n - 1; i++) if (r < thr[i]) return i;
    return n - 1;
}

static void philox_request(oenv_t *e, service_t *s) {
    uint32_t c[4] = { (uint32_t)e->req_index, (uint32_t)((uint64_t)e->req_index >> 32), e->env_id, 0u };
    philox4x32_10(c, (uint32_t)e->seed, (uint32_t)(e->seed >> 32));
    double iat = neg_log_u32(c[0]) * e->c.mean_iat;
    s->arrival = e->now + iat;
    s->holding = neg_log_u32(c[1]) * e->c.mean_holding;
    int n = e->c.num_nodes;
    int src = pick_thr(e->src_thr, n, c[2]);
    /* dst ~ p with p[src] removed: scale c[3] into the remaining mass, skip over src's interval */
    uint32_t lo = src ? e->src_thr[src - 1] : 0u;
    uint64_t hi = (src == n - 1) ? 4294967296ULL : (uint64_t)e->src_thr[src];
    uint64_t mass = hi - lo;
    uint64_t rem = 4294967296ULL - mass;
    uint64_t tt = ((uint64_t)c[3] * rem) >> 32;
    if (tt >= lo) tt += mass;
    int dst = n - 1;
    for (int i = 0; i < n - 1; i++) if (tt < (uint64_t)e->src_thr[i]) { dst = i; break; }
    if (dst == src) dst = (src + 1) % n;       /* unreachable unless p is degenerate */
    s->src = src; s->dst = dst;
    s->bit_rate = 0;
    if (e->c.kind != KIND_RWA) {
        uint32_t d[4] = { (uint32_t)e->req_index, (uint32_t)((uint64_t)e->req_index >> 32), e->env_id, 1u };
        philox4x32_10(d, (uint32_t)e->seed, (uint32_t)(e->seed >> 32));
        if (e->c.num_bit_rates > 0) {
            s->bit_rate = e->t.bit_rates[pick_thr(e->br_thr, e->c.num_bit_rates, d[0])];
        } else {
            uint32_t span = (uint32_t)(e->c.bit_rate_hi - e->c.bit_rate_lo + 1);
            s->bit_rate = e->c.bit_rate_lo + (int32_t)(((uint64_t)d[0] * span) >> 32);
        }
    }
}

static void draw_request(oenv_t *e, service_t *s) {
    memset(s, 0, sizeof(*s));
    s->path_row = -1; s->initial_slot = -1; s->core = -1; s->mod = -1;
    if (e->traffic == 0) {
        if (e->req_index >= e->tr_len) { e->error |= 1; s->arrival = e->now; s->holding = 0.0; s->src = 0; s->dst = 1; s->bit_rate = e->c.bit_rate_lo; return; }
        s->arrival = e->tr_arr[e->req_index];
        s->holding = e->tr_hold[e->req_index];
        s->src = e->tr_src[e->req_index];
        s->dst = e->tr_dst[e->req_index];
        s->bit_rate = e->tr_br ? e->tr_br[e->req_index] : 0;
    } else {
        philox_request(e, s);
    }
}

/* ------------------------------------------------------------------ event heap (heapq, optical_network_env.py:143-154) */
static void heap_push(oenv_t *e, double t, const service_t *s) {
    if (e->nheap == e->capheap) {
        e->capheap = e->capheap ? e->capheap * 2 : 256;
        e->heap = (event_t *)realloc(e->heap, sizeof(event_t) * (size_t)e->capheap);
    }
    int i = e->nheap++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (e->heap[p].t <= t) break;
        e->heap[i] = e->heap[p];
        i = p;
    }
    e->heap[i].t = t; e->heap[i].s = *s;
}

static event_t heap_pop(oenv_t *e) {
    event_t top = e->heap[0];
    event_t last = e->heap[--e->nheap];
    int i = 0, n = e->nheap;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && e->heap[c + 1].t < e->heap[c].t) c++;
        if (last.t <= e->heap[c].t) break;
        e->heap[i] = e->heap[c];
        i = c;
    }
    if (n > 0) e->heap[i] = last;
    return top;
}

/* ------------------------------------------------------------------ helpers */
#define AV(e_, co_, l_, s_) ((e_)->avail[((size_t)(co_) * (e_)->c.num_links + (l_)) * (e_)->c.num_slots + (s_)])
#define AL(e_, co_, l_, s_) ((e_)->alloc[((size_t)(co_) * (e_)->c.num_links + (l_)) * (e_)->c.num_slots + (s_)])

static int pair_row(const oenv_t *e, int src, int dst, int path) {
    int key = src * e->c.num_nodes + dst;
    if (path >= e->t.pair_count[key]) return -1;     /* the reference would raise IndexError */
    return e->t.pair_first[key] + path;
}

/* rmsa_env.py:610-621 / rmcsa_env.py:753-765: math.ceil(bit_rate / (SE * channel_width)) + 1 */
static int number_slots(const oenv_t *e, int bit_rate, int se) {
    return (int)ceil((double)bit_rate / ((double)se * e->c.channel_width)) + 1;
}

/* rmsa_env.py:623-636, rmcsa_env.py:767-794, rwa_env.py:385-400 */
static int is_path_free(const oenv_t *e, int row, int core, int initial_slot, int n) {
    if (initial_slot + n > e->c.num_slots) return 0;
    for (int h = e->t.path_link_ptr[row]; h < e->t.path_link_ptr[row + 1]; h++) {
        int l = e->t.path_links[h];
        for (int s = initial_slot; s < initial_slot + n; s++)
            if (AV(e, core, l, s) == 0) return 0;
    }
    return 1;
}

static int rle(const int8_t *a, int n, int *starts, int *values, int *lengths);

/* rmsa_env.py:651-665 on one link row */
static int rle_link(const oenv_t *e, int core, int l, int lo, int hi, int *starts, int *values, int *lengths) {
    int8_t row[1024];
    for (int s = lo; s < hi; s++) row[s - lo] = AV(e, core, l, s);
    return rle(row, hi - lo, starts, values, lengths);
}

/* _update_link_stats: rmsa_env.py:464-543, rmcsa_env.py:591-688 (the slot row of the core being touched; the statistics are per
   LINK, shared by the cores), rwa_env.py:365-383 (utilisation only) */
static void update_link_stats(oenv_t *e, int core, int l) {
    int S = e->c.num_slots;
    double last_update = e->link_last[l];
    double time_diff = e->now - e->link_last[l];
    if (e->now > 0) {
        double last_util = e->link_util[l];
        long free_sum = 0;
        for (int s = 0; s < S; s++) free_sum += AV(e, core, l, s);
        double cur_util = (double)(S - free_sum) / (double)S;
        e->link_util[l] = ((last_util * last_update) + (cur_util * time_diff)) / e->now;
        if (e->c.kind != KIND_RWA) {
            double last_external_fragmentation = e->link_frag[l];
            double last_compactness = e->link_comp[l];
            double cur_external_fragmentation = 0.0;
            double cur_link_compactness = 0.0;
            if (free_sum > 0) {
                int st[1024], va[1024], le[1024];
                int nr = rle_link(e, core, l, 0, S, st, va, le);
                /* unused_blocks = indices of the free runs; max_empty stays 0 for fewer than two of them, and when they are
                   exactly the first and the last block of the link (unused_blocks != [0, len(values) - 1]) */
                int n_unused = 0, u_first = -1, u_last = -1, longest = 0;
                for (int r = 0; r < nr; r++) if (va[r] == 1) { if (u_first < 0) u_first = r; u_last = r; n_unused++; if (le[r] > longest) longest = le[r]; }
                int max_empty = 0;
                if (n_unused > 1 && !(n_unused == 2 && u_first == 0 && u_last == nr - 1)) max_empty = longest;
                cur_external_fragmentation = 1.0 - ((double)max_empty / (double)free_sum);
                int n_used = 0, first = -1, last = -1;
                for (int r = 0; r < nr; r++) if (va[r] == 0) { if (first < 0) first = r; last = r; n_used++; }
                if (n_used > 1) {
                    int lambda_min = st[first], lambda_max = st[last] + le[last];
                    int st2[1024], va2[1024], le2[1024];
                    int nr2 = rle_link(e, core, l, lambda_min, lambda_max, st2, va2, le2);
                    long unused_spectrum_slots = 0;                 /* np.sum(1 - internal_values) */
                    for (int r = 0; r < nr2; r++) unused_spectrum_slots += 1 - va2[r];
                    if (unused_spectrum_slots > 0)
                        cur_link_compactness = ((double)(lambda_max - lambda_min) / (double)(S - free_sum)) *
                                               (1 / (double)unused_spectrum_slots);
                    else cur_link_compactness = 1.0;
                } else cur_link_compactness = 1.0;
            }
            e->link_frag[l] = ((last_external_fragmentation * last_update) + (cur_external_fragmentation * time_diff)) / e->now;
            e->link_comp[l] = ((last_compactness * last_update) + (cur_link_compactness * time_diff)) / e->now;
        }
    }
    e->link_last[l] = e->now;
}

/* _get_network_compactness: rmsa_env.py:699-744; rmcsa_env.py:825-871 looks at the links of ONE core (but sums
   number_slots * hops over all running services) */
static double network_compactness_core(const oenv_t *e, int core) {
    int S = e->c.num_slots;
    long sum_occupied = 0, sum_unused_spectrum_blocks = 0;
    for (int l = 0; l < e->c.num_links; l++) {
        int st[1024], va[1024], le[1024];
        int nr = rle_link(e, core, l, 0, S, st, va, le);
        int n_used = 0, first = -1, last = -1;
        for (int r = 0; r < nr; r++) if (va[r] == 0) { if (first < 0) first = r; last = r; n_used++; }
        if (n_used > 1) {
            int lambda_min = st[first], lambda_max = st[last] + le[last];
            sum_occupied += lambda_max - lambda_min;
            int st2[1024], va2[1024], le2[1024];
            int nr2 = rle_link(e, core, l, lambda_min, lambda_max, st2, va2, le2);
            for (int r = 0; r < nr2; r++) sum_unused_spectrum_blocks += va2[r];
        }
    }
    if (sum_unused_spectrum_blocks > 0)
        return ((double)sum_occupied / (double)e->sum_nh) * ((double)e->c.num_links / (double)sum_unused_spectrum_blocks);
    return 1.0;
}
static double network_compactness(const oenv_t *e) { return network_compactness_core(e, 0); }

/* _update_network_stats (rmsa_env.py:439-462, rmcsa_env.py:560-589; a no-op in rwa_env.py:351-363): called by
   _provision_path only, after the new service joined graph["running_services"] */
static void update_network_stats(oenv_t *e, int core) {
    double last_update = e->g_last;
    double time_diff = e->now - last_update;
    if (e->now > 0) {
        double last_throughput = e->g_thr, last_compactness = e->g_comp;
        double cur_throughput = (double)e->run_br;          /* 0.0 + integer bit rates: exact */
        e->g_thr = ((last_throughput * last_update) + (cur_throughput * time_diff)) / e->now;
        e->g_comp = ((last_compactness * last_update) + (network_compactness_core(e, core) * time_diff)) / e->now;
    }
    e->g_last = e->now;
}

/* np.mean of a float64 list: numpy's pairwise summation (8 partial sums for n <= 128, recursive above) / n */
static double np_sum(const double *a, int n) {
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2; n2 -= n2 % 8;
    return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

static double mean_over_edges(const oenv_t *e, const double *per_link) {
    double tmp[4096];
    for (int i = 0; i < e->c.num_links; i++) tmp[i] = per_link[e->t.link_order[i]];
    return np_sum(tmp, e->c.num_links) / (double)e->c.num_links;
}

/* rmsa_env.py:364-415 (masks, allocation ids, per-link float stats) */
static void provision(oenv_t *e, int row, int core, int initial_slot, int n) {
    for (int h = e->t.path_link_ptr[row]; h < e->t.path_link_ptr[row + 1]; h++) {
        int l = e->t.path_links[h];
        for (int s = initial_slot; s < initial_slot + n; s++) { AV(e, core, l, s) = 0; AL(e, core, l, s) = e->cur.id; }
        if (e->c.stats) update_link_stats(e, core, l);
    }
    e->sum_nh += (long)n * (e->t.path_link_ptr[row + 1] - e->t.path_link_ptr[row]);
    e->run_br += e->cur.bit_rate;
    e->cur.path_row = row; e->cur.initial_slot = initial_slot; e->cur.number_slots = n; e->cur.core = core;
    if (e->c.stats && e->c.kind != KIND_RWA) update_network_stats(e, core);
}

/* rmsa_env.py:417-437 */
static void release(oenv_t *e, const service_t *s) {
    for (int h = e->t.path_link_ptr[s->path_row]; h < e->t.path_link_ptr[s->path_row + 1]; h++) {
        int l = e->t.path_links[h];
        for (int k = s->initial_slot; k < s->initial_slot + s->number_slots; k++) { AV(e, s->core, l, k) = 1; AL(e, s->core, l, k) = -1; }
        if (e->c.stats) update_link_stats(e, s->core, l);
    }
    e->run_br -= s->bit_rate;
    e->sum_nh -= (long)s->number_slots * (e->t.path_link_ptr[s->path_row + 1] - e->t.path_link_ptr[s->path_row]);
}

static void release_due(oenv_t *e) {
    /* rmsa_env.py:591-597: pop; if due release, else push back and stop */
    while (e->nheap > 0) {
        if (e->heap[0].t <= e->now) { event_t ev = heap_pop(e); release(e, &ev.s); }
        else break;
    }
}

/* rmsa_env.py:545-597, rwa_env.py:258-288, rmcsa_env.py:690-739 */
static void next_service(oenv_t *e) {
    if (e->new_service) return;
    service_t s;
    draw_request(e, &s);
    e->req_index++;
    e->now = s.arrival;
    if (e->c.kind == KIND_RMSA || e->c.kind == KIND_DEEPRMSA) {
        s.id = (int32_t)e->ep_processed;
        e->cur = s; e->new_service = 1;
        e->processed++; e->ep_processed++;
        e->br_req += s.bit_rate; e->ep_br_req += s.bit_rate;
        if (e->c.kind == KIND_RMSA && br_slot(e, s.bit_rate) >= 0) e->hist_req[br_slot(e, s.bit_rate)]++;   /* rmsa_env.py:579-581 */
        release_due(e);
    } else if (e->c.kind == KIND_RWA) {
        release_due(e);
        s.id = (int32_t)e->ep_processed; s.number_slots = 1;
        e->cur = s; e->new_service = 1;
    } else {
        release_due(e);
        s.id = (int32_t)e->ep_processed;
        e->cur = s; e->new_service = 1;
        e->br_req += s.bit_rate; e->ep_br_req += s.bit_rate;
    }
}

/* rmsa_env.py:638-649: product over the path's links */
static void available_slots(const oenv_t *e, int row, int core, int8_t *out) {
    int S = e->c.num_slots;
    for (int s = 0; s < S; s++) out[s] = 1;
    for (int h = e->t.path_link_ptr[row]; h < e->t.path_link_ptr[row + 1]; h++) {
        int l = e->t.path_links[h];
        for (int s = 0; s < S; s++) out[s] = (int8_t)(out[s] * AV(e, core, l, s));
    }
}

/* rmsa_env.py:651-665: run-length encoding -> (starts, values, lengths) */
static int rle(const int8_t *a, int n, int *starts, int *values, int *lengths) {
    int nr = 0, i = 0;
    while (i < n) {
        int j = i;
        while (j + 1 < n && a[j + 1] == a[i]) j++;
        starts[nr] = i; values[nr] = a[i]; lengths[nr] = j - i + 1; nr++;
        i = j + 1;
    }
    return nr;
}

/* rmsa_env.py:667-697: first j free runs with length >= slots -> (initial index, full run length) */
static int available_blocks(const oenv_t *e, int row, int n, int *bstart, int *blen) {
    int S = e->c.num_slots;
    int8_t av[1024]; int st[1024], va[1024], le[1024];
    available_slots(e, row, 0, av);
    int nr = rle(av, S, st, va, le), nb = 0;
    for (int r = 0; r < nr && nb < e->c.j; r++)
        if (va[r] == 1 && le[r] >= n) { bstart[nb] = st[r]; blen[nb] = le[r]; nb++; }
    return nb;
}

/* rmcsa_env.py:341-384 */
static int crosstalk_ok(const oenv_t *e, int mod, double path_length, int bit_rate) {
    double average_power = 1, nf_db = 5.5;
    double nf = pow(10.0, nf_db / 10.0);
    double amp_spam = 100, amp_gain_db = 20;
    double amp_gain = pow(10.0, amp_gain_db / 10.0);
    double lambda_ = 1550, h = 6.626068e-34;
    double f_hz = 2.99e8 / (lambda_ * 1e-9);
    double inband_xt = e->t.mod_xt[mod] + 4;          /* rmcsa_env.py:127-129 */
    double worst_xt = e->c.worst_xt + 4;
    double snr_min = pow(10.0, (e->t.mod_osnr[mod] + 2) / 10);
    double lmax_snr = (average_power * amp_spam) /
        (snr_min * h * f_hz * amp_gain * nf * ((double)bit_rate / (double)e->t.mod_se[mod]) * 1e9);
    lmax_snr = lmax_snr / 1000;
    double lmax_xt = pow(10.0, (inband_xt - worst_xt - 4) / 10);
    return (path_length < lmax_xt && path_length < lmax_snr) ? 1 : 0;
}

/* ------------------------------------------------------------------ public API */
typedef struct {
    int32_t accepted, path_row, initial_slot, number_slots, core, mod, service_id;
    int32_t done;
    double reward;
    int64_t info_counters[8];   /* counters when the reference builds `info` (before _next_service) */
    double stats[4];            /* network_compactness, its difference, avg link compactness, avg link utilisation */
} ostep_t;

void *oracle_create(const ocfg_t *cfg, const otab_t *tab) {
    oenv_t *e = (oenv_t *)calloc(1, sizeof(oenv_t));
    e->c = *cfg;
    int N = cfg->num_nodes, P = cfg->num_paths;
    e->t.pair_first = dup_mem(e, tab->pair_first, sizeof(int32_t) * N * N);
    e->t.pair_count = dup_mem(e, tab->pair_count, sizeof(int32_t) * N * N);
    e->t.path_hops = dup_mem(e, tab->path_hops, sizeof(int32_t) * P);
    e->t.path_se = dup_mem(e, tab->path_se, sizeof(int32_t) * P);
    e->t.path_mod = dup_mem(e, tab->path_mod, sizeof(int32_t) * P);
    e->t.path_link_ptr = dup_mem(e, tab->path_link_ptr, sizeof(int32_t) * (P + 1));
    e->t.path_links = dup_mem(e, tab->path_links, sizeof(int32_t) * tab->path_link_ptr[P]);
    e->t.path_length = dup_mem(e, tab->path_length, sizeof(double) * P);
    e->t.mod_se = dup_mem(e, tab->mod_se, sizeof(int32_t) * cfg->num_mods);
    e->t.mod_osnr = dup_mem(e, tab->mod_osnr, sizeof(double) * cfg->num_mods);
    e->t.mod_xt = dup_mem(e, tab->mod_xt, sizeof(double) * cfg->num_mods);
    e->t.node_prob = dup_mem(e, tab->node_prob, sizeof(double) * N);
    e->t.bit_rates = dup_mem(e, tab->bit_rates, sizeof(int32_t) * cfg->num_bit_rates);
    e->t.bit_rate_prob = dup_mem(e, tab->bit_rate_prob, sizeof(double) * cfg->num_bit_rates);
    e->t.link_order = dup_mem(e, tab->link_order, sizeof(int32_t) * cfg->num_links);
    e->link_util = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    e->link_comp = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    e->link_last = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    e->link_frag = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    size_t cells = (size_t)cfg->num_cores * cfg->num_links * cfg->num_slots;
    e->avail = (int8_t *)malloc(cells);
    e->alloc = (int32_t *)malloc(cells * sizeof(int32_t));
    build_thresholds(e->t.node_prob, N, e->src_thr);
    if (cfg->num_bit_rates > 0) build_thresholds(e->t.bit_rate_prob, cfg->num_bit_rates, e->br_thr);
    e->traffic = 1; e->seed = 0; e->env_id = 0;
    return e;
}

/* same, but borrows the caller's tables (they must outlive the env): used by the batched baseline */
void *oracle_create_shared(const ocfg_t *cfg, const otab_t *tab) {
    oenv_t *e = (oenv_t *)calloc(1, sizeof(oenv_t));
    e->c = *cfg;
    e->t = *tab;
    size_t cells = (size_t)cfg->num_cores * cfg->num_links * cfg->num_slots;
    e->avail = (int8_t *)malloc(cells);
    e->alloc = (int32_t *)malloc(cells * sizeof(int32_t));
    build_thresholds(e->t.node_prob, cfg->num_nodes, e->src_thr);
    if (cfg->num_bit_rates > 0) build_thresholds(e->t.bit_rate_prob, cfg->num_bit_rates, e->br_thr);
    e->traffic = 1;
    e->link_util = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    e->link_comp = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    e->link_last = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    e->link_frag = (double *)calloc((size_t)cfg->num_links, sizeof(double));
    return e;
}

void oracle_destroy(void *p) {
    oenv_t *e = (oenv_t *)p;
    for (int i = 0; i < e->nown; i++) free(e->own[i]);
    free(e->avail); free(e->alloc); free(e->heap); free(e->link_util); free(e->link_comp); free(e->link_last); free(e->link_frag); free(e);
}

void oracle_set_trace(void *p, const double *arr, const double *hold, const int32_t *src, const int32_t *dst,
                      const int32_t *br, int64_t len) {
    oenv_t *e = (oenv_t *)p;
    e->traffic = 0; e->tr_arr = arr; e->tr_hold = hold; e->tr_src = src; e->tr_dst = dst; e->tr_br = br; e->tr_len = len;
}

void oracle_set_philox(void *p, uint64_t seed, uint32_t env_id) {
    oenv_t *e = (oenv_t *)p;
    e->traffic = 1; e->seed = seed; e->env_id = env_id;
}

/* rmsa_env.py:284-359, rwa_env.py:164-208, rmcsa_env.py:386-483, optical_network_env.py:181-203 */
void oracle_reset(void *p, int full) {
    oenv_t *e = (oenv_t *)p;
    e->ep_br_req = 0; e->ep_br_prov = 0; e->ep_processed = 0; e->ep_accepted = 0;
    if (!full) {
        if (e->c.kind != KIND_RWA && e->new_service) {
            e->ep_processed += 1;
            e->ep_br_req += e->cur.bit_rate;
        }
        return;
    }
    e->nheap = 0; e->now = 0.0;
    e->processed = e->accepted = 0; e->br_req = e->br_prov = 0;
    e->req_index = 0;
    memset(e->hist_req, 0, sizeof(e->hist_req)); memset(e->hist_prov, 0, sizeof(e->hist_prov));   /* rmsa_env.py:348-349 */
    memset(e->act_rows, 0, sizeof(e->act_rows)); memset(e->act_cols, 0, sizeof(e->act_cols)); e->act_total = 0;   /* rwa_env.py:195-201 */
    e->sum_nh = 0;
    e->run_br = 0; e->g_thr = 0.0; e->g_comp = 0.0; e->g_last = 0.0;
    for (int l = 0; l < e->c.num_links; l++) { e->link_util[l] = 0.0; e->link_comp[l] = 0.0; e->link_last[l] = 0.0; e->link_frag[l] = 0.0; }
    size_t cells = (size_t)e->c.num_cores * e->c.num_links * e->c.num_slots;
    memset(e->avail, 1, cells);
    for (size_t i = 0; i < cells; i++) e->alloc[i] = -1;
    e->new_service = 0;
    next_service(e);
}

void oracle_get_counters(void *p, int64_t *c);

static void finish_step(oenv_t *e, ostep_t *o) {
    o->accepted = e->cur.accepted; o->path_row = e->cur.path_row; o->initial_slot = e->cur.initial_slot;
    o->number_slots = e->cur.accepted ? e->cur.number_slots : -1;
    o->core = e->cur.core; o->mod = e->cur.mod; o->service_id = e->cur.id;
    if (!e->cur.accepted) { o->path_row = -1; o->initial_slot = -1; o->core = -1; o->mod = -1; }
    if (e->c.kind == KIND_DEEPRMSA) o->reward = e->cur.accepted ? 1.0 : -1.0;   /* deeprmsa_env.py:123-124 */
    else o->reward = e->cur.accepted ? 1.0 : 0.0;                                /* optical_network_env.py:178-179 */
    oracle_get_counters(e, o->info_counters);       /* rmsa_env.py:234-248 / rwa_env.py:141-152 */
    e->new_service = 0;
    next_service(e);
    o->done = (e->ep_processed == e->c.episode_length) ? 1 : 0;
}

/* rmsa_env.py:163-282 */
static int step_rmsa(oenv_t *e, int path, int initial_slot, ostep_t *o) {
    int rc = 0;
    double previous_network_compactness = e->c.stats ? network_compactness(e) : 0.0;      /* rmsa_env.py:168-170 */
    e->cur.accepted = 0;
    if (path < e->c.k_paths && initial_slot < e->c.num_slots && path >= 0 && initial_slot >= 0) {
        int row = pair_row(e, e->cur.src, e->cur.dst, path);
        if (row < 0) rc = -1;
        else {
            int n = number_slots(e, e->cur.bit_rate, e->t.path_se[row]);
            if (is_path_free(e, row, 0, initial_slot, n)) {
                provision(e, row, 0, initial_slot, n);
                e->accepted++; e->ep_accepted++;
                e->br_prov += e->cur.bit_rate; e->ep_br_prov += e->cur.bit_rate;
                e->cur.accepted = 1;
                if (e->c.kind == KIND_RMSA && br_slot(e, e->cur.bit_rate) >= 0) e->hist_prov[br_slot(e, e->cur.bit_rate)]++;   /* rmsa_env.py:408-415 */
                heap_push(e, e->cur.arrival + e->cur.holding, &e->cur);
            }
        }
    }
    if (e->c.kind == KIND_RMSA && e->c.num_bit_rates > 0) {      /* rmsa_env.py:217-227, 268-273 */
        double lo = 0.0, hi = 0.0;
        for (int i = 0; i < e->c.num_bit_rates; i++) {
            double b = 0.0;
            if (e->hist_req[i] > 0) b = (double)(e->hist_req[i] - e->hist_prov[i]) / (double)e->hist_req[i];
            e->br_blocking[i] = b;
            if (i == 0 || b < lo) lo = b;
            if (i == 0 || b > hi) hi = b;
        }
        e->br_blocking[e->c.num_bit_rates] = hi - lo;
    }
    o->stats[0] = o->stats[1] = o->stats[2] = o->stats[3] = 0.0;
    if (e->c.stats) {   /* rmsa_env.py:229-264 */
        double cur = network_compactness(e);
        o->stats[0] = cur;
        o->stats[1] = previous_network_compactness - cur;
        o->stats[2] = mean_over_edges(e, e->link_comp);
        o->stats[3] = mean_over_edges(e, e->link_util);
    }
    finish_step(e, o);
    return rc;
}

int oracle_step(void *p, const int32_t *action, ostep_t *o) {
    oenv_t *e = (oenv_t *)p;
    int k = e->c.k_paths, S = e->c.num_slots;
    switch (e->c.kind) {
    case KIND_RMSA:
        return step_rmsa(e, action[0], action[1], o);
    case KIND_DEEPRMSA: {                          /* deeprmsa_env.py:48-58 */
        int a = action[0];
        if (a >= 0 && a < k * e->c.j) {
            int route = a / e->c.j, block = a % e->c.j;
            int row = pair_row(e, e->cur.src, e->cur.dst, route);
            if (row >= 0) {
                int bs[64], bl[64];
                int n = number_slots(e, e->cur.bit_rate, e->t.path_se[row]);
                int nb = available_blocks(e, row, n, bs, bl);
                if (block < nb) return step_rmsa(e, route, bs[block], o);
            }
            return step_rmsa(e, k, S, o);
        }
        return step_rmsa(e, k, S, o);
    }
    case KIND_RWA: {                               /* rwa_env.py:101-162 */
        int path = action[0], w = action[1], rc = 0;
        e->cur.accepted = 0;
        {   /* self.actions_output[path, wavelength] += 1 (rwa_env.py:103); shape (k + rej, W + rej) */
            const int rej = e->c.allow_rejection ? 1 : 0;
            if (path >= 0 && path < k + rej && w >= 0 && w < S + rej) { e->act_rows[path]++; e->act_cols[w]++; e->act_total++; }
            else rc = -1;                              /* numpy would raise IndexError */
        }
        if (path < k && w < S && path >= 0 && w >= 0) {
            int row = pair_row(e, e->cur.src, e->cur.dst, path);
            if (row < 0) rc = -1;
            else if (is_path_free(e, row, 0, w, 1)) {
                provision(e, row, 0, w, 1);
                e->cur.accepted = 1;
                e->accepted++; e->ep_accepted++;
                heap_push(e, e->cur.arrival + e->cur.holding, &e->cur);
            }
        }
        e->processed++; e->ep_processed++;
        finish_step(e, o);
        return rc;
    }
    case KIND_RMCSA: {                             /* rmcsa_env.py:209-339 */
        int path = action[0], mod = action[1], core = action[2], slot = action[3], rc = 0;
        e->cur.accepted = 0;
        if (path < k && mod < e->c.num_mods && core < e->c.num_cores && slot < S &&
            path >= 0 && mod >= 0 && core >= 0 && slot >= 0) {
            int row = pair_row(e, e->cur.src, e->cur.dst, path);
            if (row < 0) rc = -1;
            else {
                int n = number_slots(e, e->cur.bit_rate, e->t.mod_se[mod]);
                if (is_path_free(e, row, core, slot, n) &&
                    crosstalk_ok(e, mod, e->t.path_length[row], e->cur.bit_rate)) {
                    provision(e, row, core, slot, n);
                    e->accepted++; e->ep_accepted++;
                    e->br_prov += e->cur.bit_rate; e->ep_br_prov += e->cur.bit_rate;
                    e->cur.accepted = 1; e->cur.mod = mod;
                    heap_push(e, e->cur.arrival + e->cur.holding, &e->cur);
                }
            }
        }
        e->processed++; e->ep_processed++;
        e->br_req += e->cur.bit_rate; e->ep_br_req += e->cur.bit_rate;     /* the double count, rmcsa_env.py:294-295 */
        finish_step(e, o);
        return rc;
    }
    }
    return -2;
}

/* deeprmsa_env.py:60-121, float64, the reference's operation order */
void oracle_observation(void *p, double *obs) {
    oenv_t *e = (oenv_t *)p;
    int N = e->c.num_nodes, k = e->c.k_paths, J = e->c.j, S = e->c.num_slots;
    int w = 2 * J + 3, n_obs = 1 + 2 * N + w * k;
    for (int i = 0; i < n_obs; i++) obs[i] = 0.0;
    obs[0] = (double)e->cur.bit_rate / 100;
    int mn = e->cur.src < e->cur.dst ? e->cur.src : e->cur.dst;
    int mx = e->cur.src < e->cur.dst ? e->cur.dst : e->cur.src;
    obs[1 + mn] = 1.0; obs[1 + N + mx] = 1.0;
    double *sp = obs + 1 + 2 * N;
    for (int i = 0; i < w * k; i++) sp[i] = -1.0;
    int key = e->cur.src * N + e->cur.dst;
    for (int idp = 0; idp < e->t.pair_count[key] && idp < k; idp++) {
        int row = e->t.pair_first[key] + idp;
        int8_t av[1024]; int st[1024], va[1024], le[1024], bs[64], bl[64];
        available_slots(e, row, 0, av);
        int n = number_slots(e, e->cur.bit_rate, e->t.path_se[row]);
        int nb = available_blocks(e, row, n, bs, bl);
        for (int b = 0; b < nb; b++) {
            sp[idp * w + 2 * b + 0] = 2 * ((double)bs[b] - 0.5 * S) / S;
            sp[idp * w + 2 * b + 1] = ((double)bl[b] - 8) / 8;
        }
        sp[idp * w + 2 * J] = ((double)n - 5.5) / 3.5;
        int nr = rle(av, S, st, va, le);
        int tot = 0; for (int s = 0; s < S; s++) tot += av[s];
        sp[idp * w + 2 * J + 1] = 2 * ((double)tot - 0.5 * S) / S;
        int cnt = 0; long sum = 0;
        for (int r = 0; r < nr; r++) if (va[r] == 1) { cnt++; sum += le[r]; }
        if (cnt > 0) sp[idp * w + 2 * J + 2] = ((double)sum / (double)cnt - 4) / 4;
    }
}

/* Integer pre-image of the observation: per path, for b<j: start_b,len_b (or -1),
 * then n_slots, total_free, n_free_runs.  Layout [k][2j+3] int32. */
void oracle_observation_int(void *p, int32_t *out) {
    oenv_t *e = (oenv_t *)p;
    int N = e->c.num_nodes, k = e->c.k_paths, J = e->c.j, S = e->c.num_slots, w = 2 * J + 3;
    for (int i = 0; i < w * k; i++) out[i] = -1;
    int key = e->cur.src * N + e->cur.dst;
    for (int idp = 0; idp < e->t.pair_count[key] && idp < k; idp++) {
        int row = e->t.pair_first[key] + idp;
        int8_t av[1024]; int st[1024], va[1024], le[1024], bs[64], bl[64];
        available_slots(e, row, 0, av);
        int n = number_slots(e, e->cur.bit_rate, e->t.path_se[row]);
        int nb = available_blocks(e, row, n, bs, bl);
        for (int b = 0; b < nb; b++) { out[idp * w + 2 * b] = bs[b]; out[idp * w + 2 * b + 1] = bl[b]; }
        int nr = rle(av, S, st, va, le), tot = 0, cnt = 0;
        for (int s = 0; s < S; s++) tot += av[s];
        for (int r = 0; r < nr; r++) if (va[r] == 1) cnt++;
        out[idp * w + 2 * J] = n; out[idp * w + 2 * J + 1] = tot; out[idp * w + 2 * J + 2] = cnt;
    }
}

/* Heuristic action sources.  kind: 0 SP-FF, 1 SAP-FF, 2 LLP-FF, 3 SAP-LF (RWA) */
void oracle_heuristic(void *p, int which, int32_t *action) {
    oenv_t *e = (oenv_t *)p;
    int N = e->c.num_nodes, k = e->c.k_paths, S = e->c.num_slots;
    int key = e->cur.src * N + e->cur.dst, first = e->t.pair_first[key], cnt = e->t.pair_count[key];
    if (cnt > k) cnt = k;
    switch (e->c.kind) {
    case KIND_RMSA: {                                  /* rmsa_env.py:747-803: range(0, S - n) */
        action[0] = k; action[1] = S;
        if (which == 0 || which == 1) {
            int np = which == 0 ? 1 : cnt;
            for (int idp = 0; idp < np; idp++) {
                int n = number_slots(e, e->cur.bit_rate, e->t.path_se[first + idp]);
                for (int s = 0; s < S - n; s++)
                    if (is_path_free(e, first + idp, 0, s, n)) { action[0] = idp; action[1] = s; return; }
            }
        } else {
            int max_free = 0;
            for (int idp = 0; idp < cnt; idp++) {
                int n = number_slots(e, e->cur.bit_rate, e->t.path_se[first + idp]);
                for (int s = 0; s < S - n; s++)
                    if (is_path_free(e, first + idp, 0, s, n)) {
                        int8_t av[1024]; int fr = 0;
                        available_slots(e, first + idp, 0, av);
                        for (int q = 0; q < S; q++) fr += av[q];
                        if (fr > max_free) { action[0] = idp; action[1] = s; max_free = fr; }
                        break;
                    }
            }
        }
        return;
    }
    case KIND_DEEPRMSA: {                              /* deeprmsa_env.py:135-155 */
        int J = e->c.j, bs[64], bl[64];
        if (which == 0) {
            if (!e->c.allow_rejection) { action[0] = 0; return; }
            int n = number_slots(e, e->cur.bit_rate, e->t.path_se[first]);
            action[0] = available_blocks(e, first, n, bs, bl) > 0 ? 0 : k * J;
            return;
        }
        for (int idp = 0; idp < cnt; idp++) {
            int n = number_slots(e, e->cur.bit_rate, e->t.path_se[first + idp]);
            if (available_blocks(e, first + idp, n, bs, bl) > 0) { action[0] = idp * J; return; }
        }
        action[0] = k * J;
        return;
    }
    case KIND_RWA: {                                   /* rwa_env.py:425-502 */
        action[0] = k; action[1] = S;
        if (which == 0) {
            for (int w = 0; w < S; w++) if (is_path_free(e, first, 0, w, 1)) { action[0] = 0; action[1] = w; return; }
        } else if (which == 1 || which == 3) {
            double best_hops = 1.7976931348623157e308;
            for (int idp = 0; idp < cnt; idp++) {
                if ((double)e->t.path_hops[first + idp] < best_hops) {
                    if (which == 1) {
                        for (int w = 0; w < S; w++)
                            if (is_path_free(e, first + idp, 0, w, 1)) { best_hops = e->t.path_hops[first + idp]; action[0] = idp; action[1] = w; break; }
                    } else {
                        for (int w = S - 1; w > 0; w--)
                            if (is_path_free(e, first + idp, 0, w, 1)) { best_hops = e->t.path_hops[first + idp]; action[0] = idp; action[1] = w; break; }
                    }
                }
            }
        } else {
            double best_load = -1.7976931348623157e308;
            for (int idp = 0; idp < cnt; idp++) {
                int cap = 0;
                for (int w = 0; w < S; w++) cap += is_path_free(e, first + idp, 0, w, 1);
                if ((double)cap > best_load) {
                    for (int w = 0; w < S; w++)
                        if (is_path_free(e, first + idp, 0, w, 1)) { best_load = cap; action[0] = idp; action[1] = w; break; }
                }
            }
        }
        return;
    }
    case KIND_RMCSA: {                                 /* rmcsa_env.py:882-911 (reject widened to a 4-tuple) */
        action[0] = k; action[1] = e->c.num_mods; action[2] = e->c.num_cores; action[3] = S;
        for (int idp = 0; idp < cnt; idp++) {
            int mod = e->t.path_mod[first + idp];
            int n = number_slots(e, e->cur.bit_rate, e->t.mod_se[mod]);
            for (int core = 0; core < e->c.num_cores; core++)
                for (int s = 0; s < S - n; s++)
                    if (is_path_free(e, first + idp, core, s, n)) { action[0] = idp; action[1] = mod; action[2] = core; action[3] = s; return; }
        }
        return;
    }
    }
}

/* Uniform random policy of the benchmark (DESIGN.md "Traffic": Philox stream 2) */
void oracle_random_action(void *p, int32_t *action) {
    oenv_t *e = (oenv_t *)p;
    uint64_t idx = (uint64_t)e->req_index;
    uint32_t c[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), e->env_id, 2u };
    philox4x32_10(c, (uint32_t)e->seed, (uint32_t)(e->seed >> 32));
    int rej = e->c.allow_rejection ? 1 : 0;
    switch (e->c.kind) {
    case KIND_DEEPRMSA: action[0] = (int32_t)(((uint64_t)c[0] * (uint32_t)(e->c.k_paths * e->c.j + rej)) >> 32); break;
    case KIND_RMSA: case KIND_RWA:
        action[0] = (int32_t)(((uint64_t)c[0] * (uint32_t)(e->c.k_paths + rej)) >> 32);
        action[1] = (int32_t)(((uint64_t)c[1] * (uint32_t)(e->c.num_slots + rej)) >> 32); break;
    case KIND_RMCSA:
        action[0] = (int32_t)(((uint64_t)c[0] * (uint32_t)(e->c.k_paths + rej)) >> 32);
        action[1] = (int32_t)(((uint64_t)c[1] * (uint32_t)e->c.num_mods) >> 32);
        action[2] = (int32_t)(((uint64_t)c[2] * (uint32_t)(e->c.num_cores + rej)) >> 32);
        action[3] = (int32_t)(((uint64_t)c[3] * (uint32_t)(e->c.num_slots + rej)) >> 32); break;
    }
}

void oracle_get_request(void *p, double *arrival, double *holding, int32_t *ints /* src,dst,bit_rate,id */) {
    oenv_t *e = (oenv_t *)p;
    *arrival = e->cur.arrival; *holding = e->cur.holding;
    ints[0] = e->cur.src; ints[1] = e->cur.dst; ints[2] = e->cur.bit_rate; ints[3] = e->cur.id;
}

/* info["bit_rate_blocking_<rate>"] (order of bit_rates) and info["fairness"] of the last step */
void oracle_get_bit_rate_blocking(void *p, double *out /* [num_bit_rates + 1] */) {
    oenv_t *e = (oenv_t *)p;
    for (int i = 0; i <= e->c.num_bit_rates; i++) out[i] = e->br_blocking[i];
}

/* info["path_action_probability"] ++ info["wavelength_action_probability"] (rwa_env.py:148-151):
   np.sum(actions_output, axis=1) / np.sum(actions_output), then axis=0 */
void oracle_get_action_probability(void *p, double *out /* [(k + rej) + (W + rej)] */) {
    oenv_t *e = (oenv_t *)p;
    const int rej = e->c.allow_rejection ? 1 : 0, R = e->c.k_paths + rej, Cn = e->c.num_slots + rej;
    for (int i = 0; i < R; i++) out[i] = (double)e->act_rows[i] / (double)e->act_total;
    for (int i = 0; i < Cn; i++) out[R + i] = (double)e->act_cols[i] / (double)e->act_total;
}

void oracle_get_counters(void *p, int64_t *c /* [8] */) {
    oenv_t *e = (oenv_t *)p;
    c[0] = e->processed; c[1] = e->accepted; c[2] = e->ep_processed; c[3] = e->ep_accepted;
    c[4] = e->br_req; c[5] = e->br_prov; c[6] = e->ep_br_req; c[7] = e->ep_br_prov;
}

void oracle_get_state(void *p, int8_t *avail, int32_t *alloc, double *now, int32_t *nheap) {
    oenv_t *e = (oenv_t *)p;
    size_t cells = (size_t)e->c.num_cores * e->c.num_links * e->c.num_slots;
    if (avail) memcpy(avail, e->avail, cells);
    if (alloc) memcpy(alloc, e->alloc, cells * sizeof(int32_t));
    if (now) *now = e->now;
    if (nheap) *nheap = e->nheap;
}

/* graph attributes of row f1: link [E][3] = utilization, external_fragmentation, compactness (link index order);
   graph [2] = throughput, compactness */
void oracle_get_link_stats(void *p, double *link, double *graph) {
    oenv_t *e = (oenv_t *)p;
    for (int l = 0; l < e->c.num_links; l++) { link[3 * l] = e->link_util[l]; link[3 * l + 1] = e->link_frag[l]; link[3 * l + 2] = e->link_comp[l]; }
    graph[0] = e->g_thr; graph[1] = e->g_comp;
}

int oracle_error(void *p) { return ((oenv_t *)p)->error; }

/* ------------------------------------------------------------------ batched rollout (parity at scale + CPU baseline)
 * policy: 0 = actions given ([T][A] int32 per env), 1 = uniform random (Philox stream 2),
 *         10+h = heuristic h.  Auto-reset after done (VecEnv semantics = evaluate_heuristic's reset()).
 * Outputs (any may be NULL): per step accepted/path_row/initial_slot/number_slots ([T][4] int32), reward [T] f64,
 * done [T] u8, obs [T][obs_dim] f64 (DeepRMSA; observation AFTER the step). */
long oracle_rollout(void *p, long T, int policy, const int32_t *actions, int adim,
                    int32_t *decisions, double *rewards, uint8_t *dones, double *obs, int obs_dim, int32_t *actions_out) {
    oenv_t *e = (oenv_t *)p;
    long acc = 0;
    for (long t = 0; t < T; t++) {
        int32_t a[4] = {0, 0, 0, 0};
        if (policy == 0) for (int i = 0; i < adim; i++) a[i] = actions[t * adim + i];
        else if (policy == 1) oracle_random_action(e, a);
        else oracle_heuristic(e, policy - 10, a);
        if (actions_out) for (int i = 0; i < adim; i++) actions_out[t * adim + i] = a[i];
        ostep_t o;
        oracle_step(e, a, &o);
        acc += o.accepted;
        if (decisions) { decisions[t * 4] = o.accepted; decisions[t * 4 + 1] = o.path_row; decisions[t * 4 + 2] = o.initial_slot; decisions[t * 4 + 3] = o.number_slots; }
        if (rewards) rewards[t] = o.reward;
        if (dones) dones[t] = (uint8_t)o.done;
        if (obs) oracle_observation(e, obs + (size_t)t * obs_dim);
        if (o.done) oracle_reset(e, 0);
    }
    return acc;
}

/* Multi-threaded throughput run for bench.py's cpu_baseline / --impl reference:
 * n_envs independent envs (env ids id0..), T steps each, all host threads. Returns total accepted. */
typedef struct { const ocfg_t *cfg; const otab_t *tab; uint64_t seed; uint32_t id0; int n_envs, stride, tid; long T, warm; int policy, with_obs; long acc; } mt_arg_t;

static void *mt_worker(void *vp) {
    mt_arg_t *a = (mt_arg_t *)vp;
    double obs[4096];
    for (int i = a->tid; i < a->n_envs; i += a->stride) {
        void *e = oracle_create(a->cfg, a->tab);
        oracle_set_philox(e, a->seed, a->id0 + (uint32_t)i);
        oracle_reset(e, 1);
        int adim = a->cfg->kind == KIND_DEEPRMSA ? 1 : (a->cfg->kind == KIND_RMCSA ? 4 : 2);
        for (long t = 0; t < a->T; t++) {
            int32_t act[4];
            if (a->policy == 1) oracle_random_action(e, act); else oracle_heuristic(e, a->policy - 10, act);
            (void)adim;
            ostep_t o;
            oracle_step(e, act, &o);
            a->acc += o.accepted;
            if (a->with_obs && a->cfg->kind == KIND_DEEPRMSA) oracle_observation(e, obs);
            if (o.done) oracle_reset(e, 0);
        }
        oracle_destroy(e);
    }
    return NULL;
}

long oracle_rollout_mt(const ocfg_t *cfg, const otab_t *tab, uint64_t seed, uint32_t id0, int n_envs, long T,
                       int policy, int with_obs, int n_threads) {
    pthread_t th[256]; mt_arg_t args[256];
    if (n_threads > 256) n_threads = 256;
    if (n_threads < 1) n_threads = 1;
    for (int i = 0; i < n_threads; i++) {
        mt_arg_t a = { cfg, tab, seed, id0, n_envs, n_threads, i, T, 0, policy, with_obs, 0 };
        args[i] = a;
        pthread_create(&th[i], NULL, mt_worker, &args[i]);
    }
    long acc = 0;
    for (int i = 0; i < n_threads; i++) { pthread_join(th[i], NULL); acc += args[i].acc; }
    return acc;
}

/* ------------------------------------------------------------------ persistent batch ("SubprocVecEnv"-shaped CPU baseline)
 * n independent envs kept alive across calls; oracle_vec_run advances every env by `steps` steps using
 * n_threads host threads (static partition, no per-step barrier: the envs never interact). */
typedef struct { int n; void **envs; ocfg_t cfg; otab_t tab; oenv_t *tabs_owner; } ovec_t;
typedef struct { ovec_t *v; int tid, stride; long steps; int policy, with_obs; long acc; } vec_arg_t;

void *oracle_vec_create(const ocfg_t *cfg, const otab_t *tab, uint64_t seed, uint32_t id0, int n) {
    ovec_t *v = (ovec_t *)calloc(1, sizeof(ovec_t));
    v->n = n; v->cfg = *cfg;
    v->tabs_owner = (oenv_t *)oracle_create(cfg, tab);     /* owns one copy of the tables */
    v->tab = v->tabs_owner->t;
    v->envs = (void **)calloc((size_t)n, sizeof(void *));
    for (int i = 0; i < n; i++) {
        v->envs[i] = oracle_create_shared(&v->cfg, &v->tab);
        oracle_set_philox(v->envs[i], seed, id0 + (uint32_t)i);
        oracle_reset(v->envs[i], 1);
    }
    return v;
}

void oracle_vec_destroy(void *p) {
    ovec_t *v = (ovec_t *)p;
    for (int i = 0; i < v->n; i++) oracle_destroy(v->envs[i]);
    oracle_destroy(v->tabs_owner);
    free(v->envs); free(v);
}

static void *vec_worker(void *vp) {
    vec_arg_t *a = (vec_arg_t *)vp;
    double obs[4096];
    for (int i = a->tid; i < a->v->n; i += a->stride) {
        void *e = a->v->envs[i];
        for (long t = 0; t < a->steps; t++) {
            int32_t act[4];
            if (a->policy == 1) oracle_random_action(e, act); else oracle_heuristic(e, a->policy - 10, act);
            ostep_t o;
            oracle_step(e, act, &o);
            a->acc += o.accepted;
            if (a->with_obs && a->v->cfg.kind == KIND_DEEPRMSA) oracle_observation(e, obs);
            if (o.done) oracle_reset(e, 0);
        }
    }
    return NULL;
}

long oracle_vec_run(void *p, long steps, int policy, int with_obs, int n_threads) {
    ovec_t *v = (ovec_t *)p;
    pthread_t th[512]; vec_arg_t args[512];
    if (n_threads > 512) n_threads = 512;
    if (n_threads < 1) n_threads = 1;
    for (int i = 0; i < n_threads; i++) {
        vec_arg_t a = { v, i, n_threads, steps, policy, with_obs, 0 };
        args[i] = a;
        pthread_create(&th[i], NULL, vec_worker, &args[i]);
    }
    long acc = 0;
    for (int i = 0; i < n_threads; i++) { pthread_join(th[i], NULL); acc += args[i].acc; }
    return acc;
}

void *oracle_vec_env(void *p, int i) { return ((ovec_t *)p)->envs[i]; }
