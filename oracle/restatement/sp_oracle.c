/* ORACLE / TEST INFRASTRUCTURE ONLY -- see sp_oracle.h for scope, pinning status and citation rules.
 * All `ref:` citations are relative to /root/reference/.
 */
#include "sp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ref: src/array_and_matrix_operations.cpp:476-486 */
void orc_syndrome(const orc_graph *g, const int32_t *bits, int32_t *syndrome_out)
{
    for (int32_t j = 0; j < g->m; ++j) {
        int32_t s = 0;
        for (int32_t p = g->row_ptr[j]; p < g->row_ptr[j + 1]; ++p)
            s ^= bits[g->col_idx[p]];
        syndrome_out[j] = s;
    }
}

/* ref: src/array_and_matrix_operations.cpp:96-106 */
int orc_arrays_equal(const int32_t *a, const int32_t *b, size_t n)
{
    for (size_t i = 0; i < n; ++i)
        if (a[i] != b[i])
            return 0;
    return 1;
}

/* Optional copy-out of the per-iteration intermediates the reference prints under CFG.TRACE_SUM_PRODUCT
 * (ref: src/qkd_ldpc_algorithm.cpp:251-255 E, :268-276 L and z, :279-283 s, :317-321 M); fp64 only. */
typedef struct orc_trace_sink {
    uint64_t cap;
    double *e, *l, *m;
    int32_t *z, *s;
} orc_trace_sink;
static __thread orc_trace_sink *g_trace = NULL;

#define REAL double
#define SP_TANH tanh
#define SP_ATANH atanh
#define SP_NAME sp_decode_f64
#define SP_TRACE g_trace
#include "sp_decode_body.inc"
#undef SP_TRACE
#undef REAL
#undef SP_TANH
#undef SP_ATANH
#undef SP_NAME

#define REAL float
#define SP_TANH tanhf
#define SP_ATANH atanhf
#define SP_NAME sp_decode_f32
#define SP_TRACE ((orc_trace_sink *)NULL)
#include "sp_decode_body.inc"
#undef SP_TRACE
#undef REAL
#undef SP_TANH
#undef SP_ATANH
#undef SP_NAME

orc_sp_result orc_sum_product_f64_trace(const orc_graph *g, const double *llr, const int32_t *syndrome, uint64_t max_it,
                                        int enable_threshold, double thr, uint64_t capacity, double *e_out, double *l_out,
                                        int32_t *z_out, int32_t *s_out, double *m_out, int32_t *bits_out)
{
    orc_trace_sink sink = {capacity, e_out, l_out, m_out, z_out, s_out};
    g_trace = &sink;
    const orc_sp_result r = sp_decode_f64(g, llr, syndrome, max_it, enable_threshold, thr, ORC_F32_DIVIDE, bits_out);
    g_trace = NULL;
    return r;
}

orc_sp_result orc_sum_product_f64(const orc_graph *g, const double *llr, const int32_t *syndrome, uint64_t max_it,
                                  int enable_threshold, double thr, int32_t *bits_out)
{
    return sp_decode_f64(g, llr, syndrome, max_it, enable_threshold, thr, ORC_F32_DIVIDE, bits_out);
}

orc_sp_result orc_sum_product_f32(const orc_graph *g, const float *llr, const int32_t *syndrome, uint64_t max_it,
                                  int enable_threshold, float thr, int form, int32_t *bits_out)
{
    return sp_decode_f32(g, llr, syndrome, max_it, enable_threshold, thr, form, bits_out);
}

/* ref: src/qkd_ldpc_algorithm.cpp:398-447 */
orc_ldpc_result orc_qkd_ldpc(const orc_graph *g, const int32_t *alice, const int32_t *bob, double qber, uint64_t max_it,
                             int enable_threshold, double thr, int precision, int f32_form, int32_t *syndrome_out,
                             int32_t *decoded_out)
{
    const int32_t n = g->n, m = g->m;
    orc_ldpc_result r;
    const double log_p = log((1. - qber) / qber);                       /* ref :400 */
    int32_t *syn = syndrome_out ? syndrome_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    int32_t *dec = decoded_out ? decoded_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    orc_syndrome(g, alice, syn);                                        /* ref :413-414 */
    if (precision == 32) {
        float *llr = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        const float lp = (float)log_p;
        for (int32_t i = 0; i < n; ++i)
            llr[i] = bob[i] ? -lp : lp;
        r.sp_res = sp_decode_f32(g, llr, syn, max_it, enable_threshold, (float)thr, f32_form, dec);
        free(llr);
    } else {
        double *llr = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        for (int32_t i = 0; i < n; ++i)
            llr[i] = bob[i] ? -log_p : log_p;                           /* ref :401-405 */
        r.sp_res = sp_decode_f64(g, llr, syn, max_it, enable_threshold, thr, ORC_F32_DIVIDE, dec); /* ref :424-425 */
        free(llr);
    }
    r.keys_match = orc_arrays_equal(alice, dec, (size_t)n);             /* ref :433 */
    if (!syndrome_out) free(syn);
    if (!decoded_out) free(dec);
    return r;
}

/* ---- PRNG: Reputeless/Xoshiro-cpp v1.1 (xoshiro256++ / SplitMix64, Blackman & Vigna) ---- */
static inline uint64_t rotl64(uint64_t x, int s) { return (x << s) | (x >> (64 - s)); }

void orc_prng_seed(orc_prng *p, uint64_t seed)
{
    uint64_t st = seed;
    for (int i = 0; i < 4; ++i) {
        uint64_t z = (st += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        p->s[i] = z ^ (z >> 31);
    }
}

uint64_t orc_prng_next(orc_prng *p)
{
    uint64_t *s = p->s;
    const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* ref: src/simulation.cpp:222-228. uniform_int_distribution<size_t>(0, SIZE_MAX) over a full-range 64-bit
 * URBG returns the raw draw (libstdc++ 13.3 bits/uniform_int_dist.h:319-320, the `urngrange == urange` arm). */
void orc_trial_seeds(uint64_t simulation_seed, uint64_t count, uint64_t *seeds_out)
{
    orc_prng p;
    orc_prng_seed(&p, simulation_seed);
    for (uint64_t i = 0; i < count; ++i)
        seeds_out[i] = orc_prng_next(&p);
}

/* libstdc++ 13.3 uniform_int_distribution on a 64-bit full-range URBG, range < 2^64: Lemire's
 * multiply-shift with rejection (bits/uniform_int_dist.h:257-280 `_S_nd`, selected at :292-303).
 * Returns a value uniform on [0, range). */
static uint64_t lemire_u64(orc_prng *p, uint64_t range)
{
    unsigned __int128 product = (unsigned __int128)orc_prng_next(p) * range;
    uint64_t low = (uint64_t)product;
    if (low < range) {
        const uint64_t threshold = (0 - range) % range;
        while (low < threshold) {
            product = (unsigned __int128)orc_prng_next(p) * range;
            low = (uint64_t)product;
        }
    }
    return (uint64_t)(product >> 64);
}

/* ref: src/array_and_matrix_operations.cpp:424-460 */
double orc_generate(uint64_t seed, uint64_t n, double qber, int32_t *alice_out, int32_t *bob_out)
{
    orc_prng p;
    orc_prng_seed(&p, seed);
    /* :424-431 -- uniform_int_distribution<int>(0,1): one draw per bit, range 2 => the draw's top bit */
    for (uint64_t i = 0; i < n; ++i)
        alice_out[i] = (int32_t)lemire_u64(&p, 2);

    /* :436 */
    const uint64_t num_errors = (uint64_t)((double)n * qber);
    memcpy(bob_out, alice_out, sizeof(int32_t) * n);
    if (num_errors != 0) {
        uint64_t *pos = (uint64_t *)malloc(sizeof(uint64_t) * n);
        for (uint64_t i = 0; i < n; ++i)
            pos[i] = i;
        /* :448 std::shuffle -- libstdc++ 13.3 bits/stl_algo.h:3742-3806. For a 64-bit URBG and
         * n*n <= 2^64-1 two swap positions are taken from one draw (`__gen_two_uniform_ints`, :3719-3731). */
        if (n > 1) {
            if (UINT64_MAX / n >= n) {
                uint64_t i = 1;
                if ((n % 2) == 0) {
                    const uint64_t j = lemire_u64(&p, 2);
                    uint64_t t = pos[i]; pos[i] = pos[j]; pos[j] = t;
                    ++i;
                }
                while (i != n) {
                    const uint64_t swap_range = i + 1;
                    const uint64_t b1 = swap_range + 1;
                    const uint64_t x = lemire_u64(&p, swap_range * b1);
                    const uint64_t j0 = x / b1, j1 = x % b1;
                    uint64_t t = pos[i]; pos[i] = pos[j0]; pos[j0] = t;
                    ++i;
                    t = pos[i]; pos[i] = pos[j1]; pos[j1] = t;
                    ++i;
                }
            } else {
                for (uint64_t i = 1; i < n; ++i) {
                    const uint64_t j = (i + 1 == 0) ? orc_prng_next(&p) : lemire_u64(&p, i + 1);
                    uint64_t t = pos[i]; pos[i] = pos[j]; pos[j] = t;
                }
            }
        }
        for (uint64_t k = 0; k < num_errors; ++k)      /* :451-454 */
            bob_out[pos[k]] ^= 1;
        free(pos);
    }
    return (double)num_errors / (double)n;               /* :459 */
}

/* ref: src/simulation.cpp:161-189 */
int orc_run_trial(const orc_graph *g, double qber, uint64_t seed, uint64_t max_it, int enable_threshold, double thr,
                  int precision, int f32_form, orc_ldpc_result *res_out, double *exact_qber_out)
{
    const size_t n = (size_t)g->n;
    int32_t *alice = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    int32_t *bob = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    const double exact = orc_generate(seed, n, qber, alice, bob);
    int rc = 0;
    if (exact == 0.) {
        rc = -1; /* the reference throws "Key size ... is too small for QBER." (:170-175) */
    } else {
        *res_out = orc_qkd_ldpc(g, alice, bob, exact, max_it, enable_threshold, thr, precision, f32_form, NULL, NULL);
    }
    if (exact_qber_out) *exact_qber_out = exact;
    free(alice);
    free(bob);
    return rc;
}

typedef struct trials_job {
    const orc_graph *g;
    double qber;
    const uint64_t *seeds;
    uint64_t count;
    uint64_t max_it;
    int enable_threshold;
    double thr;
    int precision;
    int f32_form;
    uint64_t *out3;
    int32_t *decoded_out;
    uint64_t next;   /* shared work counter (atomic) */
    int status;
} trials_job;

static void *trials_worker(void *arg)
{
    trials_job *jb = (trials_job *)arg;
    const size_t n = (size_t)jb->g->n;
    int32_t *alice = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    int32_t *bob = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    for (;;) {
        const uint64_t k = __atomic_fetch_add(&jb->next, 1, __ATOMIC_RELAXED);
        if (k >= jb->count)
            break;
        const double exact = orc_generate(jb->seeds[k], n, jb->qber, alice, bob);
        if (exact == 0.) {
            __atomic_store_n(&jb->status, -1, __ATOMIC_RELAXED);
            continue;
        }
        orc_ldpc_result r = orc_qkd_ldpc(jb->g, alice, bob, exact, jb->max_it, jb->enable_threshold, jb->thr,
                                         jb->precision, jb->f32_form, NULL,
                                         jb->decoded_out ? jb->decoded_out + (size_t)k * n : NULL);
        jb->out3[3 * k + 0] = r.sp_res.iterations_num;
        jb->out3[3 * k + 1] = (uint64_t)r.sp_res.syndromes_match;
        jb->out3[3 * k + 2] = (uint64_t)r.keys_match;
    }
    free(alice);
    free(bob);
    return NULL;
}

/* The per-trial work is orc_run_trial's (generate + reconcile); trials are spread over `threads` host threads
 * the way the reference spreads them over its pool (ref: src/simulation.cpp:244-250). */
int orc_run_trials(const orc_graph *g, double qber, const uint64_t *seeds, uint64_t count, int threads, uint64_t max_it,
                   int enable_threshold, double thr, int precision, int f32_form, uint64_t *out3, int32_t *decoded_out)
{
    trials_job jb = {g, qber, seeds, count, max_it, enable_threshold, thr, precision, f32_form, out3, decoded_out, 0, 0};
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 1; t < threads; ++t)
        if (pthread_create(&tid[started], NULL, trials_worker, &jb) == 0)
            ++started;
    trials_worker(&jb);
    for (int t = 0; t < started; ++t)
        pthread_join(tid[t], NULL);
    free(tid);
    return jb.status;
}

/* ref: src/simulation.cpp:48-70 -- first preset (sorted by rate) with code_rate <= preset.code_rate;
 * end-exclusive grid: steps = round((end-begin)/step), value = begin + j*step. */
int orc_qber_range(double code_rate, const double *params4, size_t n_params, double *out, size_t cap)
{
    for (size_t i = 0; i < n_params; ++i) {
        const double rate = params4[4 * i], begin = params4[4 * i + 1], end = params4[4 * i + 2], step = params4[4 * i + 3];
        if (code_rate <= rate) {
            const size_t steps = (size_t)round((end - begin) / step);
            for (size_t j = 0; j < steps && j < cap; ++j)
                out[j] = begin + (double)j * step;
            return steps ? (int)steps : -1;
        }
    }
    return -1;
}

/* ref: src/simulation.cpp:252-312 */
void orc_point_stats(const uint64_t *out3, uint64_t trials, uint64_t max_it, double *stats_out)
{
    uint64_t ok_sp = 0, ok_ldpc = 0, it_max = 0, it_min = max_it;
    double mean = 0, sd = 0;
    for (uint64_t k = 0; k < trials; ++k) {
        if (out3[3 * k + 1]) {
            const uint64_t it = out3[3 * k];
            ++ok_sp;
            if (it_max < it) it_max = it;
            if (it_min > it) it_min = it;
            if (out3[3 * k + 2]) ++ok_ldpc;
            mean += (double)it;
        }
    }
    if (ok_sp > 0) {
        mean /= (double)ok_sp;
        for (uint64_t k = 0; k < trials; ++k)
            if (out3[3 * k + 1])
                sd += pow((double)out3[3 * k] - mean, 2);
        sd /= (double)ok_sp;
        sd = sqrt(sd);
    }
    stats_out[0] = mean;
    stats_out[1] = sd;
    stats_out[2] = (it_min == max_it) ? 0. : (double)it_min;
    stats_out[3] = (double)it_max;
    stats_out[4] = (double)ok_sp / (double)trials;
    stats_out[5] = (double)ok_ldpc / (double)trials;
}
