/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the hot path of ColdCloudd/QKD_LDPC, used exclusively by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker. The product path (qkd_ldpc_b200/csrc, libqkdldpc_b200.so) never includes, links or
 * calls anything in this directory.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle_*.py) against
 *   - the unmodified reference compiled in place (oracle/_ref/libqkdref.so, oracle/Makefile),
 *   - the golden vectors the survey obtained from the reference,
 *   - the textbook known-answer example the reference ships (example/qkd_ldpc_example.cpp:34-39).
 *
 * All `ref:` citations are relative to /root/reference/.
 */
#ifndef SP_ORACLE_H
#define SP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened Tanner graph: both adjacency halves of the reference's H_matrix
 * (ref: src/array_and_matrix_operations.hpp:16-27), list order preserved. */
typedef struct orc_graph {
    int32_t n;              /* bit nodes   (num_bit_nodes)   */
    int32_t m;              /* check nodes (num_check_nodes) */
    int32_t e;              /* edges */
    const int32_t *row_ptr; /* [m+1] offsets into col_idx: check_nodes[j][k] = col_idx[row_ptr[j]+k] */
    const int32_t *col_idx; /* [e]   bit ids per check, stored order */
    const int32_t *col_ptr; /* [n+1] offsets into row_idx: bit_nodes[i][k] = row_idx[col_ptr[i]+k]   */
    const int32_t *row_idx; /* [e]   check ids per bit, stored order */
} orc_graph;

typedef struct orc_sp_result {
    uint64_t iterations_num; /* ref: src/qkd_ldpc_algorithm.hpp:14-18 */
    int32_t syndromes_match;
} orc_sp_result;

typedef struct orc_ldpc_result {
    orc_sp_result sp_res;    /* ref: src/qkd_ldpc_algorithm.hpp:20-24 */
    int32_t keys_match;
} orc_ldpc_result;

/* fp32 check-node product form (the reference has no fp32 path; SURVEY.md 8a "fp32 note") */
enum { ORC_F32_DIVIDE = 0, ORC_F32_LEAVE_ONE_OUT = 1 };

/* ref: src/array_and_matrix_operations.cpp:476-486 (and :463-473 for consistent regular matrices) */
void orc_syndrome(const orc_graph *g, const int32_t *bits, int32_t *syndrome_out);

/* ref: src/array_and_matrix_operations.cpp:96-106 */
int orc_arrays_equal(const int32_t *a, const int32_t *b, size_t n);

/* ref: src/qkd_ldpc_algorithm.cpp:175-345 (irregular) == :3-173 (regular) for consistent weights.
 * enable_threshold mirrors CFG.ENABLE_SUM_PRODUCT_MSG_LLR_THRESHOLD (:246,313). */
orc_sp_result orc_sum_product_f64(const orc_graph *g, const double *llr, const int32_t *syndrome,
                                  uint64_t max_it, int enable_threshold, double thr, int32_t *bits_out);

/* The same decode with the intermediates the reference prints under CFG.TRACE_SUM_PRODUCT copied out for the first
 * `capacity` iterations (ref: src/qkd_ldpc_algorithm.cpp:251-255 E[capacity][e] bit-major, :268-276 L[..][n] and z[..][n],
 * :279-283 s[..][m], :317-321 M[..][e] check-major; M is not produced for the converging iteration). Any sink may be NULL. */
orc_sp_result orc_sum_product_f64_trace(const orc_graph *g, const double *llr, const int32_t *syndrome, uint64_t max_it,
                                        int enable_threshold, double thr, uint64_t capacity, double *e_out, double *l_out,
                                        int32_t *z_out, int32_t *s_out, double *m_out, int32_t *bits_out);

/* Same schedule in single precision (statistical comparator for the fp32 kernels). */
orc_sp_result orc_sum_product_f32(const orc_graph *g, const float *llr, const int32_t *syndrome,
                                  uint64_t max_it, int enable_threshold, float thr, int form, int32_t *bits_out);

/* ref: src/qkd_ldpc_algorithm.cpp:398-447. precision: 64 or 32. decoded_out may be NULL. */
orc_ldpc_result orc_qkd_ldpc(const orc_graph *g, const int32_t *alice, const int32_t *bob, double qber,
                             uint64_t max_it, int enable_threshold, double thr, int precision, int f32_form,
                             int32_t *syndrome_out, int32_t *decoded_out);

/* xoshiro256++ seeded through SplitMix64: third-party module Reputeless/Xoshiro-cpp v1.1
 * (ref: CMakeLists.txt:33-37), restated from the published algorithm. */
typedef struct orc_prng { uint64_t s[4]; } orc_prng;
void orc_prng_seed(orc_prng *p, uint64_t seed);
uint64_t orc_prng_next(orc_prng *p);

/* ref: src/simulation.cpp:222-228 (seeds[k] = one raw draw each) */
void orc_trial_seeds(uint64_t simulation_seed, uint64_t count, uint64_t *seeds_out);

/* ref: src/array_and_matrix_operations.cpp:424-431 + :434-460 with libstdc++ 13.3's
 * uniform_int_distribution / std::shuffle semantics restated (bits/uniform_int_dist.h:257-328,
 * bits/stl_algo.h:3742-3806). Returns the exact QBER floor(n*q)/n. */
double orc_generate(uint64_t seed, uint64_t n, double qber, int32_t *alice_out, int32_t *bob_out);

/* ref: src/simulation.cpp:161-189. Returns 0, or -1 when floor(n*q) == 0 (the reference throws). */
int orc_run_trial(const orc_graph *g, double qber, uint64_t seed, uint64_t max_it, int enable_threshold, double thr,
                  int precision, int f32_form, orc_ldpc_result *res_out, double *exact_qber_out);

/* `count` trials with OpenMP over `threads` threads; out3[3k+{0,1,2}] = {iterations, syndromes_match, keys_match}.
 * decoded_out (may be NULL) receives count*n decoded bits. */
int orc_run_trials(const orc_graph *g, double qber, const uint64_t *seeds, uint64_t count, int threads,
                   uint64_t max_it, int enable_threshold, double thr, int precision, int f32_form,
                   uint64_t *out3, int32_t *decoded_out);

/* ref: src/simulation.cpp:48-70. Returns the number of points, or -1 if none (the reference throws). */
int orc_qber_range(double code_rate, const double *params4, size_t n_params, double *out, size_t cap);

/* ref: src/simulation.cpp:252-312. stats_out = {mean, std_dev, min, max, ratio_sp, ratio_ldpc}. */
void orc_point_stats(const uint64_t *out3, uint64_t trials, uint64_t max_it, double *stats_out);

#ifdef __cplusplus
}
#endif
#endif
