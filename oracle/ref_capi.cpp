// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI window onto the UNMODIFIED reference sources under /root/reference/src, which
// oracle/Makefile compiles in place into oracle/_ref/libqkdref.so. This file contains no
// reference code: it only includes the reference's own headers and forwards to its public
// functions, so that tests (ctypes) can (a) validate the C restatement in
// oracle/restatement/, (b) generate the golden fixtures under tests/golden/, and
// (c) time the reference's CPU path for bench.py's cpu_baseline / --impl reference legs.
//
// Forwarded reference entry points:
//   read_sparse_alist_matrix / read_dense_matrix   src/array_and_matrix_operations.cpp:109,295
//   generate_random_bit_array / introduce_errors   src/array_and_matrix_operations.cpp:424,434
//   calculate_syndrome_{regular,irregular}         src/array_and_matrix_operations.cpp:463,476
//   sum_product_decoding_{regular,irregular}       src/qkd_ldpc_algorithm.cpp:3,175
//   QKD_LDPC_{regular,irregular}                   src/qkd_ldpc_algorithm.cpp:347,398
//   run_trial                                      src/simulation.cpp:161
//   get_rate_based_QBER_range                      src/simulation.cpp:48
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>
#include <atomic>

#include "config.hpp"
#include "simulation.hpp"

config_data CFG; // the reference expects its executable's main TU to define this (src/main.cpp:13)

namespace
{
    thread_local std::string g_err;
    int fail(const std::exception &e)
    {
        g_err = e.what();
        return -1;
    }
}

extern "C"
{
    const char *ref_last_error() { return g_err.c_str(); }

    // Sets the fields of the global CFG that the hot path reads (src/qkd_ldpc_algorithm.cpp:246,313,424-425).
    void ref_set_cfg(uint64_t max_iterations, int enable_threshold, double threshold, uint64_t threads)
    {
        CFG.SUM_PRODUCT_MAX_ITERATIONS = max_iterations;
        CFG.ENABLE_SUM_PRODUCT_MSG_LLR_THRESHOLD = enable_threshold != 0;
        CFG.SUM_PRODUCT_MSG_LLR_THRESHOLD = threshold;
        CFG.THREADS_NUMBER = threads;
        CFG.TRACE_QKD_LDPC = CFG.TRACE_SUM_PRODUCT = CFG.TRACE_SUM_PRODUCT_LLR = false;
    }

    void *ref_matrix_load(const char *path, int dense)
    {
        H_matrix *h = new H_matrix();
        try
        {
            if (dense)
                read_dense_matrix(path, *h);
            else
                read_sparse_alist_matrix(path, *h);
        }
        catch (const std::exception &e)
        {
            fail(e);
            delete h; // members may be partially allocated; the reference leaks/frees inconsistently here, so do not touch them
            return nullptr;
        }
        return h;
    }

    void ref_matrix_free(void *hp)
    {
        H_matrix *h = static_cast<H_matrix *>(hp);
        if (!h)
            return;
        free_matrix_H(*h);
        delete h;
    }

    // info = {N, M, max_bit_w, max_check_w, is_regular, E_bits (sum of bit weights), E_checks (sum of check weights)}
    void ref_matrix_info(void *hp, uint64_t *info)
    {
        const H_matrix &h = *static_cast<H_matrix *>(hp);
        uint64_t eb = 0, ec = 0;
        for (size_t i = 0; i < h.num_bit_nodes; ++i)
            eb += h.bit_nodes_weight[i];
        for (size_t j = 0; j < h.num_check_nodes; ++j)
            ec += h.check_nodes_weight[j];
        info[0] = h.num_bit_nodes;
        info[1] = h.num_check_nodes;
        info[2] = h.max_bit_nodes_weight;
        info[3] = h.max_check_nodes_weight;
        info[4] = h.is_regular;
        info[5] = eb;
        info[6] = ec;
    }

    // Exports both adjacency halves exactly as stored (list order preserved): CSR over checks, CSC over bits.
    void ref_matrix_export(void *hp, int32_t *row_ptr, int32_t *col_idx, int32_t *col_ptr, int32_t *row_idx)
    {
        const H_matrix &h = *static_cast<H_matrix *>(hp);
        int32_t p = 0;
        for (size_t j = 0; j < h.num_check_nodes; ++j)
        {
            row_ptr[j] = p;
            for (int k = 0; k < h.check_nodes_weight[j]; ++k)
                col_idx[p++] = h.check_nodes[j][k];
        }
        row_ptr[h.num_check_nodes] = p;
        p = 0;
        for (size_t i = 0; i < h.num_bit_nodes; ++i)
        {
            col_ptr[i] = p;
            for (int k = 0; k < h.bit_nodes_weight[i]; ++k)
                row_idx[p++] = h.bit_nodes[i][k];
        }
        col_ptr[h.num_bit_nodes] = p;
    }

    // seeds[k] exactly as src/simulation.cpp:222-228 draws them.
    void ref_trial_seeds(uint64_t simulation_seed, uint64_t count, uint64_t *seeds_out)
    {
        XoshiroCpp::Xoshiro256PlusPlus prng(simulation_seed);
        std::uniform_int_distribution<size_t> distribution(0, std::numeric_limits<size_t>::max());
        for (uint64_t i = 0; i < count; ++i)
            seeds_out[i] = distribution(prng);
    }

    void ref_prng_raw(uint64_t seed, uint64_t count, uint64_t *out)
    {
        XoshiroCpp::Xoshiro256PlusPlus prng(seed);
        for (uint64_t i = 0; i < count; ++i)
            out[i] = prng();
    }

    // Key pair exactly as run_trial makes it (src/simulation.cpp:163-169). Returns the exact QBER.
    double ref_generate(uint64_t seed, uint64_t n, double qber, int32_t *alice_out, int32_t *bob_out)
    {
        XoshiroCpp::Xoshiro256PlusPlus prng(seed);
        generate_random_bit_array(prng, n, alice_out);
        return introduce_errors(prng, alice_out, n, qber, bob_out);
    }

    // variant: 0 = irregular functions, 1 = regular functions, -1 = dispatch on matrix.is_regular as run_trial does.
    static bool use_regular(const H_matrix &h, int variant) { return variant < 0 ? h.is_regular : variant == 1; }

    void ref_syndrome(void *hp, const int32_t *bits, int32_t *syndrome_out, int variant)
    {
        const H_matrix &h = *static_cast<H_matrix *>(hp);
        if (use_regular(h, variant))
            calculate_syndrome_regular(bits, h, syndrome_out);
        else
            calculate_syndrome_irregular(bits, h, syndrome_out);
    }

    // out = {iterations_num, syndromes_match}
    void ref_sum_product(void *hp, const double *llr, const int32_t *syndrome, uint64_t max_it, double thr,
                         int32_t *bits_out, uint64_t *out, int variant)
    {
        const H_matrix &h = *static_cast<H_matrix *>(hp);
        SP_result r = use_regular(h, variant) ? sum_product_decoding_regular(llr, h, syndrome, max_it, thr, bits_out)
                                              : sum_product_decoding_irregular(llr, h, syndrome, max_it, thr, bits_out);
        out[0] = r.iterations_num;
        out[1] = r.syndromes_match;
    }

    // out = {iterations_num, syndromes_match, keys_match}
    void ref_qkd_ldpc(void *hp, const int32_t *alice, const int32_t *bob, double qber, uint64_t *out, int variant)
    {
        const H_matrix &h = *static_cast<H_matrix *>(hp);
        LDPC_result r = use_regular(h, variant) ? QKD_LDPC_regular(alice, bob, qber, h) : QKD_LDPC_irregular(alice, bob, qber, h);
        out[0] = r.sp_res.iterations_num;
        out[1] = r.sp_res.syndromes_match;
        out[2] = r.keys_match;
    }

    // One reference trial. out = {iterations_num, syndromes_match, keys_match}; returns 0 or -1 (exception).
    int ref_run_trial(void *hp, double qber, uint64_t seed, uint64_t *out, double *exact_qber)
    {
        try
        {
            trial_result r = run_trial(*static_cast<H_matrix *>(hp), qber, seed);
            out[0] = r.ldpc_res.sp_res.iterations_num;
            out[1] = r.ldpc_res.sp_res.syndromes_match;
            out[2] = r.ldpc_res.keys_match;
            *exact_qber = r.initial_QBER;
            return 0;
        }
        catch (const std::exception &e)
        {
            return fail(e);
        }
    }

    // `count` reference trials over `threads` host threads (dynamic work distribution), trial k seeded with seeds[k].
    // This is the timed CPU arm: the per-trial work is exactly run_trial (generate + reconcile).
    // out3[k*3 + {0,1,2}] = {iterations_num, syndromes_match, keys_match}.
    int ref_run_trials(void *hp, double qber, const uint64_t *seeds, uint64_t count, uint64_t threads, uint64_t *out3)
    {
        const H_matrix &h = *static_cast<H_matrix *>(hp);
        std::atomic<uint64_t> next{0};
        std::atomic<int> status{0};
        auto body = [&]()
        {
            for (;;)
            {
                uint64_t k = next.fetch_add(1);
                if (k >= count)
                    return;
                try
                {
                    trial_result r = run_trial(h, qber, seeds[k]);
                    out3[3 * k + 0] = r.ldpc_res.sp_res.iterations_num;
                    out3[3 * k + 1] = r.ldpc_res.sp_res.syndromes_match;
                    out3[3 * k + 2] = r.ldpc_res.keys_match;
                }
                catch (const std::exception &e)
                {
                    fail(e);
                    status = -1;
                    return;
                }
            }
        };
        if (threads < 1)
            threads = 1;
        std::vector<std::thread> pool;
        for (uint64_t t = 1; t < threads; ++t)
            pool.emplace_back(body);
        body();
        for (auto &t : pool)
            t.join();
        return status.load();
    }

    // QBER grid for a code rate, via the reference's own function. Returns the number of points (<= cap) or -1.
    int ref_qber_range(double code_rate, const double *params4, uint64_t n_params, double *out, uint64_t cap)
    {
        try
        {
            std::vector<R_QBER_params> p;
            for (uint64_t i = 0; i < n_params; ++i)
                p.push_back({params4[4 * i], params4[4 * i + 1], params4[4 * i + 2], params4[4 * i + 3]});
            std::vector<double> q = get_rate_based_QBER_range(code_rate, p);
            uint64_t n = q.size() < cap ? q.size() : cap;
            std::memcpy(out, q.data(), n * sizeof(double));
            return static_cast<int>(q.size());
        }
        catch (const std::exception &e)
        {
            return fail(e);
        }
    }
}
