// The config.json-driven QBER sweep as a frame-batch scheduler.
//
// Mirrors the reference's src/simulation.cpp: write_file :4-44, get_rate_based_QBER_range :48-70, prepare_sim_inputs
// :140-158, run_trial :161-189, QKD_LDPC_batch_simulation :192-316 -- same signatures, same trial seeds
// (seeds[k] + index of the (matrix, QBER) point), same statistics, same CSV. What changes is the execution model: instead
// of one pool task per trial that generates AND decodes on a CPU core, host threads only generate keys (bit-packed, in
// batches), and every batch is reconciled by one fused kernel launch on a GPU. Trials are independent, so batches are
// dealt to all GPUs of the box with no exchange while decoding; per-point statistics are integer histograms, reduced
// over the GPUs with one NCCL all-reduce per point (qlb_stats_allreduce).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <mutex>
#include <random>
#include <stdexcept>
#include <thread>

#include "device_bridge.hpp"
#include "qkd_ldpc.hpp"
#include "trace_print.hpp"

namespace
{
    qkd_b200::sweep_report g_report;

    // ---- a plain worker pool for key generation --------------------------------------------------------------------
    class worker_pool
    {
    public:
        explicit worker_pool(size_t n)
        {
            for (size_t i = 0; i < std::max<size_t>(1, n); ++i)
                threads_.emplace_back([this] { run(); });
        }
        ~worker_pool()
        {
            {
                std::lock_guard<std::mutex> lk(mu_);
                stop_ = true;
            }
            cv_.notify_all();
            for (auto &t : threads_)
                t.join();
        }
        void submit(std::function<void()> job)
        {
            {
                std::lock_guard<std::mutex> lk(mu_);
                jobs_.push_back(std::move(job));
            }
            cv_.notify_one();
        }

    private:
        void run()
        {
            for (;;)
            {
                std::function<void()> job;
                {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
                    if (jobs_.empty())
                        return;
                    job = std::move(jobs_.front());
                    jobs_.pop_front();
                }
                job();
            }
        }
        std::vector<std::thread> threads_;
        std::deque<std::function<void()>> jobs_;
        std::mutex mu_;
        std::condition_variable cv_;
        bool stop_ = false;
    };

    // One batch of frames travelling host -> GPU -> host.
    struct batch
    {
        size_t first_trial = 0, frames = 0;
        bool on_device = false;           // keys are generated on the GPU from seeds[first_trial ...] + seed_offset
        uint64_t seed_offset = 0;
        double requested_qber = 0;
        std::vector<uint32_t> alice, bob; // packed keys
        std::vector<double> qber;         // exact QBER per frame
        std::vector<uint32_t> iterations;
        std::vector<uint8_t> result;
        std::atomic<size_t> parts_left{0};
    };

    // Thread-safe queue of batch pointers (nullptr = shut down).
    class batch_queue
    {
    public:
        void push(batch *b)
        {
            {
                std::lock_guard<std::mutex> lk(mu_);
                q_.push_back(b);
            }
            cv_.notify_one();
        }
        batch *pop()
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return !q_.empty(); });
            batch *b = q_.front();
            q_.pop_front();
            return b;
        }

    private:
        std::deque<batch *> q_;
        std::mutex mu_;
        std::condition_variable cv_;
    };

    // Keys of trial `seed` exactly as the reference's run_trial makes them (src/simulation.cpp:163-169), bit-packed.
    double make_frame(size_t n, double qber, size_t seed, std::vector<int> &alice, std::vector<int> &bob, uint32_t *alice_words, uint32_t *bob_words)
    {
        XoshiroCpp::Xoshiro256PlusPlus prng(seed);
        generate_random_bit_array(prng, n, alice.data());
        const double exact = introduce_errors(prng, alice.data(), n, qber, bob.data());
        qkd_b200::pack_bits(alice.data(), n, alice_words);
        qkd_b200::pack_bits(bob.data(), n, bob_words);
        return exact;
    }

    [[noreturn]] void key_too_small(size_t n) { throw std::runtime_error("Key size '" + std::to_string(n) + "' is too small for QBER."); }
}

namespace qkd_b200
{
    const sweep_report &last_sweep_report() { return g_report; }

    void write_report(const sweep_report &report, fs::path directory)
    {
        if (!fs::exists(directory))
            fs::create_directories(directory);
        const std::string stem = "throughput(trial_num=" + std::to_string(CFG.TRIALS_NUMBER) + ",max_sum_prod_iters=" +
                                 std::to_string(CFG.SUM_PRODUCT_MAX_ITERATIONS) + ",seed=" + std::to_string(CFG.SIMULATION_SEED) + ")";
        fs::path target = directory / (stem + ".csv");
        for (size_t dup = 1; fs::exists(target); ++dup)
            target = directory / (stem + "_" + std::to_string(dup) + ".csv");
        std::ofstream out(target, std::ios::out | std::ios::trunc);
        out << "SIM;MATRIX_FILENAME;M;N;QBER;FRAMES;SECONDS;FRAMES_PER_S;SIFTED_MBIT_PER_S;FRAME_ITERATIONS;MEAN_ITERATIONS;"
               "EFFICIENCY_F;LEAKED_BITS_PER_FRAME;GPUS;PRECISION\n";
        const std::string precision = CFG.DEVICE_PRECISION == 32 ? (CFG.DEVICE_FP32_FAST ? "fp32-fast" : "fp32") : (CFG.DEVICE_FP64_FUSED ? "fp64-fused" : "fp64");
        for (const point_report &p : report.points)
        {
            const double q = p.exact_qber, h2 = -q * std::log2(q) - (1. - q) * std::log2(1. - q);
            const double fps = p.seconds > 0 ? p.frames / p.seconds : 0.;
            out << p.sim_number << ";" << p.matrix_filename << ";" << p.num_check_nodes << ";" << p.num_bit_nodes << ";" << q << ";" << p.frames << ";"
                << p.seconds << ";" << fps << ";" << fps * p.num_bit_nodes / 1e6 << ";" << p.frame_iterations << ";"
                << static_cast<double>(p.frame_iterations) / p.frames << ";" << (static_cast<double>(p.num_check_nodes) / p.num_bit_nodes) / h2 << ";"
                << p.num_check_nodes << ";" << report.gpus << ";" << precision << "\n";
        }
    }

    std::vector<std::string> matrix_warnings(const H_matrix &h)
    {
        std::vector<std::string> w;
        size_t unsorted_bits = 0, unsorted_checks = 0, dup = 0;
        for (size_t i = 0; i < h.num_bit_nodes; ++i)
            for (int k = 1; k < h.bit_nodes_weight[i]; ++k)
            {
                unsorted_bits += h.bit_nodes[i][k] < h.bit_nodes[i][k - 1];
                dup += h.bit_nodes[i][k] == h.bit_nodes[i][k - 1];
            }
        for (size_t j = 0; j < h.num_check_nodes; ++j)
            for (int k = 1; k < h.check_nodes_weight[j]; ++k)
                unsorted_checks += h.check_nodes[j][k] < h.check_nodes[j][k - 1];
        if (unsorted_bits || unsorted_checks)
            w.push_back("adjacency lists are not sorted ascending (" + std::to_string(unsorted_bits) + " bit-list and " + std::to_string(unsorted_checks) +
                        " check-list inversions): the reference routes messages by arrival order, which is only the intended routing for sorted lists");
        if (dup)
            w.push_back(std::to_string(dup) + " duplicate entries in the bit lists (a double edge cancels in the syndrome but not in the decoder)");
        return w;
    }
}

// CSV, byte for byte the reference's format (default ostream precision, ';' separated, FER = 1 - ratio_ldpc).
void write_file(const std::vector<sim_result> &data, fs::path directory)
{
    try
    {
        if (!fs::exists(directory))
            fs::create_directories(directory);
        const std::string stem = "ldpc(trial_num=" + std::to_string(CFG.TRIALS_NUMBER) + ",max_sum_prod_iters=" +
                                 std::to_string(CFG.SUM_PRODUCT_MAX_ITERATIONS) + ",seed=" + std::to_string(CFG.SIMULATION_SEED) + ")";
        fs::path target = directory / (stem + ".csv");
        for (size_t dup = 1; fs::exists(target); ++dup)
            target = directory / (stem + "_" + std::to_string(dup) + ".csv");

        std::ofstream out(target, std::ios::out | std::ios::trunc);
        out << "№;MATRIX_FILENAME;TYPE;CODE_RATE;M;N;QBER;ITERATIONS_SUCCESSFUL_SP_MEAN;ITERATIONS_SUCCESSFUL_SP_STD_DEV;"
               "ITERATIONS_SUCCESSFUL_SP_MIN;ITERATIONS_SUCCESSFUL_SP_MAX;RATIO_TRIALS_SUCCESSFUL_SP;RATIO_TRIALS_SUCCESSFUL_LDPC;FER\n";
        for (const sim_result &r : data)
            out << r.sim_number << ";" << r.matrix_filename << ";" << (r.is_regular ? "regular" : "irregular") << ";"
                << 1. - (static_cast<double>(r.num_check_nodes) / r.num_bit_nodes) << ";" << r.num_check_nodes << ";" << r.num_bit_nodes << ";"
                << r.initial_QBER << ";" << r.iterations_successful_sp_mean << ";" << r.iterations_successful_sp_std_dev << ";"
                << r.iterations_successful_sp_min << ";" << r.iterations_successful_sp_max << ";" << r.ratio_trials_successful_sp << ";"
                << r.ratio_trials_successful_ldpc << ";" << 1. - r.ratio_trials_successful_ldpc << "\n";
    }
    catch (const std::exception &)
    {
        std::cerr << "An error occurred while writing to the file.\n";
        throw;
    }
}

// First preset (sorted by rate) whose code_rate is >= the matrix's; end-exclusive grid begin + j*step, j < round((end-begin)/step).
std::vector<double> get_rate_based_QBER_range(const double code_rate, const std::vector<R_QBER_params> &R_QBER_parameters)
{
    for (const R_QBER_params &p : R_QBER_parameters)
    {
        if (code_rate > p.code_rate)
            continue;
        const size_t steps = round((p.QBER_end - p.QBER_begin) / p.QBER_step);
        std::vector<double> grid;
        for (size_t j = 0; j < steps; ++j)
            grid.push_back(p.QBER_begin + j * p.QBER_step);
        if (!grid.empty())
            return grid;
        break;
    }
    throw std::runtime_error("An error occurred when generating a QBER range based on code rate.");
}

void prepare_sim_inputs(const std::vector<fs::path> &matrix_paths, std::vector<sim_input> &sim_inputs_out)
{
    for (size_t i = 0; i < matrix_paths.size(); ++i)
    {
        sim_input &in = sim_inputs_out[i];
        if (CFG.USE_DENSE_MATRICES)
            read_dense_matrix(matrix_paths[i], in.matrix);
        else
            read_sparse_alist_matrix(matrix_paths[i], in.matrix);
        in.matrix_path = matrix_paths[i];
        for (const std::string &warning : qkd_b200::matrix_warnings(in.matrix))
            std::cerr << "WARNING (" << matrix_paths[i].filename().string() << "): " << warning << "\n";
        const double code_rate = 1. - (static_cast<double>(in.matrix.num_check_nodes) / in.matrix.num_bit_nodes);
        in.QBER = get_rate_based_QBER_range(code_rate, CFG.R_QBER_PARAMETERS);
    }
}

// One frame per QBER value of the matrix picked on stdin, printed step by step: the reference's interactive mode
// (src/simulation.cpp:73-137, src/utils.cpp:50-67), each frame decoded on the GPU.
fs::path select_matrix_file(const std::vector<fs::path> &matrix_paths)
{
    using qkd_b200::colour;
    qkd_b200::print_coloured(colour::green, "Choose file: \n");
    for (size_t i = 0; i < matrix_paths.size(); i++)
        qkd_b200::print_coloured(colour::green, std::to_string(i + 1) + ". " + matrix_paths[i].filename().string() + "\n");
    std::fflush(stdout);
    int file_index = 0;
    std::cin >> file_index;
    file_index -= 1;
    if (file_index < 0 || file_index >= static_cast<int>(matrix_paths.size()))
        throw std::runtime_error("Wrong file number.");
    return matrix_paths[file_index];
}

void QKD_LDPC_interactive_simulation(fs::path matrix_dir_path)
{
    using qkd_b200::colour;
    using qkd_b200::print_coloured;
    H_matrix matrix;
    const std::vector<fs::path> matrix_paths = get_file_paths_in_directory(matrix_dir_path);
    const fs::path matrix_path = select_matrix_file(matrix_paths);
    if (CFG.USE_DENSE_MATRICES)
        read_dense_matrix(matrix_path, matrix);
    else
        read_sparse_alist_matrix(matrix_path, matrix);
    try
    {
        print_coloured(colour::green, std::string(matrix.is_regular ? "Matrix H is regular." : "Matrix H is irregular.") + "\n");
        const size_t n = matrix.num_bit_nodes;
        std::vector<int> alice(n), bob(n);
        XoshiroCpp::Xoshiro256PlusPlus prng(CFG.SIMULATION_SEED); // one stream for all frames, as in the reference (:95)
        const double code_rate = 1. - (static_cast<double>(matrix.num_check_nodes) / matrix.num_bit_nodes);
        const std::vector<double> QBER = get_rate_based_QBER_range(code_rate, CFG.R_QBER_PARAMETERS);
        for (size_t i = 0; i < QBER.size(); i++)
        {
            print_coloured(colour::green, "\u2116:" + std::to_string(i + 1) + "\n");
            generate_random_bit_array(prng, n, alice.data());
            const double initial_QBER = introduce_errors(prng, alice.data(), n, QBER[i], bob.data());
            print_coloured(colour::green, "Actual QBER: " + qkd_b200::format_shortest(initial_QBER) + "\n");
            if (initial_QBER == 0.)
                key_too_small(n);
            int error_num = 0;
            for (size_t k = 0; k < n; k++)
                error_num += alice[k] ^ bob[k];
            print_coloured(colour::green, "Number of errors in a key: " + std::to_string(error_num) + "\n");
            std::fflush(stdout);
            const LDPC_result r = matrix.is_regular ? QKD_LDPC_regular(alice.data(), bob.data(), initial_QBER, matrix)
                                                    : QKD_LDPC_irregular(alice.data(), bob.data(), initial_QBER, matrix);
            print_coloured(colour::green, "Iterations performed: " + std::to_string(r.sp_res.iterations_num) + "\n");
            print_coloured(colour::green, std::string(r.keys_match && r.sp_res.syndromes_match ? "Error reconciliation SUCCESSFUL" : "Error reconciliation FAILED") + "\n\n");
            std::fflush(stdout);
        }
    }
    catch (...)
    {
        free_matrix_H(matrix);
        throw;
    }
    free_matrix_H(matrix);
}

// A single trial (batch of one). Throws like the reference when floor(N * QBER) == 0.
trial_result run_trial(const H_matrix &matrix, const double QBER, size_t seed)
{
    const size_t n = matrix.num_bit_nodes;
    XoshiroCpp::Xoshiro256PlusPlus prng(seed);
    std::vector<int> alice(n), bob(n);
    trial_result result;
    generate_random_bit_array(prng, n, alice.data());
    result.initial_QBER = introduce_errors(prng, alice.data(), n, QBER, bob.data());
    if (result.initial_QBER == 0.)
        key_too_small(n);
    result.ldpc_res = matrix.is_regular ? QKD_LDPC_regular(alice.data(), bob.data(), result.initial_QBER, matrix)
                                        : QKD_LDPC_irregular(alice.data(), bob.data(), result.initial_QBER, matrix);
    return result;
}

std::vector<sim_result> QKD_LDPC_batch_simulation(const std::vector<sim_input> &sim_in)
{
    using clock = std::chrono::steady_clock;
    const auto t_start = clock::now();
    const bool timing = std::getenv("QKD_B200_TIMING") != nullptr; // coarse wall-clock marks on stderr
    auto mark = [&](const char *what)
    {
        if (timing)
            std::fprintf(stderr, "[timing] %8.3f s  %s\n", std::chrono::duration<double>(clock::now() - t_start).count(), what);
    };
    const size_t trials = CFG.TRIALS_NUMBER;
    size_t points_total = 0;
    for (const sim_input &in : sim_in)
        points_total += in.QBER.size();

    // seeds[k]: the k-th raw output of xoshiro256++(SIMULATION_SEED), drawn the way the reference draws them
    XoshiroCpp::Xoshiro256PlusPlus seed_prng(CFG.SIMULATION_SEED);
    std::uniform_int_distribution<size_t> any_size(0, std::numeric_limits<size_t>::max());
    std::vector<size_t> seeds(trials);
    for (size_t &s : seeds)
        s = any_size(seed_prng);

    mark("trial seeds drawn");
    int gpus = qkd_b200::usable_devices();
    if (gpus < 1)
        throw std::runtime_error("no CUDA device is available: this build has no CPU decoder");
    if (CFG.DEVICE_GPUS > 0)
        gpus = std::min(gpus, CFG.DEVICE_GPUS);
    // device_batch_frames is the upper bound; with many GPUs the batches shrink so that every worker still gets ~8 of them per
    // QBER point (the point ends with a barrier: 61 batches over 16 workers would leave a quarter of them idle at the end)
    const size_t batch_cap = std::max<size_t>(1, CFG.DEVICE_BATCH_FRAMES);
    const size_t batch_frames = std::min(batch_cap, std::max<size_t>(1024, trials / (static_cast<size_t>(gpus) * 2 * 8) + 1));
    const size_t max_it = CFG.SUM_PRODUCT_MAX_ITERATIONS;
    const qlb_decode_params params = qkd_b200::params_from_cfg(max_it, CFG.SUM_PRODUCT_MSG_LLR_THRESHOLD);
    const size_t stats_width = max_it + 1 + 4; // histogram of iterations of successful frames + {n_sp, n_ldpc, n_trials, sum_iterations}

    worker_pool generators(CFG.THREADS_NUMBER);
    const size_t gen_parts = std::max<size_t>(1, CFG.THREADS_NUMBER);
    // two host threads (and contexts/streams) per GPU: one batch's key generation and transfers overlap another's decode
    const int workers_per_gpu = 2, workers = gpus * workers_per_gpu;
    std::vector<batch> pool(static_cast<size_t>(workers) * 2 + 1);
    batch_queue free_batches, ready;
    for (batch &b : pool)
        free_batches.push(&b);

    // ---- GPU workers: one host thread + one context per GPU ---------------------------------------------------------
    std::vector<trial_result> trial_results(trials);
    std::vector<std::vector<uint64_t>> gpu_stats(workers, std::vector<uint64_t>(stats_width, 0)); // per worker, folded per GPU below
    std::vector<qlb_ctx *> contexts(workers, nullptr);
    std::atomic<size_t> batches_done{0};
    std::mutex err_mu, done_mu;
    std::condition_variable done_cv;
    std::string first_error;
    qlb_code *code = nullptr;                 // the current matrix
    std::atomic<uint64_t> device_ns{0}, ready_ns{0};
    std::atomic<int> contexts_ready{0};
    std::vector<std::thread> gpu_threads;
    // integer statistics of one finished trial: histogram of iterations of successful frames + {n_sp, n_ldpc, n_trials, sum_iterations}
    auto account = [max_it](std::vector<uint64_t> &st, const trial_result &tr)
    {
        const size_t it = tr.ldpc_res.sp_res.iterations_num;
        if (tr.ldpc_res.sp_res.syndromes_match)
        {
            ++st[std::min<size_t>(it, max_it)];
            ++st[max_it + 1];
            st[max_it + 2] += tr.ldpc_res.keys_match;
        }
        ++st[max_it + 3];
        st[max_it + 4] += it;
    };
    // With any console trace enabled the trials of a point run one after another on this thread through run_trial, so the
    // output reads like the reference's with threads_number = 1 (its pool would interleave the prints of concurrent trials).
    const bool traced = CFG.TRACE_QKD_LDPC || CFG.TRACE_SUM_PRODUCT || CFG.TRACE_SUM_PRODUCT_LLR;
    for (int g = 0; g < workers; ++g)
        gpu_threads.emplace_back([&, g]
                                 {
            try { contexts[g] = qkd_b200::context(g % gpus); } // worker g drives GPU g % gpus with its own context
            catch (const std::exception &e) { std::lock_guard<std::mutex> lk(err_mu); if (first_error.empty()) first_error = e.what(); }
            if (++contexts_ready == workers)
            {
                ready_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(clock::now() - t_start).count();
                mark("every worker has its context");
            }
            for (;;)
            {
                batch *b = ready.pop();
                if (!b)
                    return;
                if (contexts[g] && first_error.empty())
                {
                    const auto t0 = clock::now();
                    int rc;
                    if (b->on_device)
                    {
                        double exact = 0;
                        rc = qlb_run_trials(contexts[g], code, &params, static_cast<int64_t>(b->frames), reinterpret_cast<const uint64_t *>(&seeds[b->first_trial]),
                                            b->seed_offset, b->requested_qber, b->iterations.data(), b->result.data(), &exact);
                        std::fill(b->qber.begin(), b->qber.end(), exact);
                    }
                    else
                        rc = qlb_reconcile_batch_packed(contexts[g], code, &params, static_cast<int64_t>(b->frames), b->alice.data(), b->bob.data(),
                                                        b->qber.data(), b->iterations.data(), b->result.data(), nullptr, nullptr);
                    device_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(clock::now() - t0).count();
                    if (rc != QLB_OK)
                    {
                        std::lock_guard<std::mutex> lk(err_mu);
                        if (first_error.empty())
                            first_error = qlb_last_error();
                    }
                    else
                    {
                        for (size_t f = 0; f < b->frames; ++f)
                        {
                            trial_result &tr = trial_results[b->first_trial + f];
                            tr.ldpc_res.sp_res.iterations_num = b->iterations[f];
                            tr.ldpc_res.sp_res.syndromes_match = (b->result[f] & QLB_RES_SYNDROMES_MATCH) != 0;
                            tr.ldpc_res.keys_match = (b->result[f] & QLB_RES_KEYS_MATCH) != 0;
                            tr.initial_QBER = b->qber[f];
                            account(gpu_stats[g], tr);
                        }
                    }
                }
                free_batches.push(b);
                {
                    std::lock_guard<std::mutex> lk(done_mu);
                    ++batches_done;
                }
                done_cv.notify_all();
            } });
    // The sweep's ONE collective is a single NCCL all-reduce of the integer statistics of every point, at the end. Creating
    // the communicators takes seconds on an 8-GPU box, so it happens on a side thread (with contexts of its own) while the
    // GPUs decode.
    const bool reduce_over_nccl = gpus > 1 || std::getenv("QKD_B200_FORCE_ALLREDUCE");
    std::vector<std::vector<uint64_t>> sweep_stats(gpus, std::vector<uint64_t>(points_total * stats_width, 0));
    std::string warm_error;
    std::thread nccl_warm_up;
    if (reduce_over_nccl)
        nccl_warm_up = std::thread([&]
                                   {
            try
            {
                std::vector<qlb_ctx *> ctxs;
                std::vector<uint64_t> zero(gpus, 0);
                std::vector<uint64_t *> ptrs;
                for (int g = 0; g < gpus; ++g)
                {
                    ctxs.push_back(qkd_b200::context(g));
                    ptrs.push_back(&zero[g]);
                }
                qkd_b200::check(qlb_stats_allreduce(ctxs.data(), gpus, ptrs.data(), 1), "qlb_stats_allreduce (communicator set-up)");
                mark("NCCL communicators ready");
            }
            catch (const std::exception &e) { warm_error = e.what(); } });
    auto shut_down = [&]
    {
        for (int g = 0; g < workers; ++g)
            ready.push(nullptr);
        for (auto &t : gpu_threads)
            t.join();
        if (nccl_warm_up.joinable())
            nccl_warm_up.join();
    };

    std::vector<sim_result> sim_results(points_total);
    size_t curr_sim = 0, frames_total = 0, iterations_total = 0;
    std::vector<size_t> point_ok_sp, point_ok_ldpc;
    g_report = qkd_b200::sweep_report{};
    try
    {
        for (const sim_input &in : sim_in)
        {
            const H_matrix &matrix = in.matrix;
            const size_t n = matrix.num_bit_nodes, words = (n + 31) / 32;
            code = qkd_b200::code_for(matrix);
            mark("code layout ready");
            const std::string matrix_filename = in.matrix_path.filename().string();
            for (const double QBER : in.QBER)
            {
                if (static_cast<size_t>(n * QBER) == 0)
                    key_too_small(n); // the reference throws from inside the first trial (src/simulation.cpp:170-175)
                const auto t_point = clock::now();
                for (auto &st : gpu_stats)
                    std::fill(st.begin(), st.end(), 0);
                batches_done = 0;
                size_t issued = 0;
                for (size_t k = 0; traced && k < trials; ++k)
                {
                    trial_results[k] = run_trial(matrix, QBER, seeds[k] + curr_sim);
                    account(gpu_stats[0], trial_results[k]);
                }
                for (size_t first = 0; !traced && first < trials; first += batch_frames, ++issued)
                {
                    batch *b = free_batches.pop();
                    b->first_trial = first;
                    b->frames = std::min(batch_frames, trials - first);
                    b->qber.resize(b->frames);
                    b->iterations.resize(b->frames);
                    b->result.resize(b->frames);
                    b->on_device = CFG.DEVICE_GENERATE_KEYS;
                    b->seed_offset = curr_sim;
                    b->requested_qber = QBER;
                    if (b->on_device)
                    {
                        ready.push(b); // nothing to prepare on the host: the trial seeds are the input
                        continue;
                    }
                    b->alice.resize(b->frames * words);
                    b->bob.resize(b->frames * words);
                    const size_t parts = std::min(gen_parts, b->frames);
                    b->parts_left = parts;
                    for (size_t part = 0; part < parts; ++part)
                        generators.submit([&, b, part, parts, n, words, QBER, curr_sim]
                                          {
                            std::vector<int> alice(n), bob(n);
                            const size_t lo = b->frames * part / parts, hi = b->frames * (part + 1) / parts;
                            for (size_t f = lo; f < hi; ++f)
                                b->qber[f] = make_frame(n, QBER, seeds[b->first_trial + f] + curr_sim, alice, bob, &b->alice[f * words], &b->bob[f * words]);
                            if (--b->parts_left == 0)
                                ready.push(b); });
                }
                {
                    std::unique_lock<std::mutex> lk(done_mu);
                    done_cv.wait(lk, [&] { return batches_done == issued; });
                }
                if (!first_error.empty())
                    throw std::runtime_error(first_error);
                mark("point decoded");

                // fold the second worker of each GPU into the first and keep the point's per-GPU integer statistics for the
                // all-reduce at the end of the sweep
                for (int w = gpus; w < workers; ++w)
                    for (size_t x = 0; x < stats_width; ++x)
                        gpu_stats[w % gpus][x] += gpu_stats[w][x];
                for (int g = 0; g < gpus; ++g)
                    std::copy(gpu_stats[g].begin(), gpu_stats[g].end(), sweep_stats[g].begin() + static_cast<std::ptrdiff_t>(curr_sim * stats_width));

                // statistics exactly as the reference accumulates them, in trial order (src/simulation.cpp:252-312)
                size_t ok_sp = 0, ok_ldpc = 0, it_max = 0, it_min = max_it;
                double mean = 0, sd = 0;
                for (const trial_result &tr : trial_results)
                    if (tr.ldpc_res.sp_res.syndromes_match)
                    {
                        const size_t it = tr.ldpc_res.sp_res.iterations_num;
                        ++ok_sp;
                        it_max = std::max(it_max, it);
                        it_min = std::min(it_min, it);
                        ok_ldpc += tr.ldpc_res.keys_match;
                        mean += static_cast<double>(it);
                    }
                if (ok_sp > 0)
                {
                    mean /= static_cast<double>(ok_sp);
                    for (const trial_result &tr : trial_results)
                        if (tr.ldpc_res.sp_res.syndromes_match)
                            sd += pow(static_cast<double>(tr.ldpc_res.sp_res.iterations_num) - mean, 2);
                    sd = sqrt(sd / static_cast<double>(ok_sp));
                }
                point_ok_sp.push_back(ok_sp);
                point_ok_ldpc.push_back(ok_ldpc);

                sim_result &r = sim_results[curr_sim];
                r.sim_number = curr_sim;
                r.matrix_filename = matrix_filename;
                r.is_regular = matrix.is_regular;
                r.num_bit_nodes = matrix.num_bit_nodes;
                r.num_check_nodes = matrix.num_check_nodes;
                r.initial_QBER = trial_results[0].initial_QBER;
                r.iterations_successful_sp_max = it_max;
                r.iterations_successful_sp_min = (it_min == max_it) ? 0 : it_min;
                r.iterations_successful_sp_mean = mean;
                r.iterations_successful_sp_std_dev = sd;
                r.ratio_trials_successful_ldpc = static_cast<double>(ok_ldpc) / trials;
                r.ratio_trials_successful_sp = static_cast<double>(ok_sp) / trials;
                frames_total += trials;
                qkd_b200::point_report pr;
                pr.sim_number = curr_sim;
                pr.matrix_filename = matrix_filename;
                pr.num_bit_nodes = matrix.num_bit_nodes;
                pr.num_check_nodes = matrix.num_check_nodes;
                pr.exact_qber = r.initial_QBER;
                pr.frames = trials;
                pr.frame_iterations = 0; // filled from the reduced statistics below
                pr.seconds = std::chrono::duration<double>(clock::now() - t_point).count();
                g_report.points.push_back(pr);
                mark("point statistics done");
                ++curr_sim;
            }
        }
    }
    catch (...)
    {
        shut_down();
        throw;
    }
    shut_down();

    // ---- the sweep's one collective: SUM all-reduce of [points x (max_it + 5)] integers over the GPUs of this box -----------
    if (reduce_over_nccl && !traced)
    {
        if (!warm_error.empty())
            throw std::runtime_error(warm_error);
        std::vector<qlb_ctx *> ctxs;
        std::vector<uint64_t *> ptrs;
        for (int g = 0; g < gpus; ++g)
        {
            ctxs.push_back(qkd_b200::context(g));
            ptrs.push_back(sweep_stats[g].data());
        }
        qkd_b200::check(qlb_stats_allreduce(ctxs.data(), gpus, ptrs.data(), sweep_stats[0].size()), "qlb_stats_allreduce");
        mark("statistics all-reduced");
    }
    else
        for (int g = 1; g < gpus; ++g)
            for (size_t x = 0; x < sweep_stats[0].size(); ++x)
                sweep_stats[0][x] += sweep_stats[g][x];
    for (size_t pt = 0; pt < curr_sim; ++pt)
    {
        const uint64_t *reduced = sweep_stats[0].data() + pt * stats_width;
        if (reduced[max_it + 1] != point_ok_sp[pt] || reduced[max_it + 2] != point_ok_ldpc[pt] || reduced[max_it + 3] != trials)
            throw std::runtime_error("reduced statistics disagree with the per-trial results");
        g_report.points[pt].frame_iterations = reduced[max_it + 4];
        iterations_total += reduced[max_it + 4];
    }

    g_report.seconds_total = std::chrono::duration<double>(clock::now() - t_start).count();
    g_report.seconds_device = device_ns.load() * 1e-9 / workers;
    g_report.seconds_startup = ready_ns.load() * 1e-9; // CUDA initialisation + contexts: ~1 s on a 1-GPU box, ~8 s on an 8-GPU box
    g_report.frames = frames_total;
    g_report.frame_iterations = iterations_total;
    g_report.gpus = gpus;
    return sim_results;
}
