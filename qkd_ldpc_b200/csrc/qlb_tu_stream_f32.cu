#include "qlb_internal.hpp"
#include <algorithm>
using namespace qlb;
#include "qlb_stream_f32.cuh"
#include "qlb_stream_split.cuh"

namespace
{
    // The frame-interleaved streaming kernel (qlb_stream_f32.cuh): messages in HBM, any block length.
    template <typename Rule, bool kReconcile, int kBW, int VEC, bool kTma>
    int launch_stream(qlb_ctx *ctx, DecodeArgs &args, int stages, long long max_ctas)
    {
        auto kern = decode_stream_f32_kernel<Rule, kReconcile, kBW, VEC, kTma>;
        const size_t ring = kTma ? (size_t)(kStreamThreads / 32) * stages * ((size_t)std::max(args.code.max_check_w, kBW) * 128 * VEC + 8) + 128 : 0;
        QLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring));
        const StreamCarve cv = stream_carve(args.code.n, args.code.m, args.code.slots, VEC);
        const long long G = 32 * VEC, groups = (args.n_frames + G - 1) / G;
        long long grid = std::min<long long>(ctx->sm_count, max_ctas); // one resident CTA per SM, fewer when HBM cannot hold more groups
        if (grid > groups)
            grid = groups;
        QLB_CUDA(ctx->scratch.reserve((size_t)grid * cv.total));
        if (args.syndrome_out)
            QLB_CUDA(cudaMemsetAsync(args.syndrome_out, 0, (size_t)args.n_frames * args.code.words_m * 4, ctx->stream));
        if (std::getenv("QLB_DEBUG"))
            std::fprintf(stderr, "[qlb] decode_stream_f32_kernel VEC=%d tma=%d stages=%d ring=%zu B: %lld groups of %lld frames, grid=%lld, %zu B scratch per group\n",
                         VEC, (int)kTma, stages, ring, groups, G, grid, cv.total);
        QLB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));
        args.queue = ctx->d_counters;
        args.iter_total = ctx->d_counters + 1;
        const char *pf = std::getenv("QLB_STREAM_PREFETCH"); // nodes ahead (per warp) whose rows are prefetched into L2
        kern<<<(unsigned)grid, kStreamThreads, ring, ctx->stream>>>(args, static_cast<unsigned char *>(ctx->scratch.p), cv.total, groups,
                                                                    pf ? std::atoi(pf) : 4, stages);
        QLB_CUDA(cudaGetLastError());
        ++ctx->launches;
        return QLB_OK;
    }

    // The phase-split form (qlb_stream_split.cuh): one kernel per pass over all groups, in waves when device memory cannot
    // hold the message arrays of every group at once. Nothing here waits for the device.
    template <typename Rule, bool kReconcile, int kBW, int VEC>
    int launch_stream_split(qlb_ctx *ctx, DecodeArgs &args, size_t budget)
    {
        const long long G = 32 * VEC, groups = (args.n_frames + G - 1) / G;
        // groups per bundle: 2 KB of contiguous memory per slot when there are enough groups (see qlb_stream_split.cuh)
        int B = VEC == 4 ? 4 : 1;
        if (const char *e = std::getenv("QLB_SPLIT_BUNDLE"))
            B = std::atoi(e);
        while (B > 1 && groups < B)
            B /= 2;
        if (B != 1 && B != 2 && B != 4 && B != 8)
            B = 1;
        const size_t per_bundle = split_bundle_bytes(args.code.n, args.code.m, args.code.slots, VEC, B);
        long long fit_bundles = (long long)(budget / per_bundle);
        if (const char *e = std::getenv("QLB_SPLIT_MAX_BUNDLES")) // test hook: force several waves without filling the device memory
            if (std::atoll(e) > 0)
                fit_bundles = std::min<long long>(fit_bundles, std::atoll(e));
        if (fit_bundles < 1)
            return fail(QLB_ERR_UNSUPPORTED, "streaming decoder: device memory cannot hold the messages of one bundle of frame groups");
        const long long bundles = (groups + B - 1) / B;
        const long long waves = (bundles + fit_bundles - 1) / fit_bundles;
        const long long bundles_per_wave = (bundles + waves - 1) / waves; // equal waves instead of full ones and a remainder
        const long long per_wave = bundles_per_wave * B;                  // groups
        if (args.n_frames >= 0xFFFFFFFFLL)
            return fail(QLB_ERR_UNSUPPORTED, "streaming decoder: more than 2^32 - 2 frames in one call");
        const size_t state_words = (size_t)per_wave * 12 + (size_t)bundles_per_wave + 16 + 3 * (size_t)per_wave * G;
        QLB_CUDA(ctx->scratch.reserve((size_t)bundles_per_wave * per_bundle + state_words * 4));
        SplitState st{};
        st.bundles = static_cast<unsigned char *>(ctx->scratch.p);
        st.bundle_stride = per_bundle;
        st.bundle = B;
        st.repack_pct = 65; // measured on B200, N = 100 000: 50 / 65 / 80 / 90 % -> 0.664 / 0.694 / 0.686 / 0.651 of peak at QBER 0.085
        if (const char *e = std::getenv("QLB_SPLIT_REPACK_PCT")) // experiments
            if (std::atoi(e) > 0 && std::atoi(e) < 100)
                st.repack_pct = std::atoi(e);
        uint32_t *words = reinterpret_cast<uint32_t *>(st.bundles + (size_t)bundles_per_wave * per_bundle);
        st.act = words;
        st.bad = words + 4 * per_wave;
        st.succ = words + 8 * per_wave;
        st.list = words + 12 * per_wave;
        st.n_live = st.list + bundles_per_wave;
        st.repack = st.n_live + 4;
        st.fmap = st.n_live + 16;
        st.src_of = st.fmap + (size_t)per_wave * G;
        st.fmap_new = st.src_of + (size_t)per_wave * G;
        if (args.syndrome_out)
            QLB_CUDA(cudaMemsetAsync(args.syndrome_out, 0, (size_t)args.n_frames * args.code.words_m * 4, ctx->stream));
        args.queue = ctx->d_counters;
        args.iter_total = ctx->d_counters + 1;

        auto k_setup = stream_setup_kernel<kReconcile, VEC>;
        auto k_init = stream_init_kernel<Rule, kReconcile, kBW, VEC>;
        auto k_check = stream_check_kernel<Rule, kReconcile, VEC>;
        auto k_update = stream_update_kernel<VEC>;
        auto k_bit = stream_bit_kernel<Rule, kReconcile, kBW, VEC>;
        auto k_final = stream_finalize_kernel<kReconcile, VEC>;
        auto k_plan = stream_repack_plan_kernel<VEC>;
        auto k_mvmsg = stream_repack_msg_kernel<VEC>;
        auto k_mvbits = stream_repack_bits_kernel<VEC>;
        auto k_commit = stream_repack_commit_kernel<VEC>;
        // rounds after which a repack is attempted (decided on the device: live columns <= half of the streamed ones)
        const bool repack_on = !std::getenv("QLB_SPLIT_NO_REPACK") && per_wave <= kMaxRepackGroups && per_wave >= 2;
        int repack_every = 4;
        if (const char *e = std::getenv("QLB_SPLIT_REPACK_EVERY")) // experiments
            if (std::atoi(e) > 0)
                repack_every = std::atoi(e);
        auto repack_round = [repack_every](int it) { return it >= 4 && it % repack_every == 2 % repack_every; }; // a declined attempt costs ~15 us
        int occ_check = 1, occ_bit = 1;
        QLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_check, k_check, kSplitCheckThreads, 0));
        QLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_bit, k_bit, kSplitBitThreads, 0));
        if (const char *e = std::getenv("QLB_SPLIT_OCC")) // experiments: "check,bit" resident CTAs per SM
        {
            int a = 0, b = 0;
            if (std::sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0)
                occ_check = std::min(occ_check, a), occ_bit = std::min(occ_bit, b);
        }
        const unsigned grid_check = (unsigned)(ctx->sm_count * std::max(1, occ_check)), grid_bit = (unsigned)(ctx->sm_count * std::max(1, occ_bit));
        if (std::getenv("QLB_DEBUG"))
            std::fprintf(stderr, "[qlb] stream split VEC=%d: %lld groups of %lld frames, bundles of %d, %lld wave(s) of <= %lld groups, %zu B per bundle; check grid %u (%d/SM), bit grid %u (%d/SM)\n",
                         VEC, groups, G, B, waves, per_wave, per_bundle, grid_check, occ_check, grid_bit, occ_bit);
        for (long long w = 0; w < waves; ++w)
        {
            st.group0 = w * per_wave;
            st.n_groups = (int)std::min<long long>(per_wave, groups - st.group0);
            if (st.n_groups <= 0)
                break;
            const unsigned grid_groups = (unsigned)std::min<long long>(st.n_groups, 2LL * ctx->sm_count);
            k_setup<<<grid_groups, kSplitSetupThreads, 0, ctx->stream>>>(args, st);
            k_init<<<grid_bit, kSplitBitThreads, 0, ctx->stream>>>(args, st);
            for (int it = 0;; ++it)
            {
                k_check<<<grid_check, kSplitCheckThreads, 0, ctx->stream>>>(args, st, it);
                k_update<<<1, 1024, 0, ctx->stream>>>(args, st, it);
                if (it == args.max_it)
                    break;
                if (repack_on && repack_round(it))
                {
                    k_plan<<<1, 1024, 0, ctx->stream>>>(st);
                    k_final<<<grid_groups, kSplitSetupThreads, 0, ctx->stream>>>(args, st, 1);
                    k_mvmsg<<<grid_bit, kRepackThreads, 0, ctx->stream>>>(args, st);
                    k_mvbits<<<grid_bit, kRepackThreads, 0, ctx->stream>>>(args, st);
                    k_commit<<<1, 1024, 0, ctx->stream>>>(st);
                    ctx->launches += 5;
                }
                k_bit<<<grid_bit, kSplitBitThreads, 0, ctx->stream>>>(args, st);
            }
            k_final<<<grid_groups, kSplitSetupThreads, 0, ctx->stream>>>(args, st, 0);
            QLB_CUDA(cudaGetLastError());
            ctx->launches += 3 + 3ULL * (unsigned)args.max_it + 2;
        }
        return QLB_OK;
    }

    template <typename Rule, bool kReconcile>
    int launch_stream_bw(qlb_ctx *ctx, DecodeArgs &args)
    {
        if (!std::getenv("QLB_STREAM_PERSISTENT") && args.code.uniform_bit_w == 3)
        {
            size_t free_b = 0, total_b = 0;
            QLB_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const size_t budget = (free_b + ctx->scratch.cap) / 20 * 17;
            return args.n_frames > 32 ? launch_stream_split<Rule, kReconcile, 3, 4>(ctx, args, budget) : launch_stream_split<Rule, kReconcile, 3, 1>(ctx, args, budget);
        }
        // 128-bit accesses (128 frames per group) unless the per-SM message arrays would not fit in device memory
        size_t free_b = 0, total_b = 0;
        QLB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        // 128 frames per group (128-bit accesses, 512-byte rows) as long as device memory holds the message arrays of enough
        // groups to keep at least 40 % of the SMs busy; else 32 frames per group (measured at N = 1 000 000: 0.23 of the HBM
        // copy bandwidth with 32-frame groups on all SMs)
        const size_t per_group4 = stream_carve(args.code.n, args.code.m, args.code.slots, 4).total;
        const size_t budget = (free_b + ctx->scratch.cap) / 20 * 17;
        const long long fit4 = (long long)(budget / per_group4), fit1 = (long long)(budget / stream_carve(args.code.n, args.code.m, args.code.slots, 1).total);
        const bool vec4 = args.n_frames > 32 && fit4 * 5 >= (long long)ctx->sm_count * 2;
        const long long max_ctas = std::max<long long>(1, vec4 ? fit4 : fit1);
        // TMA rings: 16 warps x S stages x (rows x row bytes) of shared memory; needs S >= 3 and check weights <= 8
        const size_t stage_bytes = (size_t)std::max(args.code.max_check_w, args.code.uniform_bit_w) * 128 * (vec4 ? 4 : 1) + 8;
        int stages = (int)std::min<size_t>(8, ((size_t)ctx->smem_optin - 4096) / ((kStreamThreads / 32) * stage_bytes));
        // Measured on B200 (N = 100 000, 18 944 frames): per-warp TMA rings of 512-byte bulk copies reach 0.49 of the HBM copy
        // bandwidth, plain 128-bit loads + L2 software prefetch 0.57-0.58 -- the rings are kept as an opt-in experiment.
        const bool tma = stages >= 3 && args.code.max_check_w <= 8 && args.code.uniform_bit_w == 3 && std::getenv("QLB_STREAM_TMA");
        if (tma)
            return vec4 ? launch_stream<Rule, kReconcile, 3, 4, true>(ctx, args, stages, max_ctas)
                        : launch_stream<Rule, kReconcile, 3, 1, true>(ctx, args, stages, max_ctas);
        switch (args.code.uniform_bit_w * 10 + (vec4 ? 4 : 1))
        {
        case 34: return launch_stream<Rule, kReconcile, 3, 4, false>(ctx, args, 0, max_ctas);
        case 31: return launch_stream<Rule, kReconcile, 3, 1, false>(ctx, args, 0, max_ctas);
        default: return fail(QLB_ERR_UNSUPPORTED, "streaming kernel: unsupported bit weight");
        }
    }

    bool stream_eligible_impl(const CodeDev &c)
    {
        // column weight 3 (the code family of BASELINE.json) only: other weights take the generic kernel
        return c.uniform_bit_w == 3 && c.max_check_w <= kResidentMaxCW && c.m <= c.n;
    }
}
namespace qlb
{
    bool stream_f32_eligible(const CodeDev &c) { return stream_eligible_impl(c); }
    int launch_stream_f32(qlb_ctx *ctx, DecodeArgs &args, bool reconcile, bool fast)
    {
        if (fast)
            return reconcile ? launch_stream_bw<RuleF32Fast, true>(ctx, args) : launch_stream_bw<RuleF32Fast, false>(ctx, args);
        return reconcile ? launch_stream_bw<RuleF32Accurate, true>(ctx, args) : launch_stream_bw<RuleF32Accurate, false>(ctx, args);
    }
}
