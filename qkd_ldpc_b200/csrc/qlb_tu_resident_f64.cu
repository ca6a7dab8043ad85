#include "qlb_internal.hpp"
#include <algorithm>
using namespace qlb;
#include "qlb_resident_f64.cuh"

namespace
{
    // The fp64 SM-resident kernel (qlb_resident_f64.cuh): reference arithmetic, messages split shared memory / small scratch.
    uint32_t resident64_smem_slots(const qlb_ctx *ctx, const CodeDev &c)
    {
        const size_t avail = (size_t)ctx->smem_optin - kResident64StaticSmem - resident64_small_bytes(c.n, c.m);
        const size_t fit = avail / 8 / 2 * 2;
        return (uint32_t)std::min<size_t>(fit, (size_t)c.slots);
    }
    bool resident64_eligible_impl(const qlb_ctx *ctx, const CodeDev &c)
    {
        return c.slots < 65535 && c.n < 65536 && c.bit_slots16 && c.col_of_slot16 && c.uniform_bit_w >= 2 && c.uniform_bit_w <= 4 &&
               c.max_check_w <= kResidentMaxCW && c.n % 32 == 0 && balanced_block_size(c, kResident64Threads, 0.85, kResident64Threads * 3 / 4, 0.05) > 0 &&
               (size_t)ctx->smem_optin > kResident64StaticSmem + resident64_small_bytes(c.n, c.m) + 8192 &&
               resident64_smem_slots(ctx, c) >= (uint32_t)c.slots / 2;
    }
    template <typename Math, bool kReconcile, int kBW>
    int launch_resident64(qlb_ctx *ctx, DecodeArgs &args)
    {
        auto kern = decode_resident_f64_kernel<Math, kReconcile, kBW, kResident64Threads>;
        int kThreads = balanced_block_size(args.code, kResident64Threads, 0.85, kResident64Threads * 3 / 4, 0.05);
        if (const char *e = std::getenv("QLB_RES64_THREADS")) // experiments only
        {
            const int t = std::atoi(e) / 32 * 32;
            if (t >= 32 && t <= kResident64Threads && 32 * t >= args.code.n && 32 * t >= args.code.m)
                kThreads = t;
        }
        const uint32_t smem_slots = resident64_smem_slots(ctx, args.code);
        const size_t smem = (size_t)smem_slots * 8 + resident64_small_bytes(args.code.n, args.code.m);
        QLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long grid = ctx->sm_count;
        if (grid > args.n_frames)
            grid = args.n_frames;
        const size_t tail = align_up((size_t)(args.code.slots - smem_slots) * 8 + 16, 256);
        QLB_CUDA(ctx->scratch.reserve((size_t)grid * tail));
        args.scratch = static_cast<unsigned char *>(ctx->scratch.p);
        args.scratch_stride = tail;
        if (std::getenv("QLB_DEBUG"))
            std::fprintf(stderr, "[qlb] decode_resident_f64_kernel: %d threads, %u of %d slots in shared memory (%zu B), %zu B tail scratch per CTA, grid=%lld\n",
                         kThreads, smem_slots, args.code.slots, smem, tail, grid);
        QLB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));
        args.queue = ctx->d_counters;
        args.iter_total = ctx->d_counters + 1;
        kern<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(args, smem_slots, args.code.col_of_slot16);
        QLB_CUDA(cudaGetLastError());
        ++ctx->launches;
        return QLB_OK;
    }
    template <typename Math, bool kReconcile>
    int launch_resident64_bw(qlb_ctx *ctx, DecodeArgs &args)
    {
        switch (args.code.uniform_bit_w)
        {
        case 2: return launch_resident64<Math, kReconcile, 2>(ctx, args);
        case 3: return launch_resident64<Math, kReconcile, 3>(ctx, args);
        case 4: return launch_resident64<Math, kReconcile, 4>(ctx, args);
        default: return fail(QLB_ERR_UNSUPPORTED, "fp64 resident kernel: unsupported bit weight");
        }
    }
}
namespace qlb
{
    bool resident_f64_eligible(const qlb_ctx *ctx, const CodeDev &c) { return resident64_eligible_impl(ctx, c); }
    int launch_resident_f64(qlb_ctx *ctx, DecodeArgs &args, bool reconcile, bool fused)
    {
        if (fused)
            return reconcile ? launch_resident64_bw<MathF64Fused, true>(ctx, args) : launch_resident64_bw<MathF64Fused, false>(ctx, args);
        return reconcile ? launch_resident64_bw<MathF64, true>(ctx, args) : launch_resident64_bw<MathF64, false>(ctx, args);
    }
}
