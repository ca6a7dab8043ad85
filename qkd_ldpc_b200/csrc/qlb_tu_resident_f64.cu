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
    int resident64_threads(const CodeDev &c) { return balanced_block_size(c, kResident64Threads, 0.85, kResident64Threads * 3 / 4, 0.05); }

    // The walk over the sorted checks, cut where the shared-memory / tail split of the message array changes sides (see
    // qlb_resident_f64.cuh). Returns false when the table or a thread's 32 per-walk rounds would overflow.
    bool resident64_segments(const CodeDev &c, uint32_t smem_slots, int threads, SegTable64 &t)
    {
        t.n = 0;
        long long rounds = 0;
        auto push = [&](uint32_t lo, uint32_t hi, int w, int tail)
        {
            if (lo >= hi)
                return true;
            if (t.n == kResident64MaxSegs)
                return false;
            t.seg[t.n++] = Seg64{lo, hi, w, tail};
            rounds += (hi - lo + threads - 1) / threads;
            return true;
        };
        for (int w = c.max_check_w; w >= 0; --w)
        {
            const uint32_t lo = (w < c.max_check_w) ? c.cnt[w] : 0u, hi = (w > 0) ? c.cnt[w - 1] : (uint32_t)c.m;
            if (lo >= hi)
                continue;
            if (w == 0 || w > kResident64FastW)
            {
                if (!push(lo, hi, w, -1))
                    return false;
                continue;
            }
            // positions below split(k) keep row k in shared memory; rows are laid out in increasing k, so split(k) decreases with k
            auto split = [&](int k) -> uint32_t
            {
                const long long s = (long long)smem_slots - (long long)c.base[k];
                return (uint32_t)std::min<long long>(hi, std::max<long long>(lo, s));
            };
            const uint32_t s1 = split(w - 1), s2 = w >= 2 ? split(w - 2) : hi; // [lo, s1): all rows shared; [s1, s2): the last row in the tail
            if (!push(lo, s1, w, 0) || !push(s1, s2, w, 1) || !push(s2, hi, w, -1))
                return false;
        }
        return rounds <= 32 && (c.n + threads - 1) / threads <= 32;
    }

    bool resident64_eligible_impl(const qlb_ctx *ctx, const CodeDev &c)
    {
        if (!(c.slots < 65535 && c.n < 65536 && c.bit_slots16 && c.col_of_slot16 && c.uniform_bit_w >= 2 && c.uniform_bit_w <= 4 &&
              c.max_check_w <= kResidentMaxCW && c.n % 32 == 0 && resident64_threads(c) > 0 &&
              (size_t)ctx->smem_optin > kResident64StaticSmem + resident64_small_bytes(c.n, c.m) + 8192 &&
              resident64_smem_slots(ctx, c) >= (uint32_t)c.slots / 2))
            return false;
        SegTable64 t;
        return resident64_segments(c, resident64_smem_slots(ctx, c), resident64_threads(c), t);
    }
    template <typename Math, bool kReconcile, int kBW>
    int launch_resident64(qlb_ctx *ctx, DecodeArgs &args)
    {
        auto kern = decode_resident_f64_kernel<Math, kReconcile, kBW, kResident64Threads>;
        const uint32_t smem_slots = resident64_smem_slots(ctx, args.code);
        int kThreads = resident64_threads(args.code);
        SegTable64 segs;
        if (const int t = args.block_threads / 32 * 32) // qlb_decode_params.block_threads
            if (t >= 32 && t <= kResident64Threads && resident64_segments(args.code, smem_slots, t, segs))
                kThreads = t;
        if (!resident64_segments(args.code, smem_slots, kThreads, segs))
            return fail(QLB_ERR_UNSUPPORTED, "fp64 resident kernel: the check walk does not fit its segment table");
        const size_t smem = (size_t)smem_slots * 8 + resident64_small_bytes(args.code.n, args.code.m);
        QLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long grid = ctx->sm_count;
        if (grid > args.n_frames)
            grid = args.n_frames;
        const size_t tail = align_up((size_t)(args.code.slots - smem_slots) * 8 + 16, 256);
        QLB_CUDA(ctx->scratch.reserve((size_t)grid * tail));
        args.scratch = static_cast<unsigned char *>(ctx->scratch.p);
        args.scratch_stride = tail;
#ifdef QLB_DEBUG_LAUNCH
        std::fprintf(stderr, "[qlb] decode_resident_f64_kernel: %d threads, %u of %d slots in shared memory (%zu B), %zu B tail scratch per CTA, grid=%lld, %d segments\n",
                     kThreads, smem_slots, args.code.slots, smem, tail, grid, segs.n);
#endif
        QLB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));
        args.queue = ctx->d_counters;
        args.iter_total = ctx->d_counters + 1;
        kern<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(args, segs, smem_slots, args.code.col_of_slot16);
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
