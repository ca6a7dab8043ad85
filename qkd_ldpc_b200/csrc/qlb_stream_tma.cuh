// TMA-pipelined passes of the fp32 streaming decoder (qlb_stream_f32.cuh).
//
// In the plain streaming kernel a warp can only keep as many message rows in flight as it has registers for (W rows of
// 512 B), which at 16 warps per SM does not cover HBM latency (ncu: 71 % long-scoreboard stalls, 0.53-0.58 of the measured
// copy bandwidth). Here every warp owns a private ring of S stages in shared memory: lane 0 streams the rows of the
// warp's next nodes into it with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), the warp computes a node out of
// shared memory (128-bit, conflict-free), writes the result back into the stage and lane 0 sends it home with TMA bulk
// stores (cp.async.bulk.global.shared::cta, bulk async-groups). Loads run S-2 nodes ahead of the arithmetic, stores drain
// behind it, no registers are tied up, and no inter-warp synchronisation exists inside a pass.
//
// STATUS (measured on B200, round 1): correct (tests force it with QLB_STREAM_TMA=1) but SLOWER than the plain kernel with L2
// software prefetch -- 0.49 vs 0.57-0.58 of the measured HBM copy bandwidth at N = 100 000: one bulk copy per 512-byte row
// is too fine a grain for the TMA unit. Kept opt-in; the next step is one bulk copy per run of consecutive checks
// (rows of consecutive checks are contiguous: 16 checks = 8 KB per edge position).
#pragma once
// (included by qlb_stream_f32.cuh after its helpers; not a stand-alone header)

namespace qlb
{
    struct WarpPipe
    {
        unsigned char *stages; // this warp's S stages
        uint64_t *bars;        // this warp's S "full" mbarriers
        uint32_t stage_bytes;
        int S;
        int use_stage, use_phase, issue_stage; // ring positions (warp-uniform)
        __device__ __forceinline__ void advance_use()
        {
            if (++use_stage == S)
            {
                use_stage = 0;
                use_phase ^= 1;
            }
        }
        __device__ __forceinline__ void advance_issue()
        {
            if (++issue_stage == S)
                issue_stage = 0;
        }
    };

    // All checks of weight exactly W assigned to this warp in [lo, hi): p = lo + warp, lo + warp + nwarps, ...
    template <typename Rule, int W, int VEC>
    __device__ __forceinline__ void tma_check_segment(float *__restrict__ msg, const CodeDev &code, uint32_t lo, uint32_t hi, int lane, int warp,
                                                      int nwarps, uint32_t *__restrict__ synT, float cap, bool first, uint32_t (&bad)[VEC], WarpPipe &pp)
    {
        constexpr int G = 32 * VEC;
        constexpr uint32_t kRowBytes = 4 * G;
        const uint32_t first_p = lo + warp;
        if (first_p >= hi)
            return;
        const uint32_t cnt = (hi - first_p + nwarps - 1) / nwarps;
        const uint32_t ahead = (uint32_t)pp.S - 2;
        auto issue = [&](uint32_t idx) // lane 0
        {
            const uint32_t p = first_p + idx * nwarps;
            unsigned char *dst = pp.stages + (size_t)pp.issue_stage * pp.stage_bytes;
            uint64_t *bar = &pp.bars[pp.issue_stage];
            mbar_expect_tx(bar, W * kRowBytes);
#pragma unroll
            for (int k = 0; k < W; ++k)
                tma_bulk_g2s(dst + k * kRowBytes, msg + (size_t)(code.base[k] + p) * G, kRowBytes, bar);
        };
        for (uint32_t idx = 0; idx < ahead && idx < cnt; ++idx)
        {
            if (lane == 0)
                issue(idx);
            pp.advance_issue();
        }
#pragma unroll 1
        for (uint32_t idx = 0; idx < cnt; ++idx)
        {
            if (idx + ahead < cnt)
            {
                if (lane == 0)
                {
                    bulk_wait_read<1>(); // the stores that last used the target stage (two nodes ago) have left shared memory
                    issue(idx + ahead);
                }
                pp.advance_issue();
            }
            const uint32_t p = first_p + idx * nwarps;
            float *st = reinterpret_cast<float *>(pp.stages + (size_t)pp.use_stage * pp.stage_bytes);
            uint32_t syn_words[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                syn_words[j] = first ? 0u : synT[(size_t)p * VEC + j];
            mbar_wait(&pp.bars[pp.use_stage], (uint32_t)pp.use_phase);
            float v[VEC][W];
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                float t[VEC];
                VecIO<VEC>::load(st + k * G + VEC * lane, t);
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    v[j][k] = t[j];
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j)
            {
                uint32_t xr = 0;
#pragma unroll
                for (int k = 0; k < W; ++k)
                    xr ^= __float_as_uint(v[j][k]);
                uint32_t sb;
                if (first)
                {
                    sb = xr & 1u;
                    const uint32_t word = __ballot_sync(0xffffffffu, sb != 0);
                    if (lane == 0)
                        synT[(size_t)p * VEC + j] = word;
                }
                else
                {
                    sb = (syn_words[j] >> lane) & 1u;
                    bad[j] |= (xr ^ sb) & 1u;
                }
                xr ^= sb << 31;
                Rule::template apply<W>(v[j], xr, cap);
            }
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                float t[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    t[j] = v[j][k];
                VecIO<VEC>::store(st + k * G + VEC * lane, t);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
            {
#pragma unroll
                for (int k = 0; k < W; ++k)
                    tma_bulk_s2g(msg + (size_t)(code.base[k] + p) * G, st + k * G, kRowBytes);
                bulk_commit();
            }
            pp.advance_use();
        }
    }

    template <typename Rule, int VEC>
    __device__ __forceinline__ void tma_check_pass(float *__restrict__ msg, const CodeDev &code, const uint32_t *s_seg_w, const uint32_t *s_seg_lo,
                                                   const uint32_t *s_seg_hi, int nseg, uint32_t *__restrict__ synT, float cap, bool first,
                                                   uint32_t (&bad)[VEC], WarpPipe &pp)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            bad[j] = 0;
#pragma unroll 1
        for (int sg = 0; sg < nseg; ++sg)
        {
            const uint32_t lo = s_seg_lo[sg], hi = s_seg_hi[sg];
            switch (s_seg_w[sg])
            {
#define QLB_TSEG(W_) case W_: tma_check_segment<Rule, W_, VEC>(msg, code, lo, hi, lane, warp, nwarps, synT, cap, first, bad, pp); break;
                QLB_TSEG(1) QLB_TSEG(2) QLB_TSEG(3) QLB_TSEG(4) QLB_TSEG(5) QLB_TSEG(6) QLB_TSEG(7) QLB_TSEG(8)
#undef QLB_TSEG
            default: // weight 0 (weights above 8 are routed to the plain streaming kernel by the host)
                for (uint32_t p = lo + warp; p < hi; p += nwarps)
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                    {
                        if (first)
                        {
                            if (lane == 0)
                                synT[(size_t)p * VEC + j] = 0;
                        }
                        else
                            bad[j] |= (synT[(size_t)p * VEC + j] >> lane) & 1u;
                    }
                break;
            }
        }
        if (lane == 0)
            bulk_wait_all(); // results are in global memory before the block barrier that starts the bit pass
        __syncwarp();
    }

    // Bit pass over this warp's bits i = warp, warp + nwarps, ...
    template <bool kReconcile, int kBW, int VEC>
    __device__ __forceinline__ void tma_bit_pass(float *__restrict__ msg, const DecodeArgs &args, const uint32_t *__restrict__ bobT, uint32_t *__restrict__ zT,
                                                 const float (&lp)[VEC], const uint32_t (&act_word)[VEC], long long f0, float unit, float cap,
                                                 bool clamp_b2c, WarpPipe &pp)
    {
        constexpr int G = 32 * VEC;
        constexpr uint32_t kRowBytes = 4 * G;
        const CodeDev &code = args.code;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, n = code.n;
        if (warp >= n)
            return;
        const uint32_t cnt = (uint32_t)(n - warp + nwarps - 1) / nwarps;
        const uint32_t ahead = (uint32_t)pp.S - 2;
        auto issue = [&](uint32_t idx) // lane 0
        {
            const int i = warp + (int)idx * nwarps;
            unsigned char *dst = pp.stages + (size_t)pp.issue_stage * pp.stage_bytes;
            uint64_t *bar = &pp.bars[pp.issue_stage];
            mbar_expect_tx(bar, kBW * kRowBytes);
#pragma unroll
            for (int a = 0; a < kBW; ++a)
                tma_bulk_g2s(dst + a * kRowBytes, msg + (size_t)code.bit_slots32[(size_t)a * n + i] * G, kRowBytes, bar);
        };
        for (uint32_t idx = 0; idx < ahead && idx < cnt; ++idx)
        {
            if (lane == 0)
                issue(idx);
            pp.advance_issue();
        }
#pragma unroll 1
        for (uint32_t idx = 0; idx < cnt; ++idx)
        {
            if (idx + ahead < cnt)
            {
                if (lane == 0)
                {
                    bulk_wait_read<1>();
                    issue(idx + ahead);
                }
                pp.advance_issue();
            }
            const int i = warp + (int)idx * nwarps;
            float *st = reinterpret_cast<float *>(pp.stages + (size_t)pp.use_stage * pp.stage_bytes);
            float prior[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
            {
                if (kReconcile)
                    prior[j] = __uint_as_float(__float_as_uint(lp[j]) ^ (((bobT[(size_t)i * VEC + j] >> lane) & 1u) << 31));
                else
                {
                    const long long f = f0 + (long long)VEC * lane + j;
                    prior[j] = f < args.n_frames ? __fmul_rn(unit, (float)args.llr[f * n + i]) : 0.f;
                }
            }
            mbar_wait(&pp.bars[pp.use_stage], (uint32_t)pp.use_phase);
            float c[kBW][VEC];
#pragma unroll
            for (int a = 0; a < kBW; ++a)
                VecIO<VEC>::load(st + a * G + VEC * lane, c[a]);
            float total[VEC];
            uint32_t zbits = 0;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
            {
                float t = prior[j];
#pragma unroll
                for (int a = 0; a < kBW; ++a)
                    t = t + c[a][j];
                total[j] = t;
                zbits |= (uint32_t)(t <= 0.f) << j;
            }
#pragma unroll
            for (int a = 0; a < kBW; ++a)
            {
                float o[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                {
                    float v = total[j] - c[a][j];
                    if (clamp_b2c)
                        v = fminf(fmaxf(v, -cap), cap);
                    o[j] = __uint_as_float((__float_as_uint(v) & ~1u) | ((zbits >> j) & 1u));
                }
                VecIO<VEC>::store(st + a * G + VEC * lane, o);
            }
            fence_proxy_async();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < VEC; ++j)
            {
                const uint32_t word = __ballot_sync(0xffffffffu, (zbits >> j) & 1u);
                if (lane == 0)
                {
                    const size_t at = (size_t)i * VEC + j;
                    zT[at] = (zT[at] & ~act_word[j]) | (word & act_word[j]);
                }
            }
            if (lane == 0)
            {
#pragma unroll
                for (int a = 0; a < kBW; ++a)
                    tma_bulk_s2g(msg + (size_t)code.bit_slots32[(size_t)a * n + i] * G, st + a * G, kRowBytes);
                bulk_commit();
            }
            pp.advance_use();
        }
        if (lane == 0)
            bulk_wait_all();
        __syncwarp();
    }
}
