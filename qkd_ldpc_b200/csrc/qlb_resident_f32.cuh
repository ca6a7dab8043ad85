// fp32 "SM-resident" decoder: the specialised hot kernel for codes whose whole message state fits in one SM's shared
// memory and whose bits all have the same weight (the CW=3 family of BASELINE.json, N=10240: 120 KB of messages +
// 60 KB of slot indices). Same schedule and node arithmetic as decode_kernel (qlb_kernels.cuh), re-organised so that
// the inner loops carry (almost) nothing but the node arithmetic:
//
//   * a warp's checks all have the same weight (checks are sorted by weight), so the check update is dispatched once
//     per check to code fully unrolled for that weight -- no per-edge predicates, no local arrays;
//   * the hard decision z of a bit travels in the least-significant mantissa bit of the bit-to-check messages that bit
//     sends (a <= 1 ulp perturbation; fp32 has no bit-exactness contract, its bar is statistical). The check pass
//     XORs the raw words it loads anyway: bit 31 of the XOR is the product's sign, bit 0 is the check's parity. The
//     separate parity phase, its barrier, and the per-edge byte array of decode_kernel disappear;
//   * convergence of iteration t is therefore seen by the check pass of iteration t+1 (one speculative check pass per
//     successful frame, < 1 % of the sweep's work, against ~20 % saved in every iteration);
//   * two block barriers per iteration.
//
// Reference semantics restated (paths relative to the reference repository): src/qkd_ldpc_algorithm.cpp:175-345, 398-447;
// iteration counts, success flags and decoded keys follow the reference's definitions exactly
// (iterations_num = index of the first bit pass whose hard decision satisfies the syndrome, else max_it).
#pragma once
#include "qlb_kernels.cuh"

namespace qlb
{
    __device__ __forceinline__ float ex2_approx(float x)
    {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    __device__ __forceinline__ float lg2_approx(float x)
    {
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    __device__ __forceinline__ float rcp_approx(float x)
    {
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }

    // Check rules for a check of weight exactly W. v[] holds the raw incoming messages and receives the outgoing ones.
    // `xr` is the XOR of the raw message words with the syndrome bit folded into bit 31 (sign of the seeded product).
    struct RuleF32Fast
    {
        // v = exp(-|m|): tanh(|m|/2) = (1-v)/(1+v); leave-one-out products A_k = prod(1-v), B_k = prod(1+v);
        // 2 atanh(A_k/B_k) = ln((B_k+A_k)/(B_k-A_k)). One MUFU.EX2 + MUFU.RCP + MUFU.LG2 per edge.
        template <int W>
        static __device__ __forceinline__ void apply(float (&v)[W], uint32_t xr, uint32_t sbit, float cap)
        {
            float a[W], b[W], preA[W], preB[W];
            float runA = 1.f, runB = 1.f;
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                const float e = ex2_approx(-1.4426950408889634f * fabsf(v[k]));
                a[k] = 1.f - e;
                b[k] = 1.f + e;
                preA[k] = runA;
                preB[k] = runB;
                runA *= a[k];
                runB *= b[k];
            }
            float sufA = 1.f, sufB = 1.f;
#pragma unroll
            for (int k = W - 1; k >= 0; --k)
            {
                const float Ak = preA[k] * sufA, Bk = preB[k] * sufB;
                const uint32_t sk = (xr ^ __float_as_uint(v[k])) & 0x80000000u;
                float mag = 0.6931471805599453f * lg2_approx((Bk + Ak) * rcp_approx(Bk - Ak));
                mag = fminf(mag, cap); // +inf (saturated product) -> threshold; the clamp of :246-249
                v[k] = __uint_as_float(__float_as_uint(mag) | sk);
                sufA *= a[k];
                sufB *= b[k];
            }
        }
    };

    struct RuleF32Accurate
    {
        // libdevice tanhf / atanhf, leave-one-out product by prefix * suffix
        template <int W>
        static __device__ __forceinline__ void apply(float (&v)[W], uint32_t xr, uint32_t sbit, float cap)
        {
            float t[W], pre[W];
            float run = sbit ? -1.f : 1.f;
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                t[k] = tanhf(0.5f * v[k]);
                pre[k] = run;
                run *= t[k];
            }
            float suf = 1.f;
#pragma unroll
            for (int k = W - 1; k >= 0; --k)
            {
                float o = 2.f * atanhf(pre[k] * suf);
                o = fminf(fmaxf(o, -cap), cap); // NaN cannot arise here: |pre*suf| <= 1
                v[k] = o;
                suf *= t[k];
            }
        }
    };

    // One check of weight exactly W at sorted position p. base4[k] (kernel parameter => constant bank, compile-time k)
    // is the byte offset of edge position k's slot row; p4 = 4 * p.
    template <typename Rule, int W>
    __device__ __forceinline__ uint32_t check_fixed(unsigned char *__restrict__ msg_bytes, const DecodeArgs &args, uint32_t p4,
                                                    uint32_t sbit, float cap)
    {
        float v[W];
        uint32_t xr = sbit << 31;
#pragma unroll
        for (int k = 0; k < W; ++k)
        {
            v[k] = *reinterpret_cast<const float *>(msg_bytes + (args.code.base4[k] + p4));
            xr ^= __float_as_uint(v[k]);
        }
        Rule::template apply<W>(v, xr, sbit, cap);
#pragma unroll
        for (int k = 0; k < W; ++k)
            *reinterpret_cast<float *>(msg_bytes + (args.code.base4[k] + p4)) = v[k];
        return (xr ^ sbit) & 1u; // parity of the hard decisions riding in bit 0, against the syndrome bit
    }

    template <typename Rule>
    __device__ __forceinline__ uint32_t check_any(unsigned char *__restrict__ msg_bytes, const DecodeArgs &args, uint32_t p4, int w,
                                                  uint32_t sbit, float cap)
    {
        switch (w)
        {
#define QLB_CASE(W_) case W_: return check_fixed<Rule, W_>(msg_bytes, args, p4, sbit, cap);
            QLB_CASE(1) QLB_CASE(2) QLB_CASE(3) QLB_CASE(4) QLB_CASE(5) QLB_CASE(6) QLB_CASE(7) QLB_CASE(8)
            QLB_CASE(9) QLB_CASE(10) QLB_CASE(11) QLB_CASE(12) QLB_CASE(13) QLB_CASE(14) QLB_CASE(15) QLB_CASE(16)
#undef QLB_CASE
        default: return 0;
        }
    }

    constexpr int kResidentMaxCW = 16;

    __host__ __device__ inline size_t resident_smem_bytes(int n, int m, int e, int bw)
    {
        const size_t wn = align_up((size_t)(n + 31) / 32 * 4, 16), wm = align_up((size_t)(m + 31) / 32 * 4, 16);
        return align_up((size_t)e * 4, 16) + align_up((size_t)bw * n * 2, 16) + 3 * wn + 2 * wm + 256;
    }

    // kBW: the (uniform) bit weight. Requirements checked by the host: e < 65535, max_check_w <= 16, every bit of weight kBW.
    template <typename Rule, bool kReconcile, int kBW, int kThreads>
    __global__ void __launch_bounds__(kThreads, 1) decode_resident_f32_kernel(const DecodeArgs args)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        const CodeDev &code = args.code;
        const int n = code.n, m = code.m, tid = threadIdx.x, lane = tid & 31;
        const int words_n = code.words_n, words_m = code.words_m;
        const size_t wn = align_up((size_t)words_n * 4, 16), wm = align_up((size_t)words_m * 4, 16);

        float *msg = reinterpret_cast<float *>(smem);
        uint16_t *bslot = reinterpret_cast<uint16_t *>(smem + align_up((size_t)code.e * 4, 16));
        unsigned char *tail = reinterpret_cast<unsigned char *>(bslot) + align_up((size_t)kBW * n * 2, 16);
        uint32_t *s_bob = reinterpret_cast<uint32_t *>(tail);
        uint32_t *s_alice = reinterpret_cast<uint32_t *>(tail + wn);
        uint32_t *s_z = reinterpret_cast<uint32_t *>(tail + 2 * wn);
        uint32_t *s_synp = reinterpret_cast<uint32_t *>(tail + 3 * wn);
        uint32_t *s_synn = reinterpret_cast<uint32_t *>(tail + 3 * wn + wm);
        uint32_t *s_cnt = reinterpret_cast<uint32_t *>(tail + 3 * wn + 2 * wm); // [16]
        long long *s_frame = reinterpret_cast<long long *>(s_cnt + kResidentMaxCW);

        for (int i = tid; i < kBW * n; i += kThreads)
            bslot[i] = code.bit_slots16[i];
        if (tid < kResidentMaxCW)
            s_cnt[tid] = code.cnt[tid];
        const int wmax = code.max_check_w;
        const float cap = args.enable_thr ? (float)args.thr : __int_as_float(0x7f800000);
        const int n_round = (n + 31) & ~31, m_round = (m + 31) & ~31;

        for (;;)
        {
            __syncthreads();
            if (tid == 0)
                *s_frame = (long long)atomicAdd(args.queue, 1ULL);
            __syncthreads();
            const long long f = *s_frame;
            if (f >= args.n_frames)
                break;

            // ---- frame set-up ------------------------------------------------------------------------------------
            float lp = 0.f;
            const double *llr_f = nullptr;
            if (kReconcile)
            {
                lp = (float)args.log_prior[f];
                for (int w = tid; w < words_n; w += kThreads)
                {
                    s_bob[w] = args.bob[f * words_n + w];
                    s_alice[w] = args.alice[f * words_n + w];
                }
                for (int w = tid; w < words_m; w += kThreads)
                    s_synn[w] = 0;
            }
            else
            {
                llr_f = args.llr + f * n;
                for (int w = tid; w < words_m; w += kThreads)
                    s_synn[w] = args.syndrome_in[f * words_m + w];
            }
            __syncthreads();

            // messages <- priors (src/qkd_ldpc_algorithm.cpp:182-190); in reconcile mode Alice's bit rides in bit 0 so
            // that the parity of the first pass over the checks is her syndrome (:413-414)
            for (int i = tid; i < n; i += kThreads)
            {
                float prior;
                uint32_t abit = 0;
                if (kReconcile)
                {
                    const uint32_t bb = (s_bob[i >> 5] >> (i & 31)) & 1u;
                    abit = (s_alice[i >> 5] >> (i & 31)) & 1u;
                    prior = bb ? -lp : lp;
                }
                else
                    prior = (float)llr_f[i];
                const float pv = __uint_as_float((__float_as_uint(prior) & ~1u) | abit);
#pragma unroll
                for (int a = 0; a < kBW; ++a)
                    msg[bslot[a * n + i]] = pv;
            }
            __syncthreads();

            for (int p = tid; p < m_round; p += kThreads)
            {
                uint32_t bit = 0;
                if (p < m)
                {
                    if (kReconcile)
                    {
                        for (int k = 0; k < wmax; ++k)
                            if ((uint32_t)p < s_cnt[k])
                                bit ^= __float_as_uint(msg[code.base[k] + p]);
                        bit &= 1u;
                        if (bit)
                        {
                            const uint32_t j = code.check_order[p];
                            atomicOr(&s_synn[j >> 5], 1u << (j & 31));
                        }
                    }
                    else
                    {
                        const uint32_t j = code.check_order[p];
                        bit = (s_synn[j >> 5] >> (j & 31)) & 1u;
                    }
                }
                const uint32_t word = __ballot_sync(0xffffffffu, bit != 0);
                if (lane == 0)
                    s_synp[p >> 5] = word;
            }
            __syncthreads();

            // ---- iterations ----------------------------------------------------------------------------------------
            // Loop index `it` counts completed bit passes. The check pass of round `it` also evaluates the parity of the
            // hard decisions of bit pass `it` (meaningless for it == 0: bit 0 then still holds Alice's / zero bits).
            int it = 0;
            bool success = false;
            for (;;)
            {
                uint32_t bad = 0;
                int w = wmax;
                for (int p = tid; p < m; p += kThreads)
                {
                    const uint32_t sbit = (s_synp[p >> 5] >> lane) & 1u;
                    if ((uint32_t)p >= s_cnt[0])
                    {
                        bad |= sbit; // a check without edges can only be satisfied by a zero syndrome bit
                        continue;
                    }
                    while (w > 1 && (uint32_t)p >= s_cnt[w - 1])
                        --w;
                    bad |= check_any<Rule>(reinterpret_cast<unsigned char *>(msg), args, 4u * (uint32_t)p, w, sbit, cap);
                }
                const int any_bad = __syncthreads_or((int)bad);
                if (it > 0 && !any_bad)
                {
                    success = true; // the decisions of bit pass `it` satisfy the syndrome (:285-298)
                    break;
                }
                if (it == args.max_it)
                    break; // :337-344
                // bit pass: total (:256-258), decision (:259-266), extrinsic + clamp (:300-316)
                for (int i = tid; i < n_round; i += kThreads)
                {
                    bool z = false;
                    if (i < n)
                    {
                        float prior;
                        if (kReconcile)
                            prior = ((s_bob[i >> 5] >> lane) & 1u) ? -lp : lp;
                        else
                            prior = (float)llr_f[i];
                        uint32_t sl[kBW];
                        float c[kBW];
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            sl[a] = bslot[a * n + i];
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            c[a] = msg[sl[a]];
                        float total = prior;
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            total = total + c[a];
                        z = total <= 0.f;
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                        {
                            float v = total - c[a];
                            v = fminf(fmaxf(v, -cap), cap);
                            msg[sl[a]] = __uint_as_float((__float_as_uint(v) & ~1u) | (uint32_t)z);
                        }
                    }
                    const uint32_t word = __ballot_sync(0xffffffffu, z);
                    if (lane == 0)
                        s_z[i >> 5] = word;
                }
                ++it;
                __syncthreads();
            }

            // ---- results -------------------------------------------------------------------------------------------
            int differs = 0;
            for (int w = tid; w < words_n; w += kThreads)
            {
                const uint32_t zw = s_z[w];
                if (args.decoded)
                    args.decoded[f * words_n + w] = zw;
                if (kReconcile)
                    differs |= (zw != s_alice[w]);
            }
            if (kReconcile && args.syndrome_out)
                for (int w = tid; w < words_m; w += kThreads)
                    args.syndrome_out[f * words_m + w] = s_synn[w];
            const int any_diff = __syncthreads_or(differs);
            if (tid == 0)
            {
                uint8_t r = success ? 1 : 0;
                if (kReconcile && !any_diff)
                    r |= 2;
                args.iterations[f] = (uint32_t)it;
                args.result[f] = r;
                atomicAdd(args.iter_total, (unsigned long long)it);
            }
        }
    }
}
