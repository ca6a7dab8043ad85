// fp32 frame-interleaved "streaming" decoder: the kernel for codes whose messages do not fit in shared memory
// (BASELINE.json configs[3]: N = 100 000 ... 1 000 000) -- the HBM-bound design point of SURVEY.md 8d.
//
// A frame GROUP of G = 32 * VEC frames (VEC = 4: 128-bit accesses, VEC = 1: large N) is decoded by one persistent CTA.
// Messages are stored slot-major, frame-minor: msg[slot][G]. A WARP works on one node at a time and lane l owns frames
// VEC*l ... VEC*l+VEC-1 of the group, so whatever the Tanner graph looks like, every message access of the warp is one
// fully coalesced row of G floats (512 B for VEC = 4), and the graph indices are warp-uniform (one broadcast load per
// node, amortised over the G frames). Traffic per executed iteration is the algorithmic 16 B per edge and frame
// (read + write in the check pass, read + write in the bit pass) plus < 2 % of indices and packed bits.
// Keys, decisions and syndromes are kept bit-transposed per group ([node][VEC] words, word j bit l = frame VEC*l + j) so a
// lane extracts its frames' bits with one shift; the transposes run once per group with __ballot_sync.
// Node arithmetic, the decision-in-LSB trick and the convergence rule are those of the SM-resident kernel
// (qlb_resident_f32.cuh); a converged frame is frozen (decisions and counters kept) while its group finishes.
#pragma once
#include "qlb_resident_f32.cuh"

namespace qlb
{
    constexpr int kStreamThreads = 512;

    template <int VEC>
    struct VecIO;
    template <>
    struct VecIO<4>
    {
        static __device__ __forceinline__ void load(const float *p, float (&v)[4])
        {
#ifdef QLB_STREAM_CS
            const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
#else
            const float4 t = *reinterpret_cast<const float4 *>(p);
#endif
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
        static __device__ __forceinline__ void store(float *p, const float (&v)[4])
        {
#ifdef QLB_STREAM_CS
            __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
#else
            *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
#endif
        }
    };
    template <>
    struct VecIO<1>
    {
        static __device__ __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = *p; }
        static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
    };

    // Software prefetch into L2: registers bound how many demand loads a warp can keep in flight (W rows of 512 B), which is
    // not enough to cover HBM latency at 16 warps per SM; prefetching the NEXT node's rows costs no registers and turns the
    // demand loads that follow into L2 hits.
    __device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

    // per-group scratch carve-up (bytes); G = 32 * VEC
    struct StreamCarve
    {
        size_t msg, bobT, aliceT, zT, synT, total;
    };
    __host__ __device__ inline StreamCarve stream_carve(int n, int m, int slots, int vec)
    {
        StreamCarve c{};
        const size_t G = 32 * (size_t)vec;
        size_t o = 0;
        c.msg = o; o += align_up((size_t)slots * G * 4, 256);
        c.bobT = o; o += align_up((size_t)n * vec * 4, 256);
        c.aliceT = o; o += align_up((size_t)n * vec * 4, 256);
        c.zT = o; o += align_up((size_t)n * vec * 4, 256);
        c.synT = o; o += align_up((size_t)m * vec * 4, 256);
        c.total = o;
        return c;
    }

    // One check of weight exactly W for the VEC frames of this lane.
    // row_stride: floats between the rows of consecutive slots (G, or B * G when B groups are interleaved slot by slot)
    template <typename Rule, int W, int VEC>
    __device__ __forceinline__ void stream_check(float *__restrict__ msg, const CodeDev &code, uint32_t p, int lane, uint32_t *__restrict__ synT,
                                                 float cap, bool first, uint32_t (&bad)[VEC], size_t row_stride = 32 * VEC)
    {
        float v[VEC][W];
        float *row[W];
#pragma unroll
        for (int k = 0; k < W; ++k)
        {
            row[k] = msg + ((size_t)(code.base[k] + p) * row_stride + VEC * lane);
            float t[VEC];
            VecIO<VEC>::load(row[k], t);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                v[j][k] = t[j];
        }
        uint32_t syn_words[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            syn_words[j] = first ? 0u : synT[(size_t)p * VEC + j];
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
                sb = xr & 1u; // Alice's bits ride in bit 0 during the first pass: their parity IS her syndrome bit
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
            VecIO<VEC>::store(row[k], t);
        }
    }

    // `first`: the reconcile-mode pass that derives Alice's syndrome from the bits riding in the freshly initialised messages
    template <typename Rule, int VEC>
    __device__ __forceinline__ void stream_check_pass(float *__restrict__ msg, const CodeDev &code, const uint32_t *s_seg_w, const uint32_t *s_seg_lo,
                                                      const uint32_t *s_seg_hi, int nseg, uint32_t *__restrict__ synT, float cap, bool first,
                                                      uint32_t ahead, uint32_t (&bad)[VEC])
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
#define QLB_SSEG(W_)                                                                    \
    case W_:                                                                            \
        _Pragma("unroll 1") for (uint32_t p = lo + warp; p < hi; p += nwarps)           \
        {                                                                               \
            if (p + ahead < hi)                                                         \
                _Pragma("unroll") for (int k = 0; k < W_; ++k)                          \
                    prefetch_l2(msg + ((size_t)(code.base[k] + p + ahead) * (32 * VEC) + VEC * lane)); \
            stream_check<Rule, W_, VEC>(msg, code, p, lane, synT, cap, first, bad);     \
        }                                                                               \
        break;
                QLB_SSEG(1) QLB_SSEG(2) QLB_SSEG(3) QLB_SSEG(4) QLB_SSEG(5) QLB_SSEG(6) QLB_SSEG(7) QLB_SSEG(8)
                QLB_SSEG(9) QLB_SSEG(10) QLB_SSEG(11) QLB_SSEG(12) QLB_SSEG(13) QLB_SSEG(14) QLB_SSEG(15) QLB_SSEG(16)
#undef QLB_SSEG
            default: // checks without edges: satisfied only by a zero syndrome bit
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
    }

    // 32 x 32 bit transposes between frame-major packed words and the per-group node-major layout.
    // in:  word `wd` of frames f0 + VEC*l + j (lane l)      out: T[(32*wd + b) * VEC + j] = word whose bit l is bit b of that frame's word
    template <int VEC>
    __device__ __forceinline__ void transpose_in(const uint32_t *__restrict__ frames, long long f0, long long n_frames, int words, int n, uint32_t *__restrict__ T)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int item = warp; item < words * VEC; item += nwarps)
        {
            const int wd = item / VEC, j = item % VEC;
            const long long f = f0 + (long long)VEC * lane + j;
            const uint32_t x = f < n_frames ? frames[f * words + wd] : 0u;
            uint32_t mine = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b)
            {
                const uint32_t col = __ballot_sync(0xffffffffu, (x >> b) & 1u);
                if (lane == b)
                    mine = col;
            }
            const int bit = 32 * wd + lane;
            if (bit < n)
                T[(size_t)bit * VEC + j] = mine;
        }
    }
    // fmap_g (optional): frame index of each of the group's 32 * VEC columns (0xFFFFFFFF = none) instead of f0 + column;
    // keep[j] (with fmap_g): only the columns whose bit is set in keep[j] are written
    template <int VEC>
    __device__ __forceinline__ void transpose_out(const uint32_t *__restrict__ T, long long f0, long long n_frames, int words, int n, uint32_t *__restrict__ frames,
                                                  const uint32_t *__restrict__ fmap_g = nullptr, const uint32_t *keep = nullptr)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int item = warp; item < words * VEC; item += nwarps)
        {
            const int wd = item / VEC, j = item % VEC;
            const int bit = 32 * wd + lane;
            const uint32_t x = bit < n ? T[(size_t)bit * VEC + j] : 0u;
            uint32_t mine = 0;
#pragma unroll
            for (int l = 0; l < 32; ++l)
            {
                const uint32_t row = __ballot_sync(0xffffffffu, (x >> l) & 1u);
                if (lane == l)
                    mine = row;
            }
            long long f = f0 + (long long)VEC * lane + j;
            if (fmap_g)
            {
                const uint32_t fm = fmap_g[VEC * lane + j];
                f = (fm == 0xFFFFFFFFu || (keep && !((keep[j] >> lane) & 1u))) ? n_frames : (long long)fm;
            }
            if (f < n_frames)
                frames[f * words + wd] = mine;
        }
    }

}
#include "qlb_stream_tma.cuh"
namespace qlb
{
    // Requirements (host-checked): max_check_w <= 16 (<= 8 for kTma), uniform bit weight kBW. Grid: persistent, one group per
    // CTA at a time. kTma: the passes run through per-warp TMA rings in dynamic shared memory (`stages` per warp).
    template <typename Rule, bool kReconcile, int kBW, int VEC, bool kTma>
    __global__ void __launch_bounds__(kStreamThreads, 1) decode_stream_f32_kernel(const DecodeArgs args, unsigned char *__restrict__ group_scratch,
                                                                                  size_t group_stride, long long n_groups, int prefetch_nodes, int stages)
    {
        constexpr int G = 32 * VEC;
        constexpr int kWarps = kStreamThreads / 32;
        const int ahead = prefetch_nodes * kWarps; // this warp's node `prefetch_nodes` steps ahead
        extern __shared__ __align__(128) unsigned char ring_smem[];
        WarpPipe pp{};
        if (kTma)
        {
            const uint32_t rows = (uint32_t)max(args.code.max_check_w, kBW);
            pp.stage_bytes = rows * 4u * G;
            pp.S = stages;
            pp.stages = ring_smem + (size_t)(threadIdx.x >> 5) * stages * pp.stage_bytes;
            pp.bars = reinterpret_cast<uint64_t *>(ring_smem + (size_t)kWarps * stages * pp.stage_bytes) + (size_t)(threadIdx.x >> 5) * stages;
            if ((threadIdx.x & 31) == 0)
                for (int s = 0; s < stages; ++s)
                    mbar_init(&pp.bars[s], 1);
        }
        __shared__ uint32_t s_seg_w[kResidentMaxCW + 1], s_seg_lo[kResidentMaxCW + 1], s_seg_hi[kResidentMaxCW + 1];
        __shared__ uint32_t s_flags[kWarps][32];
        __shared__ int s_nseg;
        __shared__ long long s_group;

        const CodeDev &code = args.code;
        const int n = code.n, m = code.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int words_n = code.words_n, words_m = code.words_m;
        const StreamCarve cv = stream_carve(n, m, code.slots, VEC);
        unsigned char *scratch = group_scratch + (size_t)blockIdx.x * group_stride;
        float *msg = reinterpret_cast<float *>(scratch + cv.msg);
        uint32_t *bobT = reinterpret_cast<uint32_t *>(scratch + cv.bobT);
        uint32_t *aliceT = reinterpret_cast<uint32_t *>(scratch + cv.aliceT);
        uint32_t *zT = reinterpret_cast<uint32_t *>(scratch + cv.zT);
        uint32_t *synT = reinterpret_cast<uint32_t *>(scratch + cv.synT);

        if (tid == 0)
        {
            int ns = 0;
            for (int w = code.max_check_w; w >= 0; --w)
            {
                const uint32_t lo = (w < code.max_check_w) ? code.cnt[w] : 0u, hi = (w > 0) ? code.cnt[w - 1] : (uint32_t)m;
                if (lo < hi)
                {
                    s_seg_w[ns] = (uint32_t)w;
                    s_seg_lo[ns] = lo;
                    s_seg_hi[ns] = hi;
                    ++ns;
                }
            }
            s_nseg = ns;
        }
        __syncthreads();
        const int nseg = s_nseg;
        const float unit = Rule::kUnit;
        const float cap = args.cap_f32 * unit;
        const bool clamp_b2c = !(Rule::kUnit != 1.f && cap >= 25.f);

        for (;;)
        {
            __syncthreads();
            if (tid == 0)
                s_group = (long long)atomicAdd(args.queue, 1ULL);
            __syncthreads();
            const long long grp = s_group;
            if (grp >= n_groups)
                break;
            const long long f0 = grp * G;

            // ---- group set-up: transposed keys, priors, messages ----------------------------------------------------------
            float lp[VEC];
            uint32_t active = 0; // bit j: frame VEC*lane + j still decoding
            uint32_t iters[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
            {
                const long long f = f0 + (long long)VEC * lane + j;
                lp[j] = 0.f;
                iters[j] = (uint32_t)args.max_it;
                if (f < args.n_frames)
                {
                    active |= 1u << j;
                    if (kReconcile)
                        lp[j] = unit * (float)args.log_prior[f];
                }
            }
            uint32_t success = 0;
            if (kReconcile)
            {
                transpose_in<VEC>(args.bob, f0, args.n_frames, words_n, n, bobT);
                transpose_in<VEC>(args.alice, f0, args.n_frames, words_n, n, aliceT);
            }
            else
            {
                // target syndromes arrive in natural check order: transpose into the (unused) Alice area, then sort
                transpose_in<VEC>(args.syndrome_in, f0, args.n_frames, words_m, m, aliceT);
                __syncthreads();
                for (int p = tid; p < m; p += kStreamThreads)
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        synT[(size_t)p * VEC + j] = aliceT[(size_t)code.check_order[p] * VEC + j];
            }
            __syncthreads();

            for (int i = warp; i < n; i += kWarps)
            {
                float pv[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                {
                    float prior;
                    uint32_t abit = 0;
                    if (kReconcile)
                    {
                        const uint32_t bb = (bobT[(size_t)i * VEC + j] >> lane) & 1u;
                        abit = (aliceT[(size_t)i * VEC + j] >> lane) & 1u;
                        prior = __uint_as_float(__float_as_uint(lp[j]) ^ (bb << 31));
                    }
                    else
                    {
                        const long long f = f0 + (long long)VEC * lane + j;
                        prior = f < args.n_frames ? __fmul_rn(unit, (float)args.llr[f * n + i]) : 0.f;
                    }
                    pv[j] = __uint_as_float((__float_as_uint(prior) & ~1u) | abit);
                }
#pragma unroll
                for (int a = 0; a < kBW; ++a)
                    VecIO<VEC>::store(msg + ((size_t)code.bit_slots32[(size_t)a * n + i] * G + VEC * lane), pv);
                if (lane == 0)
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        zT[(size_t)i * VEC + j] = 0;
            }
            __syncthreads();

            // ---- iterations (`it` = completed bit passes; the check pass of round it > 0 sees the parity of bit pass it) ----
            int it = 0;
            for (;;)
            {
                uint32_t bad[VEC];
                if constexpr (kTma)
                    tma_check_pass<Rule, VEC>(msg, code, s_seg_w, s_seg_lo, s_seg_hi, nseg, synT, cap, kReconcile && it == 0, bad, pp);
                else
                    stream_check_pass<Rule, VEC>(msg, code, s_seg_w, s_seg_lo, s_seg_hi, nseg, synT, cap, kReconcile && it == 0, (uint32_t)ahead, bad);
                uint32_t nib = 0;
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    nib |= (bad[j] & 1u) << j;
                s_flags[warp][lane] = nib;
                __syncthreads();
                uint32_t frame_bad = 0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w)
                    frame_bad |= s_flags[w][lane];
                if (it > 0)
                {
                    const uint32_t done = active & ~frame_bad; // decisions of bit pass `it` satisfy the syndrome
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        if ((done >> j) & 1u)
                            iters[j] = (uint32_t)it;
                    success |= done;
                    active &= ~done;
                }
                const int any_active = __syncthreads_or((int)active); // also orders the s_flags reads before the next writes
                if (!any_active || it == args.max_it)
                    break;

                // bit pass
                uint32_t act_word[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    act_word[j] = __ballot_sync(0xffffffffu, (active >> j) & 1u);
                if constexpr (kTma)
                    tma_bit_pass<kReconcile, kBW, VEC>(msg, args, bobT, zT, lp, act_word, f0, unit, cap, clamp_b2c, pp);
                else
                for (int i = warp; i < n; i += kWarps)
                {
                    if (i + ahead < n)
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            prefetch_l2(msg + ((size_t)code.bit_slots32[(size_t)a * n + i + ahead] * G + VEC * lane));
                    float *row[kBW];
                    float c[kBW][VEC];
#pragma unroll
                    for (int a = 0; a < kBW; ++a)
                    {
                        row[a] = msg + ((size_t)code.bit_slots32[(size_t)a * n + i] * G + VEC * lane);
                        VecIO<VEC>::load(row[a], c[a]);
                    }
                    float total[VEC];
                    uint32_t zbits = 0;
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                    {
                        float prior;
                        if (kReconcile)
                            prior = __uint_as_float(__float_as_uint(lp[j]) ^ (((bobT[(size_t)i * VEC + j] >> lane) & 1u) << 31));
                        else
                        {
                            const long long f = f0 + (long long)VEC * lane + j;
                            prior = f < args.n_frames ? __fmul_rn(unit, (float)args.llr[f * n + i]) : 0.f;
                        }
                        float t = prior;
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
                        VecIO<VEC>::store(row[a], o);
                    }
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                    {
                        const uint32_t word = __ballot_sync(0xffffffffu, (zbits >> j) & 1u);
                        if (lane == 0)
                        {
                            const size_t at = (size_t)i * VEC + j;
                            zT[at] = (zT[at] & ~act_word[j]) | (word & act_word[j]); // converged frames keep their decision
                        }
                    }
                }
                ++it;
                __syncthreads();
            }

            // ---- results ---------------------------------------------------------------------------------------------------
            __syncthreads();
            uint32_t differs = 0;
            if (kReconcile)
            {
                uint32_t d[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    d[j] = 0;
                for (int i = tid; i < n; i += kStreamThreads)
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        d[j] |= zT[(size_t)i * VEC + j] ^ aliceT[(size_t)i * VEC + j];
                // d[j] bit l set <=> frame VEC*l + j differs somewhere among this thread's bits: OR over the block per (j, l)
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                {
                    uint32_t x = d[j];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        x |= __shfl_xor_sync(0xffffffffu, x, o);
                    if (lane == 0)
                        s_flags[warp][j] = x;
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                {
                    uint32_t x = 0;
                    for (int w = 0; w < kWarps; ++w)
                        x |= s_flags[w][j];
                    differs |= ((x >> lane) & 1u) << j;
                }
            }
            if (warp == 0)
            {
                unsigned long long it_sum = 0;
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                {
                    const long long f = f0 + (long long)VEC * lane + j;
                    if (f < args.n_frames)
                    {
                        uint8_t r = (success >> j) & 1u ? 1 : 0;
                        if (kReconcile && !((differs >> j) & 1u))
                            r |= 2;
                        args.iterations[f] = iters[j];
                        args.result[f] = r;
                        it_sum += iters[j];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    it_sum += __shfl_xor_sync(0xffffffffu, it_sum, o);
                if (lane == 0)
                    atomicAdd(args.iter_total, it_sum);
            }
            if (args.decoded)
                transpose_out<VEC>(zT, f0, args.n_frames, words_n, n, args.decoded);
            if (kReconcile && args.syndrome_out)
                // synT is indexed by sorted check position: un-sort on the way out through a per-frame bit scatter
                for (int p = warp; p < m; p += kWarps)
                {
                    const uint32_t jn = code.check_order[p];
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                    {
                        const long long f = f0 + (long long)VEC * lane + j;
                        if (f < args.n_frames && ((synT[(size_t)p * VEC + j] >> lane) & 1u))
                            atomicOr(&args.syndrome_out[f * words_m + (jn >> 5)], 1u << (jn & 31));
                    }
                }
        }
    }
}
