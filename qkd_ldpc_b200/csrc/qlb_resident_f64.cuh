// fp64 SM-resident decoder: the reference's arithmetic and operation order (MathF64, qlb_kernels.cuh) with the frame's
// messages kept on the SM instead of in an L2-resident scratch.
//
// 30 720 doubles (240 KB) do not fit in the 227 KB of shared memory of one SM, so the message array is SPLIT: the first
// `smem_slots` physical slots (~92 % on the N=10240 code: all of edge positions 0..4 and the head of position 5) live in
// shared memory, the tail in a small per-CTA global scratch (coalesced in the check pass, L2 hits in the bit pass). To
// make room, the slot table and the bit index of every slot are streamed from global memory (read-only, coalesced, shared
// by all CTAs through L2), and the hard decisions are kept bit-packed (the bit pass ballots them into the key words); the
// parity of a check is formed from those packed decisions through the slot->bit table at the start of the NEXT check pass
// (bit-exact fp64 messages leave no spare mantissa bit to carry the decision as the fp32 kernel does). Iteration counts,
// flags and keys follow the reference's definitions exactly (src/qkd_ldpc_algorithm.cpp:175-345, 398-447); the generic
// decode_kernel<MathF64> stays as the fallback for codes this kernel does not take.
#pragma once
#include "qlb_resident_f32.cuh"

namespace qlb
{
    constexpr int kResident64Threads = 768; // launch bound; the host picks the block size that balances the node walks
    constexpr size_t kResident64StaticSmem = 2 * kResident64Threads * 4 + 1024;

    struct Split64
    {
        double *smem;
        double *gmem;
        uint32_t smem_slots;
        __device__ __forceinline__ double ld(uint32_t slot) const { return slot < smem_slots ? smem[slot] : gmem[slot - smem_slots]; }
        __device__ __forceinline__ void st(uint32_t slot, double v) const
        {
            if (slot < smem_slots)
                smem[slot] = v;
            else
                gmem[slot - smem_slots] = v;
        }
    };

    __host__ __device__ inline size_t resident64_small_bytes(int n, int m)
    {
        const size_t wn = align_up((size_t)(n + 31) / 32 * 4, 16), wm = align_up((size_t)(m + 31) / 32 * 4, 16);
        return 3 * wn + wm;
    }

    // All checks of weight exactly W in [lo, hi). zsrc: packed bits whose parity per check is wanted (the last hard decision,
    // or Alice's key during the first pass, which yields her syndrome: src/qkd_ldpc_algorithm.cpp:413-414).
    struct Base64FromParams
    {
        const DecodeArgs &args;
        __device__ __forceinline__ uint32_t operator()(int k) const { return args.code.base[k]; }
    };
    struct Base64FromSmem
    {
        const uint32_t *base;
        __device__ __forceinline__ uint32_t operator()(int k) const { return base[k]; }
    };

    template <typename Math, int W, typename Base>
    __device__ __forceinline__ uint32_t check_segment64(int kThreads, const Split64 &msg, const Base base, const uint16_t *__restrict__ col_of_slot,
                                                        const uint32_t *__restrict__ zsrc, uint32_t lo, uint32_t hi, uint32_t &my_syn, int &rbit,
                                                        bool first, bool en, double thr)
    {
        uint32_t bad = 0;
#pragma unroll 1
        for (uint32_t p = lo + threadIdx.x; p < hi; p += kThreads, ++rbit)
        {
            double v[W];
            uint32_t par = 0;
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                const uint32_t slot = base(k) + p;
                const uint32_t col = col_of_slot[slot];
                par ^= zsrc[col >> 5] >> (col & 31);
                v[k] = msg.ld(slot);
            }
            par &= 1u;
            uint32_t sb;
            if (first)
            {
                sb = par;
                my_syn |= sb << rbit;
            }
            else
            {
                sb = (my_syn >> rbit) & 1u;
                bad |= par ^ sb; // calculate_syndrome + arrays_equal of :277-298, one iteration late
            }
            Math::template check<W>(v, W, sb != 0, en, thr);
#pragma unroll
            for (int k = 0; k < W; ++k)
                msg.st(base(k) + p, v[k]);
        }
        return bad;
    }

    // Weights 9..16 (the R >= 0.7 codes of the CW = 3 family) are kept out of line: their register appetite (2 x W doubles live) must
    // not leak into the hot instantiations; some spilling inside them is still far cheaper than the generic kernel's L2 round trips.
    // Returns {bit 0: parity failure, bits 1..: advanced round counter} and the (first pass) accumulated syndrome bits.
    template <typename Math, int W>
    __device__ __noinline__ uint2 check_segment64_wide(int kThreads, const Split64 msg, const uint32_t *s_base, const uint16_t *col_of_slot,
                                                       const uint32_t *zsrc, uint32_t lo, uint32_t hi, uint32_t my_syn, int rbit, bool first, bool en,
                                                       double thr)
    {
        const uint32_t bad = check_segment64<Math, W>(kThreads, msg, Base64FromSmem{s_base}, col_of_slot, zsrc, lo, hi, my_syn, rbit, first, en, thr);
        return make_uint2((bad & 1u) | ((uint32_t)rbit << 1), my_syn);
    }

    // kBW: uniform bit weight. Host-checked: slots < 65535, n < 65536, max_check_w <= 16, n % 32 == 0, n, m <= 32 * kThreads.
    template <typename Math, bool kReconcile, int kBW, int kMaxThreads>
    __global__ void __launch_bounds__(kMaxThreads, 1) decode_resident_f64_kernel(const DecodeArgs args, uint32_t smem_slots,
                                                                                 const uint16_t *__restrict__ col_of_slot)
    {
        const int kThreads = blockDim.x; // multiple of 32, <= kMaxThreads
        extern __shared__ __align__(16) unsigned char smem[];
        __shared__ uint32_t s_seg_w[kResidentMaxCW + 1], s_seg_lo[kResidentMaxCW + 1], s_seg_hi[kResidentMaxCW + 1];
        __shared__ uint32_t s_park_bob[kMaxThreads], s_park_syn[kMaxThreads];
        __shared__ uint32_t s_base[kResidentMaxCW];
        __shared__ int s_nseg;
        __shared__ long long s_frame;

        const CodeDev &code = args.code;
        const int n = code.n, m = code.m, tid = threadIdx.x, lane = tid & 31;
        const int words_n = code.words_n, words_m = code.words_m;
        const size_t wn = align_up((size_t)words_n * 4, 16);
        if (tid < kResidentMaxCW)
            s_base[tid] = code.base[tid];

        Split64 msg;
        msg.smem = reinterpret_cast<double *>(smem);
        msg.smem_slots = smem_slots;
        msg.gmem = reinterpret_cast<double *>(args.scratch + (size_t)blockIdx.x * args.scratch_stride);
        unsigned char *tail = smem + (size_t)smem_slots * 8;
        uint32_t *s_bob = reinterpret_cast<uint32_t *>(tail);
        uint32_t *s_alice = reinterpret_cast<uint32_t *>(tail + wn);
        uint32_t *s_z = reinterpret_cast<uint32_t *>(tail + 2 * wn);
        uint32_t *s_synn = reinterpret_cast<uint32_t *>(tail + 3 * wn);
        const uint16_t *bslot = code.bit_slots16;

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
        const bool en = args.enable_thr != 0;
        const double thr = args.thr;

        for (;;)
        {
            __syncthreads();
            if (tid == 0)
                s_frame = (long long)atomicAdd(args.queue, 1ULL);
            __syncthreads();
            const long long f = s_frame;
            if (f >= args.n_frames)
                break;

            double lp = 0.;
            const double *llr_f = nullptr;
            if (kReconcile)
            {
                lp = args.log_prior[f];
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

            // messages <- priors, unclamped (:182-190); Bob's bit of the r-th bit this thread visits -> bit r
            {
                uint32_t my_bob = 0;
                int r = 0;
                for (int i = tid; i < n; i += kThreads, ++r)
                {
                    double prior;
                    if (kReconcile)
                    {
                        const uint32_t bb = (s_bob[i >> 5] >> lane) & 1u;
                        my_bob |= bb << r;
                        prior = bb ? -lp : lp; // :401-405
                    }
                    else
                        prior = llr_f[i];
#pragma unroll
                    for (int a = 0; a < kBW; ++a)
                        msg.st(bslot[a * n + i], prior);
                }
                s_park_bob[tid] = my_bob;
            }
            if (!kReconcile)
            {
                // target syndrome bits of this thread's checks, in its walk order
                uint32_t my_syn = 0;
                int r = 0;
                for (int sg = 0; sg < nseg; ++sg)
                    for (uint32_t p = s_seg_lo[sg] + tid; p < s_seg_hi[sg]; p += kThreads, ++r)
                    {
                        const uint32_t j = code.check_order[p];
                        my_syn |= ((s_synn[j >> 5] >> (j & 31)) & 1u) << r;
                    }
                s_park_syn[tid] = my_syn;
            }
            __syncthreads();

            // `it` counts completed bit passes; the check pass of round it > 0 first evaluates the parity of bit pass `it`
            // As in the fp32 kernel: when the previous check pass saw only a handful of unsatisfied checks, look at the parity on its
            // own before paying a whole check pass (85 % of an fp64 iteration) to find out that the frame has converged.
            int it = 0;
            bool success = false, quiet = false;
            for (;;)
            {
                if (quiet && it > 0)
                {
                    uint32_t wrong = 0;
                    const uint32_t my_syn = s_park_syn[tid];
                    int r = 0;
                    for (int sg = 0; sg < nseg; ++sg)
                    {
                        const int w = (int)s_seg_w[sg];
                        for (uint32_t p = s_seg_lo[sg] + tid; p < s_seg_hi[sg]; p += kThreads, ++r)
                        {
                            uint32_t x = my_syn >> r;
                            for (int k = 0; k < w; ++k)
                            {
                                const uint32_t col = col_of_slot[code.base[k] + p];
                                x ^= s_z[col >> 5] >> (col & 31);
                            }
                            wrong |= x;
                        }
                    }
                    if (!__syncthreads_or((int)(wrong & 1u)))
                    {
                        success = true; // :285-298
                        break;
                    }
                }
                const bool first = kReconcile && it == 0;
                uint32_t bad = 0;
                {
                    uint32_t my_syn = first ? 0u : s_park_syn[tid];
                    const uint32_t *zsrc = first ? s_alice : s_z;
                    int rbit = 0;
#pragma unroll 1
                    for (int sg = 0; sg < nseg; ++sg)
                    {
                        const uint32_t lo = s_seg_lo[sg], hi = s_seg_hi[sg];
                        switch (s_seg_w[sg])
                        {
#define QLB_SEG64(W_) case W_: bad |= check_segment64<Math, W_>(kThreads, msg, Base64FromParams{args}, col_of_slot, zsrc, lo, hi, my_syn, rbit, first, en, thr); break;
#define QLB_SEG64W(W_) case W_: { const uint2 rv = check_segment64_wide<Math, W_>(kThreads, msg, s_base, col_of_slot, zsrc, lo, hi, my_syn, rbit, first, en, thr); \
                                  bad |= rv.x & 1u; rbit = (int)(rv.x >> 1); my_syn = rv.y; } break;
                            QLB_SEG64(1) QLB_SEG64(2) QLB_SEG64(3) QLB_SEG64(4) QLB_SEG64(5) QLB_SEG64(6) QLB_SEG64(7) QLB_SEG64(8)
                            QLB_SEG64W(9) QLB_SEG64W(10) QLB_SEG64W(11) QLB_SEG64W(12) QLB_SEG64W(13) QLB_SEG64W(14) QLB_SEG64W(15) QLB_SEG64W(16)
#undef QLB_SEG64
#undef QLB_SEG64W
                        default: // checks without edges: satisfied only by a zero syndrome bit
                            for (uint32_t p = lo + tid; p < hi; p += kThreads, ++rbit)
                                if (!first)
                                    bad |= (my_syn >> rbit) & 1u;
                            break;
                        }
                    }
                    if (first)
                    {
                        s_park_syn[tid] = my_syn;
                        if (args.syndrome_out)
                        {
                            int r = 0;
                            for (int sg = 0; sg < nseg; ++sg)
                                for (uint32_t p = s_seg_lo[sg] + tid; p < s_seg_hi[sg]; p += kThreads, ++r)
                                    if ((my_syn >> r) & 1u)
                                    {
                                        const uint32_t j = code.check_order[p];
                                        atomicOr(&s_synn[j >> 5], 1u << (j & 31));
                                    }
                        }
                    }
                }
                const int any_bad = __syncthreads_count((int)(bad & 1u)); // threads with an unsatisfied check
                if (it > 0 && !any_bad)
                {
                    success = true; // :285-298
                    break;
                }
                quiet = it > 0 && any_bad <= kQuietChecks;
                if (it == args.max_it)
                    break; // :337-344
                // bit pass: total (:256-258), decision (:259-266), extrinsic + clamp (:300-316)
                {
                    uint32_t my_bob = s_park_bob[tid];
                    const uint16_t *bs = bslot + tid;
                    uint32_t *zw = s_z + (tid >> 5);

#pragma unroll 1
                    for (int i = tid; i < n; i += kThreads)
                    {
                        double prior;
                        if (kReconcile)
                        {
                            prior = (my_bob & 1u) ? -lp : lp;
                            my_bob >>= 1;
                        }
                        else
                            prior = llr_f[i];
                        uint32_t sl[kBW];
                        double c[kBW];
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            sl[a] = bs[a * n];
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            c[a] = msg.ld(sl[a]);
                        double total = prior;
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            total = total + c[a];
                        const bool z = total <= 0.;
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            msg.st(sl[a], clamp_msg(total - c[a], thr, en));
                        const uint32_t word = __ballot_sync(0xffffffffu, z);
                        if (lane == 0)
                            *zw = word;
                        bs += kThreads;
                        zw += kThreads / 32;
                    }
                }
                ++it;
                __syncthreads();
            }

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
