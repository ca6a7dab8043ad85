// fp64 SM-resident decoder: the reference's arithmetic and operation order (MathF64, qlb_kernels.cuh) with the frame's
// messages kept on the SM instead of in an L2-resident scratch.
//
// 30 720 doubles (240 KB) do not fit in the 227 KB of shared memory of one SM, so the message array is SPLIT: the first
// `smem_slots` physical slots (~92 % on the N=10240 code: all of edge positions 0..4 and the head of position 5) live in
// shared memory, the tail in a small per-CTA global scratch (coalesced in the check pass, L2 hits in the bit pass). The
// host cuts the walk over the sorted checks into segments (SegTable64) inside which the split falls the same way for every
// check -- all rows in shared memory, or exactly the last row in the tail -- so the hot loops address their messages
// without a per-access test.
//
// Convergence (calculate_syndrome + arrays_equal after every bit pass, src/qkd_ldpc_algorithm.cpp:277-298) is tracked
// INCREMENTALLY, in integers, exactly: s_unsat holds one bit per check, syndrome(z) ^ target. It is initialised once per
// frame from the hard decision of the priors (a walk over the slot->bit table), and from then on a bit whose decision
// flips in a bit pass toggles the bits of its checks (a few hundred shared-memory atomics in the first rounds, a handful
// later, against a 30 720-edge gather per round before). The frame has converged when the words are all zero -- known right
// after the bit pass, so a converged frame no longer pays a speculative check pass, and the check pass itself carries
// nothing but the check rule (bit-exact fp64 messages leave no spare mantissa bit to carry the decision as the fp32 kernel
// does). Iteration counts, flags and keys follow the reference's definitions exactly (src/qkd_ldpc_algorithm.cpp:175-345,
// 398-447); the generic decode_kernel<MathF64> stays as the fallback for codes this kernel does not take.
#pragma once
#include "qlb_resident_f32.cuh"

namespace qlb
{
    constexpr int kResident64Threads = 768; // launch bound; the host picks the block size that balances the node walks
    constexpr size_t kResident64StaticSmem = 2 * kResident64Threads * 4 + 1024;
    constexpr int kResident64MaxSegs = 24;
    constexpr int kResident64FastW = 8; // weights up to this get the split-specialised loops

    // One run of sorted check positions [lo, hi): weight exactly w; `tail` rows (edge positions w - tail .. w - 1) live in the
    // global tail for every check of the run (0 or 1), or tail = -1: decide per access.
    struct Seg64
    {
        uint32_t lo, hi;
        int32_t w, tail;
    };
    struct SegTable64
    {
        int32_t n;
        Seg64 seg[kResident64MaxSegs];
    };

    struct Split64
    {
        double *smem;
        double *gmem; // slot s >= smem_slots is gmem[s - smem_slots]
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
        return 3 * wn + 2 * wm; // Bob, Alice, decisions | syndrome (natural order), unsatisfied checks (sorted order)
    }

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

    // The check rule (:220-249) on all checks of one segment. kTail as Seg64::tail.
    template <typename Math, int W, int kTail, typename Base>
    __device__ __forceinline__ void check_segment64(int kThreads, const Split64 &msg, const Base base, uint32_t lo, uint32_t hi, uint32_t my_syn,
                                                    int &rbit, bool en, double thr)
    {
        double *gm = msg.gmem - msg.smem_slots;
#pragma unroll 1
        for (uint32_t p = lo + threadIdx.x; p < hi; p += kThreads, ++rbit)
        {
            double v[W];
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                const uint32_t slot = base(k) + p;
                v[k] = kTail < 0 ? msg.ld(slot) : (k < W - kTail ? msg.smem[slot] : gm[slot]);
            }
            Math::template check<W>(v, W, ((my_syn >> rbit) & 1u) != 0, en, thr);
#pragma unroll
            for (int k = 0; k < W; ++k)
            {
                const uint32_t slot = base(k) + p;
                if (kTail < 0)
                    msg.st(slot, v[k]);
                else if (k < W - kTail)
                    msg.smem[slot] = v[k];
                else
                    gm[slot] = v[k];
            }
        }
    }

    // Weights 9..16 (the R >= 0.7 codes of the CW = 3 family) are kept out of line: their register appetite (2 x W doubles live) must
    // not leak into the hot instantiations; some spilling inside them is still far cheaper than the generic kernel's L2 round trips.
    template <typename Math, int W>
    __device__ __noinline__ int check_segment64_wide(int kThreads, const Split64 msg, const uint32_t *s_base, uint32_t lo, uint32_t hi, uint32_t my_syn,
                                                     int rbit, bool en, double thr)
    {
        check_segment64<Math, W, -1>(kThreads, msg, Base64FromSmem{s_base}, lo, hi, my_syn, rbit, en, thr);
        return rbit;
    }

    // Sorted position of the check that owns `slot` (rows are laid out one edge position after the other, qlb_layout.hpp).
    static __device__ __noinline__ uint32_t check_of_slot64(uint32_t slot, const uint32_t *s_base, int max_cw)
    {
        int k = max_cw - 1;
        while (k > 0 && slot < s_base[k])
            --k;
        return slot - s_base[k];
    }

    // kBW: uniform bit weight. Host-checked: slots < 65535, n < 65536, max_check_w <= 16, n % 32 == 0, and every thread's walk over
    // the segments / the bits stays within 32 rounds (one register bit per node a thread visits).
    template <typename Math, bool kReconcile, int kBW, int kMaxThreads>
    __global__ void __launch_bounds__(kMaxThreads, 1) decode_resident_f64_kernel(const DecodeArgs args, const SegTable64 segs, uint32_t smem_slots,
                                                                                 const uint16_t *__restrict__ col_of_slot)
    {
        const int kThreads = blockDim.x; // multiple of 32, <= kMaxThreads
        extern __shared__ __align__(16) unsigned char smem[];
        __shared__ uint32_t s_park_bob[kMaxThreads], s_park_syn[kMaxThreads];
        __shared__ uint32_t s_base[kResidentMaxCW];
        __shared__ long long s_frame;

        const CodeDev &code = args.code;
        const int n = code.n, m = code.m, tid = threadIdx.x, lane = tid & 31;
        const int words_n = code.words_n, words_m = code.words_m;
        const size_t wn = align_up((size_t)words_n * 4, 16), wm = align_up((size_t)words_m * 4, 16);
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
        uint32_t *s_unsat = reinterpret_cast<uint32_t *>(tail + 3 * wn + wm);
        const uint16_t *bslot = code.bit_slots16;
        const int nseg = segs.n;
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
            for (int w = tid; w < words_m; w += kThreads)
                s_unsat[w] = 0;
            __syncthreads();

            // messages <- priors, unclamped (:182-190); z0 = the priors' own hard decision (the base of the incremental
            // syndrome); Bob's bit of the r-th bit this thread visits -> bit r of s_park_bob
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
                    const uint32_t word = __ballot_sync(0xffffffffu, prior <= 0.); // n % 32 == 0: whole warps only
                    if (lane == 0)
                        s_z[i >> 5] = word;
                }
                s_park_bob[tid] = my_bob;
            }
            __syncthreads();

            // One walk over this thread's checks through the slot -> bit table: the target syndrome bit of its r-th check -> bit
            // r of my_syn (reconcile mode: Alice's syndrome, :413-414), and s_unsat <- syndrome(z0) ^ target.
            {
                uint32_t my_syn = 0;
                int r = 0;
                for (int sg = 0; sg < nseg; ++sg)
                {
                    const int w = segs.seg[sg].w;
                    for (uint32_t p = segs.seg[sg].lo + tid; p < segs.seg[sg].hi; p += kThreads, ++r)
                    {
                        uint32_t pa = 0, pz = 0;
                        for (int k = 0; k < w; ++k)
                        {
                            const uint32_t col = col_of_slot[s_base[k] + p];
                            pz ^= s_z[col >> 5] >> (col & 31);
                            if (kReconcile)
                                pa ^= s_alice[col >> 5] >> (col & 31);
                        }
                        const uint32_t j = code.check_order[p];
                        uint32_t sb;
                        if (kReconcile)
                        {
                            sb = pa & 1u;
                            if (sb && args.syndrome_out)
                                atomicOr(&s_synn[j >> 5], 1u << (j & 31));
                        }
                        else
                            sb = (s_synn[j >> 5] >> (j & 31)) & 1u;
                        my_syn |= sb << r;
                        if ((sb ^ pz) & 1u)
                            atomicOr(&s_unsat[p >> 5], 1u << (p & 31));
                    }
                }
                s_park_syn[tid] = my_syn;
            }
            // (no barrier needed before the first check pass: it touches messages only; s_unsat is next touched after two barriers)

            int it = 0; // completed bit passes
            bool success = false;
            while (it < args.max_it)
            {
                // check pass (:220-249)
                {
                    const uint32_t my_syn = s_park_syn[tid];
                    int rbit = 0;
#pragma unroll 1
                    for (int sg = 0; sg < nseg; ++sg)
                    {
                        const uint32_t lo = segs.seg[sg].lo, hi = segs.seg[sg].hi;
                        switch (segs.seg[sg].w * 4 + (segs.seg[sg].tail & 3))
                        {
#define QLB_SEG64(W_)                                                                                                                     \
    case W_ * 4 + 0: check_segment64<Math, W_, 0>(kThreads, msg, Base64FromParams{args}, lo, hi, my_syn, rbit, en, thr); break;          \
    case W_ * 4 + 1: check_segment64<Math, W_, 1>(kThreads, msg, Base64FromParams{args}, lo, hi, my_syn, rbit, en, thr); break;          \
    case W_ * 4 + 3: check_segment64<Math, W_, -1>(kThreads, msg, Base64FromParams{args}, lo, hi, my_syn, rbit, en, thr); break;
#define QLB_SEG64W(W_) \
    case W_ * 4 + 3: rbit = check_segment64_wide<Math, W_>(kThreads, msg, s_base, lo, hi, my_syn, rbit, en, thr); break;
                            QLB_SEG64(1) QLB_SEG64(2) QLB_SEG64(3) QLB_SEG64(4) QLB_SEG64(5) QLB_SEG64(6) QLB_SEG64(7) QLB_SEG64(8)
                            QLB_SEG64W(9) QLB_SEG64W(10) QLB_SEG64W(11) QLB_SEG64W(12) QLB_SEG64W(13) QLB_SEG64W(14) QLB_SEG64W(15) QLB_SEG64W(16)
#undef QLB_SEG64
#undef QLB_SEG64W
                        default: // checks without edges send nothing
                            for (uint32_t p = lo + tid; p < hi; p += kThreads)
                                ++rbit;
                            break;
                        }
                    }
                }
                __syncthreads();
                // bit pass: total (:256-258), decision (:259-266), extrinsic + clamp (:300-316); a flipped decision toggles the
                // unsatisfied-bits of the bit's checks
                {
                    uint32_t my_bob = s_park_bob[tid];
                    const uint16_t *bs = bslot + tid;
                    uint32_t *zw = s_z + (tid >> 5);
                    uint32_t nx[kBW];
#pragma unroll
                    for (int a = 0; a < kBW; ++a)
                        nx[a] = tid < n ? bs[a * n] : 0;
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
                            sl[a] = nx[a];
                        bs += kThreads;
                        if (i + kThreads < n) // next round's slot indices (L2) while this round's messages are gathered
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                nx[a] = bs[a * n];
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
                        const uint32_t flips = word ^ *zw;
                        if ((flips >> lane) & 1u)
#pragma unroll 1
                            for (int a = 0; a < kBW; ++a)
                            {
                                const uint32_t p = check_of_slot64(sl[a], s_base, code.max_check_w);
                                atomicXor(&s_unsat[p >> 5], 1u << (p & 31));
                            }
                        __syncwarp();
                        if (lane == 0)
                            *zw = word;
                        zw += kThreads / 32;
                    }
                }
                ++it;
                __syncthreads();
                // calculate_syndrome + arrays_equal of :277-298: all checks satisfied?
                int open = 0;
                for (int w = tid; w < words_m; w += kThreads)
                    open |= s_unsat[w] != 0u;
                if (!__syncthreads_or(open))
                {
                    success = true; // :285-298
                    break;
                }
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
                args.iterations[f] = (uint32_t)it; // == max_it on failure (:344)
                args.result[f] = r;
                atomicAdd(args.iter_total, (unsigned long long)it);
            }
        }
    }
}
