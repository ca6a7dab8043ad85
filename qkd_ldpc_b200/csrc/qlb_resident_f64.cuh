// fp64 SM-resident decoder: the reference's arithmetic and operation order (MathF64, qlb_kernels.cuh) with the frame's
// messages kept on the SM instead of in an L2-resident scratch.
//
// Storage. 30 720 doubles (240 KB) do not fit in the 227 KB of shared memory of one SM, so the message array is SPLIT: the
// first `smem_slots` physical slots (~92 % on the N=10240 code: all of edge positions 0..4 and the head of position 5) live
// in shared memory, the tail in a small per-CTA global scratch (L2-resident). Two properties of the code layout
// (qlb_layout.hpp) keep the split out of the hot loops:
//   * a run of sorted checks of one weight sees the split the same way for every check -- all rows in shared memory, or
//     exactly the last row in the tail -- so the check pass addresses its messages without a per-access test;
//   * inside a weight class the checks are ordered by the bit index of their last edge, so the tail slots belong to bits of
//     high index only: ~85 % of the 32-bit groups of the bit pass never leave shared memory and run a loop without any
//     shared / global predicate (`fast` groups); the others take the generic path.
//
// Schedule. Work is dealt to WARPS in groups of 32 consecutive nodes through a shared-memory counter (the slow groups first):
// a warp that finishes early takes the next group, so the two block barriers of an iteration wait for one group at most,
// not for the slowest of 24 fixed shares (ncu of the fixed-share version: 16 % of the time in those barriers). The tables
// that describe the groups (built once per code and device by resident64_build_tables) sit in shared memory.
//
// Convergence (calculate_syndrome + arrays_equal after every bit pass, src/qkd_ldpc_algorithm.cpp:277-298) is tracked
// INCREMENTALLY, in integers, exactly: s_unsat holds one bit per check, syndrome(z) ^ target. It is initialised once per
// frame from the hard decision of the priors (a walk over the slot->bit table), and from then on a bit whose decision
// flips in a bit pass toggles the bits of its checks. The frame has converged when the words are all zero -- known right
// after the bit pass, so a converged frame pays no speculative check pass, and the check pass itself carries nothing but
// the check rule (bit-exact fp64 messages leave no spare mantissa bit to carry the decision as the fp32 kernel does).
// Iteration counts, flags and keys follow the reference's definitions exactly (src/qkd_ldpc_algorithm.cpp:175-345,
// 398-447); the generic decode_kernel<MathF64> stays as the fallback for codes this kernel does not take.
#pragma once
#include "qlb_resident_f32.cuh"

namespace qlb
{
    constexpr int kResident64Threads = 768; // launch bound (80 registers)
    constexpr int kResident64FastW = 8;     // weights up to this get the split-specialised loops
    constexpr size_t kResident64StaticSmem = 256;

    // A group of the check walk: sorted positions p0 + l0 .. p0 + l1 (p0 a multiple of 32), all of weight w, the split falling
    // the same way for each: `tail` rows (edge positions w - tail .. w - 1) in the global tail (0 or 1), or 3: decide per access.
    //   bits 0..10  p0 / 32     bits 11..15  l0     bits 16..20  l1     bits 21..27  type = w * 4 + tail
    __host__ __device__ inline uint32_t r64_pack_group(uint32_t p0, int l0, int l1, int w, int tail)
    {
        return (p0 >> 5) | ((uint32_t)l0 << 11) | ((uint32_t)l1 << 16) | ((uint32_t)(w * 4 + tail) << 21);
    }
    // A group of the bit walk: bits 32 g .. 32 g + 31; bit 15 set: every slot of the group lies in shared memory.
    constexpr uint16_t kR64BitGroupFast = 0x8000;

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

    // shared memory beside the messages: Bob, decisions | target syndrome, unsatisfied checks (both in sorted check order) | tables
    __host__ __device__ inline size_t resident64_small_bytes(int n, int m, int check_groups, int bit_groups)
    {
        const size_t wn = align_up((size_t)(n + 31) / 32 * 4, 16), wm = align_up((size_t)(m + 31) / 32 * 4, 16);
        return 2 * wn + 2 * wm + align_up((size_t)check_groups * 4, 16) + align_up((size_t)bit_groups * 2, 16);
    }

    // The clamp of :313-316 (threshold_matrix: x > thr -> thr, x < -thr -> -thr, NaN untouched) with ONE FP64 compare.
    __device__ __forceinline__ double clamp_f64(double x, double thr, bool en)
    {
        if (en && fabs(x) > thr)
            x = copysign(thr, x);
        return x;
    }

    struct Base64FromParams // constant-bank operands
    {
        const DecodeArgs &args;
        __device__ __forceinline__ uint32_t operator()(int k) const { return args.code.base[k]; }
    };
    struct Base64FromSmem
    {
        const uint32_t *base;
        __device__ __forceinline__ uint32_t operator()(int k) const { return base[k]; }
    };

    // The check rule (:220-249) on the checks p0 + lane of one group.
    template <typename Math, int W, int kTail, typename Base>
    __device__ __forceinline__ void check_group64(const Split64 &msg, const Base s_base, uint32_t p, bool active, bool syn, bool en, double thr)
    {
        if (!active)
            return;
        double *gm = msg.gmem - msg.smem_slots;
        double v[W];
#pragma unroll
        for (int k = 0; k < W; ++k)
        {
            const uint32_t slot = s_base(k) + p;
            v[k] = kTail == 3 ? msg.ld(slot) : (k < W - kTail ? msg.smem[slot] : gm[slot]);
        }
        Math::template check<W>(v, W, syn, en, thr);
#pragma unroll
        for (int k = 0; k < W; ++k)
        {
            const uint32_t slot = s_base(k) + p;
            if (kTail == 3)
                msg.st(slot, v[k]);
            else if (k < W - kTail)
                msg.smem[slot] = v[k];
            else
                gm[slot] = v[k];
        }
    }
    // Weights 9..16 (the R >= 0.7 codes of the CW = 3 family) are kept out of line: their register appetite (2 x W doubles live) must
    // not leak into the hot instantiations; some spilling inside them is still far cheaper than the generic kernel's L2 round trips.
    template <typename Math, int W>
    __device__ __noinline__ void check_group64_wide(const Split64 msg, const uint32_t *s_base, uint32_t p, bool active, bool syn, bool en, double thr)
    {
        check_group64<Math, W, 3>(msg, Base64FromSmem{s_base}, p, active, syn, en, thr);
    }

    // Sorted position of the check that owns `slot` (rows are laid out one edge position after the other, qlb_layout.hpp).
    __device__ __forceinline__ uint32_t check_of_slot64(uint32_t slot, const uint32_t *s_base, int max_cw)
    {
        uint32_t b = 0;
        for (int k = 1; k < max_cw; ++k)
            b = slot >= s_base[k] ? s_base[k] : b;
        return slot - b;
    }

    // Next group of a walk for this warp: lane 0 draws from the shared counter.
    __device__ __forceinline__ int draw_group(int *counter, int lane)
    {
        int g = 0;
        if (lane == 0)
            g = atomicAdd(counter, 1);
        return g;
    }

    // kBW: uniform bit weight. Host-checked: slots < 65535, n < 65536, m < 65536, max_check_w <= 16, n % 32 == 0.
    template <typename Math, bool kReconcile, int kBW, int kMaxThreads>
    __global__ void __launch_bounds__(kMaxThreads, 1) decode_resident_f64_kernel(const DecodeArgs args)
    {
        const int kThreads = blockDim.x; // multiple of 32, <= kMaxThreads
        extern __shared__ __align__(16) unsigned char smem[];
        __shared__ uint32_t s_base[kResidentMaxCW];
        __shared__ int s_ctr[2]; // next group of the check walk / of the bit walk
        __shared__ long long s_frame;

        const CodeDev &code = args.code;
        const int n = code.n, m = code.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = kThreads >> 5;
        const int words_n = code.words_n, words_m = code.words_m;
        const size_t wn = align_up((size_t)words_n * 4, 16), wm = align_up((size_t)words_m * 4, 16);
        const int n_cg = code.r64_check_groups, n_bg = code.r64_bit_groups;
        const uint32_t smem_slots = code.r64_smem_slots;

        Split64 msg;
        msg.smem = reinterpret_cast<double *>(smem);
        msg.smem_slots = smem_slots;
        msg.gmem = reinterpret_cast<double *>(args.scratch + (size_t)blockIdx.x * args.scratch_stride);
        unsigned char *tail = smem + (size_t)smem_slots * 8;
        uint32_t *s_bob = reinterpret_cast<uint32_t *>(tail);
        uint32_t *s_z = reinterpret_cast<uint32_t *>(tail + wn);
        uint32_t *s_syn = reinterpret_cast<uint32_t *>(tail + 2 * wn);         // target syndrome, sorted check order
        uint32_t *s_unsat = reinterpret_cast<uint32_t *>(tail + 2 * wn + wm);  // syndrome(z) ^ target, sorted check order
        uint32_t *s_cg = reinterpret_cast<uint32_t *>(tail + 2 * wn + 2 * wm);
        uint16_t *s_bg = reinterpret_cast<uint16_t *>(tail + 2 * wn + 2 * wm + align_up((size_t)n_cg * 4, 16));
        // frame set-up only, before the messages are written: Alice's key and the syndrome in natural check order
        uint32_t *t_alice = reinterpret_cast<uint32_t *>(smem);
        uint32_t *t_synn = reinterpret_cast<uint32_t *>(smem + wn);

        if (tid < kResidentMaxCW)
            s_base[tid] = code.base[tid];
        for (int g = tid; g < n_cg; g += kThreads)
            s_cg[g] = code.r64_check_group_table[g];
        for (int g = tid; g < n_bg; g += kThreads)
            s_bg[g] = code.r64_bit_group_table[g];
        const uint16_t *bslot = code.bit_slots16;
        const uint16_t *col_of_slot = code.col_of_slot16;
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

            // ---- A: keys / target syndrome into shared memory; z0 = the priors' own hard decision ---------------------------
            double lp = 0.;
            const double *llr_f = nullptr;
            if (kReconcile)
            {
                lp = args.log_prior[f];
                for (int w = tid; w < words_n; w += kThreads)
                {
                    s_bob[w] = args.bob[f * words_n + w];
                    t_alice[w] = args.alice[f * words_n + w];
                }
                for (int w = tid; w < words_m; w += kThreads)
                    t_synn[w] = 0;
            }
            else
            {
                llr_f = args.llr + f * n;
                for (int w = tid; w < words_m; w += kThreads)
                    t_synn[w] = args.syndrome_in[f * words_m + w];
            }
            for (int w = tid; w < words_m; w += kThreads)
            {
                s_unsat[w] = 0;
                s_syn[w] = 0;
            }
            __syncthreads();
            for (int g = warp; g < words_n; g += warps) // n % 32 == 0: whole warps only
            {
                const int i = g * 32 + lane;
                double prior;
                if (kReconcile)
                    prior = ((s_bob[g] >> lane) & 1u) ? -lp : lp; // :401-405
                else
                    prior = llr_f[i];
                const uint32_t word = __ballot_sync(0xffffffffu, prior <= 0.);
                if (lane == 0)
                    s_z[g] = word;
            }
            __syncthreads();
            // ---- B: one walk over the checks through the slot -> bit table: the target syndrome (reconcile mode: Alice's
            // syndrome, :413-414) and s_unsat <- syndrome(z0) ^ target, both in sorted check order ---------------------------
            for (int g = warp; g < n_cg; g += warps)
            {
                const uint32_t ent = s_cg[g];
                const uint32_t p = ((ent & 0x7ffu) << 5) + lane;
                const bool active = lane >= (int)((ent >> 11) & 31u) && lane <= (int)((ent >> 16) & 31u);
                const int w = (int)(ent >> 23);
                uint32_t sb = 0, pz = 0;
                if (active)
                {
                    uint32_t pa = 0;
                    for (int k = 0; k < w; ++k)
                    {
                        const uint32_t col = col_of_slot[s_base[k] + p];
                        pz ^= s_z[col >> 5] >> (col & 31);
                        if (kReconcile)
                            pa ^= t_alice[col >> 5] >> (col & 31);
                    }
                    const uint32_t j = code.check_order[p];
                    if (kReconcile)
                    {
                        sb = pa & 1u;
                        if (sb && args.syndrome_out)
                            atomicOr(&t_synn[j >> 5], 1u << (j & 31));
                    }
                    else
                        sb = (t_synn[j >> 5] >> (j & 31)) & 1u;
                }
                const uint32_t syn_word = __ballot_sync(0xffffffffu, active && sb);
                const uint32_t unsat_word = __ballot_sync(0xffffffffu, active && ((sb ^ pz) & 1u));
                if (lane == 0) // two groups can share a word (a weight class ending inside it)
                {
                    if (syn_word)
                        atomicOr(&s_syn[p >> 5], syn_word);
                    if (unsat_word)
                        atomicOr(&s_unsat[p >> 5], unsat_word);
                }
            }
            __syncthreads();
            if (kReconcile && args.syndrome_out)
            {
                for (int w = tid; w < words_m; w += kThreads)
                    args.syndrome_out[f * words_m + w] = t_synn[w];
                __syncthreads();
            }
            // ---- C: messages <- priors, unclamped (:182-190) -----------------------------------------------------------------
            for (int g = warp; g < words_n; g += warps)
            {
                const int i = g * 32 + lane;
                double prior;
                if (kReconcile)
                    prior = ((s_bob[g] >> lane) & 1u) ? -lp : lp;
                else
                    prior = llr_f[i];
#pragma unroll
                for (int a = 0; a < kBW; ++a)
                    msg.st(bslot[a * n + i], prior);
            }
            if (tid == 0)
                s_ctr[0] = 0;
            __syncthreads();

            int it = 0; // completed bit passes
            bool success = false;
            while (it < args.max_it)
            {
                // ---- check pass (:220-249) -----------------------------------------------------------------------------------
                if (tid == 0)
                    s_ctr[1] = 0; // every warp has left the previous bit pass
                {
                    int nxt = draw_group(&s_ctr[0], lane);
                    for (;;)
                    {
                        const int g = __shfl_sync(0xffffffffu, nxt, 0);
                        if (g >= n_cg)
                            break;
                        nxt = draw_group(&s_ctr[0], lane);
                        const uint32_t ent = s_cg[g];
                        const uint32_t p = ((ent & 0x7ffu) << 5) + lane;
                        const bool active = lane >= (int)((ent >> 11) & 31u) && lane <= (int)((ent >> 16) & 31u);
                        const bool syn = ((s_syn[ent & 0x7ffu] >> lane) & 1u) != 0;
                        switch (ent >> 21)
                        {
#define QLB_GRP64(W_)                                                                                            \
    case W_ * 4 + 0: check_group64<Math, W_, 0>(msg, Base64FromParams{args}, p, active, syn, en, thr); break;     \
    case W_ * 4 + 1: check_group64<Math, W_, 1>(msg, Base64FromParams{args}, p, active, syn, en, thr); break;     \
    case W_ * 4 + 3: check_group64<Math, W_, 3>(msg, Base64FromParams{args}, p, active, syn, en, thr); break;
#define QLB_GRP64W(W_) \
    case W_ * 4 + 3: check_group64_wide<Math, W_>(msg, s_base, p, active, syn, en, thr); break;
                            QLB_GRP64(1) QLB_GRP64(2) QLB_GRP64(3) QLB_GRP64(4) QLB_GRP64(5) QLB_GRP64(6) QLB_GRP64(7) QLB_GRP64(8)
                            QLB_GRP64W(9) QLB_GRP64W(10) QLB_GRP64W(11) QLB_GRP64W(12) QLB_GRP64W(13) QLB_GRP64W(14) QLB_GRP64W(15) QLB_GRP64W(16)
#undef QLB_GRP64
#undef QLB_GRP64W
                        default: // checks without edges send nothing
                            break;
                        }
                    }
                }
                __syncthreads();
                // ---- bit pass: total (:256-258), decision (:259-266), extrinsic + clamp (:300-316); a flipped decision toggles
                // the unsatisfied-bits of the bit's checks ------------------------------------------------------------------
                if (tid == 0)
                    s_ctr[0] = 0; // every warp has left the check pass
                {
                    int nxt = draw_group(&s_ctr[1], lane);
                    int g = __shfl_sync(0xffffffffu, nxt, 0);
                    uint32_t ent = 0, sl[kBW];
                    if (g < n_bg)
                    {
                        ent = s_bg[g];
                        nxt = draw_group(&s_ctr[1], lane);
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            sl[a] = bslot[a * n + (int)(ent & 0x7fffu) * 32 + lane];
                    }
                    while (g < n_bg)
                    {
                        // the next group's slot indices (L2) travel while this group's messages are gathered
                        const int g2 = __shfl_sync(0xffffffffu, nxt, 0);
                        uint32_t ent2 = 0, sl2[kBW];
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            sl2[a] = 0;
                        if (g2 < n_bg)
                        {
                            ent2 = s_bg[g2];
                            nxt = draw_group(&s_ctr[1], lane);
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                sl2[a] = bslot[a * n + (int)(ent2 & 0x7fffu) * 32 + lane];
                        }
                        const int grp = (int)(ent & 0x7fffu);
                        double prior;
                        if (kReconcile)
                            prior = ((s_bob[grp] >> lane) & 1u) ? -lp : lp;
                        else
                            prior = llr_f[grp * 32 + lane];
                        double c[kBW];
                        bool z;
                        if (ent & kR64BitGroupFast)
                        {
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                c[a] = msg.smem[sl[a]];
                            double total = prior;
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                total = total + c[a];
                            z = total <= 0.;
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                msg.smem[sl[a]] = clamp_f64(total - c[a], thr, en);
                        }
                        else
                        {
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                c[a] = msg.ld(sl[a]);
                            double total = prior;
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                total = total + c[a];
                            z = total <= 0.;
#pragma unroll
                            for (int a = 0; a < kBW; ++a)
                                msg.st(sl[a], clamp_f64(total - c[a], thr, en));
                        }
                        const uint32_t word = __ballot_sync(0xffffffffu, z);
                        const uint32_t flips = word ^ s_z[grp];
                        if ((flips >> lane) & 1u)
                        {
#pragma unroll 1
                            for (int a = 0; a < kBW; ++a)
                            {
                                const uint32_t p = check_of_slot64(sl[a], s_base, code.max_check_w);
                                atomicXor(&s_unsat[p >> 5], 1u << (p & 31));
                            }
                        }
                        __syncwarp();
                        if (lane == 0 && flips)
                            s_z[grp] = word;
                        g = g2;
                        ent = ent2;
#pragma unroll
                        for (int a = 0; a < kBW; ++a)
                            sl[a] = sl2[a];
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
                    differs |= (zw != args.alice[f * words_n + w]);
            }
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
