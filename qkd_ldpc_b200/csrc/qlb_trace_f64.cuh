// Trace form of the fp64 decoder: ONE frame on one CTA, every intermediate the reference prints under CFG.TRACE_SUM_PRODUCT
// (src/qkd_ldpc_algorithm.cpp:214-327) copied out per iteration. It walks the same slot layout and calls the same MathF64
// node rule as the throughput kernels, so its (iterations, success, decoded bits) equal theirs; speed is not a goal here.
//
// Per executed iteration t < capacity the kernel stores
//   snap_e[t][slots]  the message array after the check pass + clamp  = check_to_bit_msg (:228-249)
//   tot[t][n], z[t][n] total LLR and hard decision                     (:256-267)
//   s[t][m]           syndrome of the decision, natural check order    (:277)
//   snap_m[t][slots]  the message array after the bit pass + clamp     = bit_to_check_msg (:300-316); absent for the
//                     iteration that converges, as in the reference (it returns at :285-298 first)
// The host reorders the slot snapshots into the reference's jagged row order (qlb_sum_product_trace in qlb_api.cu).
#pragma once
#include "qlb_kernels.cuh"

namespace qlb
{
    constexpr int kTraceThreads = 1024;

    __global__ void __launch_bounds__(kTraceThreads, 1)
        trace_f64_kernel(const CodeDev code, const double *__restrict__ llr, const uint32_t *__restrict__ syn_packed, int max_it, int en,
                         double thr, int capacity, double *__restrict__ msg, int32_t *__restrict__ z_cur, double *__restrict__ snap_e,
                         double *__restrict__ snap_m, double *__restrict__ tot_out, int32_t *__restrict__ z_out, int32_t *__restrict__ s_out,
                         uint32_t *__restrict__ iterations, uint8_t *__restrict__ result)
    {
        const int tid = threadIdx.x, n = code.n, m = code.m, max_cw = code.max_check_w, max_bw = code.max_bit_w;
        const uint32_t *bit_slots = code.bit_slots32;
        constexpr uint32_t kNone = 0xFFFFFFFFu;

        // initialisation: every edge carries the prior of its bit (:182-190)
        for (int i = tid; i < n; i += kTraceThreads)
            for (int a = 0; a < max_bw; ++a)
            {
                const uint32_t s = bit_slots[(size_t)a * n + i];
                if (s != kNone)
                    msg[s] = llr[i];
            }
        __syncthreads();

        int it = 0;
        bool success = false;
        while (it < max_it)
        {
            const bool keep = it < capacity;
            // check pass (:220-249): tanh of every incoming message, product seeded by the syndrome bit, divide, 2 atanh, clamp
            for (int p = tid; p < m; p += kTraceThreads)
            {
                const uint32_t j = code.check_order[p];
                const bool sbit = (syn_packed[j >> 5] >> (j & 31)) & 1u;
                double row = sbit ? -1. : 1.;
                for (int k = 0; k < max_cw; ++k)
                    if ((uint32_t)p < code.cnt[k])
                    {
                        const double t = TwoPass<MathF64>::t(msg[code.base[k] + p]); // NaN stays NaN
                        msg[code.base[k] + p] = t;
                        row *= t;
                    }
                for (int k = 0; k < max_cw; ++k)
                    if ((uint32_t)p < code.cnt[k])
                        msg[code.base[k] + p] = TwoPass<MathF64>::out(MathF64::quotient(row, msg[code.base[k] + p]), en != 0, thr);
            }
            __syncthreads();
            if (keep)
                for (int s = tid; s < code.slots; s += kTraceThreads)
                    snap_e[(size_t)it * code.slots + s] = msg[s];

            // totals and hard decisions (:256-267)
            for (int i = tid; i < n; i += kTraceThreads)
            {
                double total = llr[i];
                for (int a = 0; a < max_bw; ++a)
                {
                    const uint32_t s = bit_slots[(size_t)a * n + i];
                    if (s != kNone)
                        total = total + msg[s];
                }
                const int z = (total <= 0.) ? 1 : 0;
                z_cur[i] = z;
                if (keep)
                {
                    tot_out[(size_t)it * n + i] = total;
                    z_out[(size_t)it * n + i] = z;
                }
            }
            __syncthreads();

            // syndrome of the decision (:277) against the target (:285)
            int bad = 0;
            for (int j = tid; j < m; j += kTraceThreads)
            {
                int sb = 0;
                for (int q = code.row_ptr[j]; q < code.row_ptr[j + 1]; ++q)
                    sb ^= z_cur[code.col_idx[q]];
                if (keep)
                    s_out[(size_t)it * m + j] = sb;
                bad |= sb ^ (int)((syn_packed[j >> 5] >> (j & 31)) & 1u);
            }
            if (!__syncthreads_or(bad))
            {
                success = true;
                ++it;
                break;
            }

            // extrinsic bit-to-check messages (:300-316)
            for (int i = tid; i < n; i += kTraceThreads)
            {
                double total = llr[i];
                for (int a = 0; a < max_bw; ++a)
                {
                    const uint32_t s = bit_slots[(size_t)a * n + i];
                    if (s != kNone)
                        total = total + msg[s];
                }
                for (int a = 0; a < max_bw; ++a)
                {
                    const uint32_t s = bit_slots[(size_t)a * n + i];
                    if (s != kNone)
                        msg[s] = clamp_msg(total - msg[s], thr, en != 0);
                }
            }
            __syncthreads();
            if (keep)
                for (int s = tid; s < code.slots; s += kTraceThreads)
                    snap_m[(size_t)it * code.slots + s] = msg[s];
            ++it;
        }
        if (tid == 0)
        {
            *iterations = (uint32_t)it; // max_it on failure (:344)
            *result = success ? 1 : 0;
        }
    }
}
