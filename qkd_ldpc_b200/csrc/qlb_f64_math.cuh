// Branch-free double-precision exp / log / divide for the fp64 check rule.
//
// ncu on the fp64 kernels showed that 70 % of the issued instructions inside libdevice's exp / log / divide are not
// FP64 arithmetic but their scaffolding: 64-bit polynomial constants materialised through UMOV pairs, register moves and
// special-case branches (profiles/r01_summary.md). These versions keep every coefficient in constant memory (a direct DFMA
// operand), have no branches, and are specialised to the argument ranges the decoder produces:
//   exp_neg(x)   x in [-64, 0]           Cody-Waite reduction by ln 2 (round-to-nearest through the 1.5*2^52 shift), degree-13
//                                        Taylor polynomial in Horner form, exponent insertion by integer add.  <= 1 ulp
//   div_any(a,d) d normal, either sign    MUFU.RCP64H seed (rcp.approx.ftz.f64: ONE instruction on the double's high word, no
//                                        fp64<->fp32 conversions), one Newton step, one residual correction: the quotient before
//                                        its final rounding is within 2^-80 of a / d. d = 0 gives NaN for a = 0 (the reference's
//                                        0/0) and for any other a (never asked); a subnormal d counts as 0 (ftz)
//   log_pos(y)   y >= 2^-60, finite      the classical fdlibm scheme: y = 2^k m, m in [sqrt(1/2), sqrt 2), s = f/(2+f), f = m-1,
//                                        log m = f - (f^2/2 - s (f^2/2 + R(s^2))), R = the degree-14 minimax polynomial
//                                        Lg1..Lg7 of fdlibm's e_log.c (published algorithm and constants).     <= 1 ulp
// Accuracy was measured on the CPU with the same FMA arithmetic against glibc over 4e6 random arguments each
// (/tmp prototype, numbers above). Like any libm they differ from glibc at the ulp level; the decoder's contract is the
// per-frame outcome (profiles/parity_r01.md).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb
{
    namespace f64m
    {
        static __constant__ double kExp[14] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                                               1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
        static __constant__ double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                                             1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};
        static __constant__ double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;

        __device__ __forceinline__ double exp_neg(double x)
        {
            const double kMagic = 6755399441055744.0; // 1.5 * 2^52: adding it rounds to the nearest integer in the low word
            const double t = fma(x, 1.4426950408889634, kMagic);
            const int k = __double2loint(t);
            const double kf = t - kMagic;
            double r = fma(kf, -kLn2Hi, x);
            r = fma(kf, -kLn2Lo, r);
            double p = kExp[13];
#pragma unroll
            for (int i = 12; i >= 0; --i)
                p = fma(p, r, kExp[i]);
            return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
        }

        // e^-min(|m|, ~1024) for ANY double m, with the cap and the |.| made on the integer pipe (2 ALU instructions + one on the
        // exponent) instead of the FP64 pipe's DADD / DSETP and two selects: the high word of |m| is limited to that of 1024.0
        // (an infinity becomes 1024, a NaN some value in [1024, 1025): callers that can see a NaN track it themselves), and the
        // binary exponent k is kept above -1000 so that the integer exponent insertion cannot wrap: beyond |m| ~ 693 the result is
        // a positive number far below 2^-54 instead of e^-|m|, which is all tanh(|m| / 2) = (1 - e) / (1 + e) = 1 needs.
        __device__ __forceinline__ double exp_neg_abs(double m)
        {
            const double kMagic = 6755399441055744.0;
            const double a = __hiloint2double(min(__double2hiint(m) & 0x7fffffff, 0x40900000), __double2loint(m));
            const double t = fma(-a, 1.4426950408889634, kMagic);
            const int k = max(__double2loint(t), -1000);
            const double kf = t - kMagic;
            double r = fma(kf, -kLn2Hi, -a);
            r = fma(kf, -kLn2Lo, r);
            double p = kExp[13];
#pragma unroll
            for (int i = 12; i >= 0; --i)
                p = fma(p, r, kExp[i]);
            return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
        }

        // MUFU.RCP64H seed (relative error ~2^-20) and ONE Newton step: y = (1/d)(1 + eps), |eps| <~ 2^-40. Either sign of d.
        __device__ __forceinline__ double rcp_any(double d)
        {
            double y0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
            return fma(y0, fma(-d, y0, 1.0), y0);
        }
        // q = a y is ~40 bits good; the residual a - d q is exact in an FMA and |residual * y| carries a relative error of 2^-40
        // again, so the corrected quotient is the rounding of a / d (1 + O(2^-80)): a second Newton step on y would buy nothing.
        __device__ __forceinline__ double div_any(double a, double d)
        {
            const double y = rcp_any(d);
            const double q = a * y;
            return fma(fma(-d, q, a), y, q);
        }
        __device__ __forceinline__ double div_pos(double a, double d) { return div_any(a, d); }

        __device__ __forceinline__ double log_pos(double y)
        {
            const int hi = __double2hiint(y);
            int k = (hi >> 20) - 1023;
            double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(y));
            const bool big = m > 1.4142135623730951;
            m = big ? 0.5 * m : m;
            k += big ? 1 : 0;
            const double f = m - 1.0;
            const double s = div_pos(f, 2.0 + f);
            const double z = s * s, w = z * z;
            const double t1 = w * fma(w, fma(w, kLg[5], kLg[3]), kLg[1]);
            const double t2 = z * fma(w, fma(w, fma(w, kLg[6], kLg[4]), kLg[2]), kLg[0]);
            const double R = t2 + t1;
            const double hfsq = 0.5 * f * f, dk = (double)k;
            return dk * kLn2Hi - ((hfsq - (s * (hfsq + R) + dk * kLn2Lo)) - f);
        }

        // ln(num / den) for 0 < den <= num without forming the quotient: ONE division for the ratio and the logarithm together.
        // den is rescaled by the power of two 2^j that brings m = num / (den 2^j) into ~[sqrt(1/2), sqrt 2] (exponent difference,
        // corrected by one through an integer comparison of the leading mantissa bits; the boundary need not be exact, the
        // polynomial holds a little beyond it). Then s = (m - 1) / (m + 1) = (num - den') / (num + den') -- the numerator is exact
        // (Sterbenz) -- and ln m = 2 s + s R(s^2) with fdlibm's R (e_log.c: "log(1+f) = 2s + s*R"), result j ln 2 + ln m. <= 2 ulp.
        // (the special outcomes are selected by the caller from `num`, `den`: for operands outside 0 < den <= num the value
        // returned here is meaningless, but nothing traps)
        __device__ __forceinline__ double log_ratio(double num, double den)
        {
            const int hn = __double2hiint(num), hd = __double2hiint(den);
            int j = (hn >> 20) - (hd >> 20);
            // leading 16 mantissa bits (with the hidden one) of both: u / d in (1/2, 2)
            const uint32_t u = (((uint32_t)hn & 0x000fffffu) | 0x00100000u) >> 5, d = (((uint32_t)hd & 0x000fffffu) | 0x00100000u) >> 5;
            j += (u * 46341u > d * 65536u) ? 1 : 0; // u / d > sqrt 2
            j -= (u * 65536u < d * 46341u) ? 1 : 0; // u / d < sqrt(1/2)
            const double ds = __hiloint2double(hd + (j << 20), __double2loint(den));
            const double s = div_pos(num - ds, num + ds);
            const double z = s * s, w = z * z;
            const double t1 = w * fma(w, fma(w, kLg[5], kLg[3]), kLg[1]);
            const double t2 = z * fma(w, fma(w, fma(w, kLg[6], kLg[4]), kLg[2]), kLg[0]);
            const double dj = (double)j;
            return fma(dj, kLn2Hi, fma(2., s, fma(s, t2 + t1, dj * kLn2Lo)));
        }

        // tanh(m / 2) = (1 - e^-|m|) / (1 + e^-|m|); |m| capped (exp_neg_abs; the quotient is exactly +-1 far earlier).
        // m = NaN comes back as +-1, NOT NaN: callers that can see a NaN message track it themselves (MathF64::check poisons the
        // row product, TwoPass<MathF64>::t passes the NaN through) so that the hot loop pays one compare instead of a select pair.
        __device__ __forceinline__ double tanh_half(double m)
        {
            const double e = exp_neg_abs(m);
            return copysign(div_any(1. - e, 1. + e), m);
        }
        // 2 atanh(p) = ln((1 + |p|) / (1 - |p|)) with p's sign, |p| <= 1 (the caller's p is a product of tanh values divided by one
        // of its own factors, which cannot exceed 1 in magnitude: every partial product of factors <= 1 is <= each factor, and the
        // quotient of a <= d rounds to <= 1). ONE division serves the ratio and the logarithm (log_ratio). Edges: |p| = 1 gives
        // log_ratio(2, 0) = 1024 ln 2 -- finite, above any clamp the decoder can apply -- unless `want_inf`, the IEEE outcome +-inf
        // of the literal expression (needed only when the clamp is off); p = NaN gives NaN through the arithmetic itself.
        __device__ __forceinline__ double two_atanh_mag(double p, bool want_inf)
        {
            const double a = fabs(p);
            const double den = 1. - a;
            double r = log_ratio(1. + a, den);
            if (want_inf)
                r = den == 0. ? __longlong_as_double(0x7ff0000000000000LL) : r;
            return r; // >= 0 or NaN; the caller clamps the magnitude and then gives it p's sign
        }
        __device__ __forceinline__ double two_atanh(double p, bool want_inf) { return copysign(two_atanh_mag(p, want_inf), p); }
    }
}
