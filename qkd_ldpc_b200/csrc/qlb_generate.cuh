// On-device key / error generator, bit-exact with the reference's run_trial inputs.
//
// The reference draws, per trial, from one xoshiro256++ stream seeded through SplitMix64 (third-party module
// Reputeless/Xoshiro-cpp v1.1; reference src/simulation.cpp:163):
//   Alice  : generate_random_bit_array (src/array_and_matrix_operations.cpp:424-431) -- uniform_int_distribution<int>(0,1)
//            per bit: libstdc++ 13.3 maps a 64-bit draw to [0, 2) by Lemire's multiply-shift, i.e. the draw's top bit;
//   Bob    : introduce_errors (:434-460) -- std::shuffle of 0..N-1 (libstdc++ 13.3 bits/stl_algo.h:3742-3806: one optional
//            single swap, then two swap positions per draw from uniform_int_distribution{0, (i+1)(i+2)-1}, with the
//            rejection loop of bits/uniform_int_dist.h:257-280), the first floor(N*q) entries are flipped.
// One thread replays one trial's stream sequentially (the stream and the Fisher-Yates chain are inherently serial, and a
// trial costs ~15 k draws -- a few percent of the cheapest decode); thousands of trials run side by side. During the
// forward shuffle position i still holds i when it is visited, so a swap is one dependent load and two stores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb
{
    struct Xoshiro256pp
    {
        uint64_t s0, s1, s2, s3;
        __device__ __forceinline__ explicit Xoshiro256pp(uint64_t seed)
        {
            uint64_t x = seed;
            s0 = splitmix(x);
            s1 = splitmix(x);
            s2 = splitmix(x);
            s3 = splitmix(x);
        }
        static __device__ __forceinline__ uint64_t splitmix(uint64_t &x)
        {
            uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            return z ^ (z >> 31);
        }
        static __device__ __forceinline__ uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
        __device__ __forceinline__ uint64_t next()
        {
            const uint64_t out = rotl(s0 + s3, 23) + s0;
            const uint64_t t = s1 << 17;
            s2 ^= s0;
            s3 ^= s1;
            s1 ^= s2;
            s0 ^= s3;
            s2 ^= t;
            s3 = rotl(s3, 45);
            return out;
        }
        // uniform on [0, range): Lemire's method as libstdc++ 13.3 implements it (bits/uniform_int_dist.h:257-280)
        __device__ __forceinline__ uint64_t below(uint64_t range)
        {
            uint64_t g = next();
            uint64_t lo = g * range, hi = __umul64hi(g, range);
            if (lo < range)
            {
                const uint64_t threshold = (0 - range) % range;
                while (lo < threshold)
                {
                    g = next();
                    lo = g * range;
                    hi = __umul64hi(g, range);
                }
            }
            return hi;
        }
    };

    // perm: per-frame scratch of n entries (PosT). n_err = floor(n * q), computed by the host in double as the reference does.
    template <typename PosT>
    __global__ void __launch_bounds__(128) generate_keys_kernel(long long n_frames, int n, int words, long long n_err, const uint64_t *__restrict__ seeds,
                                                                uint64_t seed_offset, PosT *__restrict__ perm_all, uint32_t *__restrict__ alice_out,
                                                                uint32_t *__restrict__ bob_out)
    {
        const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (f >= n_frames)
            return;
        Xoshiro256pp rng(seeds[f] + seed_offset);
        uint32_t *alice = alice_out + f * words, *bob = bob_out + f * words;

        // Alice: bit i = top bit of draw i
        for (int w = 0; w < words; ++w)
        {
            uint32_t word = 0;
            const int lim = min(32, n - w * 32);
            for (int b = 0; b < lim; ++b)
                word |= (uint32_t)(rng.next() >> 63) << b;
            alice[w] = word;
            bob[w] = word;
        }
        if (n_err == 0)
            return;

        // std::shuffle of 0..n-1 (forward Fisher-Yates: a[i] == i when position i is first visited)
        PosT *a = perm_all + (size_t)f * n;
        a[0] = 0;
        int i = 1;
        auto swap_in = [&](int pos, uint64_t j)
        {
            // swap(a[pos], a[j]) with a[pos] == pos still untouched
            if ((int)j == pos)
                a[pos] = (PosT)pos;
            else
            {
                a[pos] = a[j];
                a[j] = (PosT)pos;
            }
        };
        if (n > 1)
        {
            if ((n & 1) == 0)
            {
                swap_in(1, rng.below(2));
                i = 2;
            }
            for (; i < n; i += 2)
            {
                const uint64_t swap_range = (uint64_t)i + 1, b1 = swap_range + 1;
                const uint64_t x = rng.below(swap_range * b1);
                uint64_t j0, j1;
                if (swap_range * b1 <= 0xffffffffULL)
                {
                    j0 = (uint32_t)x / (uint32_t)b1;
                    j1 = (uint32_t)x % (uint32_t)b1;
                }
                else
                {
                    j0 = x / b1;
                    j1 = x % b1;
                }
                swap_in(i, j0);
                swap_in(i + 1, j1);
            }
        }
        // flip the first floor(n*q) positions of the shuffled order
        for (long long k = 0; k < n_err; ++k)
        {
            const uint32_t pos = (uint32_t)a[k];
            bob[pos >> 5] ^= 1u << (pos & 31);
        }
    }
}
