// ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Stand-in for the un-vendored third-party module Reputeless/Xoshiro-cpp v1.1
// (pinned at /root/reference/CMakeLists.txt:33-37), written from the published
// xoshiro256++ / SplitMix64 algorithms (Blackman & Vigna, public domain).
// Only the surface the reference uses is provided: a C++ UniformRandomBitGenerator
// `XoshiroCpp::Xoshiro256PlusPlus(uint64 seed)` (call sites:
// /root/reference/src/simulation.cpp:95,163,222 and
// /root/reference/src/array_and_matrix_operations.cpp:424,434).
//
// Known answers checked in tests/test_oracle_prng.py: state {1,2,3,4} ->
// 41943041, 58720359, 3588806011781223; seed 777 -> 2066146677187504009, ...
#pragma once
#include <array>
#include <cstdint>
#include <limits>

namespace XoshiroCpp
{
    class SplitMix64
    {
    public:
        using result_type = std::uint64_t;
        explicit constexpr SplitMix64(std::uint64_t seed = 0) noexcept : m_state(seed) {}
        constexpr result_type operator()() noexcept
        {
            std::uint64_t z = (m_state += 0x9e3779b97f4a7c15ULL);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            return z ^ (z >> 31);
        }
        static constexpr result_type min() noexcept { return 0; }
        static constexpr result_type max() noexcept { return std::numeric_limits<result_type>::max(); }

    private:
        std::uint64_t m_state;
    };

    class Xoshiro256PlusPlus
    {
    public:
        using state_type = std::array<std::uint64_t, 4>;
        using result_type = std::uint64_t;

        explicit constexpr Xoshiro256PlusPlus(std::uint64_t seed = 0x2545F4914F6CDD1DULL) noexcept : m_state{}
        {
            SplitMix64 sm(seed);
            for (auto &s : m_state)
            {
                s = sm();
            }
        }
        explicit constexpr Xoshiro256PlusPlus(state_type state) noexcept : m_state(state) {}

        constexpr result_type operator()() noexcept
        {
            const std::uint64_t result = rotl(m_state[0] + m_state[3], 23) + m_state[0];
            const std::uint64_t t = m_state[1] << 17;
            m_state[2] ^= m_state[0];
            m_state[3] ^= m_state[1];
            m_state[1] ^= m_state[2];
            m_state[0] ^= m_state[3];
            m_state[2] ^= t;
            m_state[3] = rotl(m_state[3], 45);
            return result;
        }
        static constexpr result_type min() noexcept { return 0; }
        static constexpr result_type max() noexcept { return std::numeric_limits<result_type>::max(); }

    private:
        static constexpr std::uint64_t rotl(std::uint64_t x, int s) noexcept { return (x << s) | (x >> (64 - s)); }
        state_type m_state;
    };
}
