// xoshiro256++ seeded through SplitMix64, exposing the one type the reference's signatures name
// (`XoshiroCpp::Xoshiro256PlusPlus`, reference src/array_and_matrix_operations.hpp:29-53 / CMakeLists.txt:33-37:
// third-party module Reputeless/Xoshiro-cpp v1.1, not vendored there). Written from the published algorithm
// (Blackman & Vigna); a C++ UniformRandomBitGenerator, so <random> distributions and std::shuffle accept it.
#pragma once
#include <array>
#include <cstdint>
#include <limits>

namespace XoshiroCpp
{
    class Xoshiro256PlusPlus
    {
    public:
        using result_type = std::uint64_t;
        using state_type = std::array<std::uint64_t, 4>;

        explicit constexpr Xoshiro256PlusPlus(std::uint64_t seed = 0x2545F4914F6CDD1DULL) noexcept : s_{}
        {
            // SplitMix64 expands the seed into the four state words
            std::uint64_t x = seed;
            for (auto &w : s_)
            {
                std::uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
                w = z ^ (z >> 31);
            }
        }
        explicit constexpr Xoshiro256PlusPlus(state_type state) noexcept : s_(state) {}

        constexpr result_type operator()() noexcept
        {
            const std::uint64_t out = rotl(s_[0] + s_[3], 23) + s_[0];
            const std::uint64_t t = s_[1] << 17;
            s_[2] ^= s_[0];
            s_[3] ^= s_[1];
            s_[1] ^= s_[2];
            s_[0] ^= s_[3];
            s_[2] ^= t;
            s_[3] = rotl(s_[3], 45);
            return out;
        }
        static constexpr result_type min() noexcept { return 0; }
        static constexpr result_type max() noexcept { return std::numeric_limits<result_type>::max(); }
        constexpr state_type serialize() const noexcept { return s_; }

    private:
        static constexpr std::uint64_t rotl(std::uint64_t v, int k) noexcept { return (v << k) | (v >> (64 - k)); }
        state_type s_;
    };
}
