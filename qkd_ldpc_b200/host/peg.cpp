// Seeded progressive-edge-growth (PEG) construction of column-weight-regular parity-check matrices, written as alist.
//
// The reference ships one code, `(N=10240,M=5231,R=0.49,CW=3,SEED=666).txt`, but not the program that made it
// (its .gitignore hides `alist_sparse_matrices_1/`). BASELINE.json's configs[3] (N = 100 000 and 1 000 000) and configs[4]
// (rates 0.3 ... 0.8 at N = 10 240) need codes "of the same construction": the shipped file has the signature of PEG
// (processing bits in order, every bit's first edge goes to the lowest-index minimum-degree check, all
// edges go to checks of degree <= min+1, row degrees differ by at most one, no 4-cycles), which is what this does
// (Hu, Eleftheriou, Arnold, "Regular and irregular progressive edge-growth Tanner graphs", 2005):
//   edge 0 of bit j  -> a check of minimum current degree (lowest index);
//   edge k > 0       -> grow the tree rooted at bit j through the graph built so far, level by level; when it stops
//                       reaching new checks, or is about to cover all of them, take a minimum-degree check among those
//                       farthest from / not reached by the tree (ties broken by the seeded generator).
// For large N the tree is cut off after `bfs_limit` checks (the girth guarantee is then local, which is all a
// depth-limited search can give); any check outside the partial tree qualifies.
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "XoshiroCpp.hpp"
#include "qkd_ldpc.hpp"

namespace qkd_b200
{
    void generate_peg_alist(size_t n, size_t m, size_t dv, uint64_t seed, size_t bfs_limit, const fs::path &out_path)
    {
        if (n == 0 || m == 0 || dv == 0 || dv > m)
            throw std::runtime_error("PEG: need n, m > 0 and 1 <= column weight <= m");
        XoshiroCpp::Xoshiro256PlusPlus rng(seed);
        std::vector<std::vector<uint32_t>> bit_adj(n), chk_adj(m);
        std::vector<uint32_t> degree(m, 0);
        std::vector<uint32_t> seen_epoch(m, 0), bit_epoch(n, 0);
        uint32_t epoch = 0;
        std::vector<uint32_t> frontier, next_frontier, last_level, candidates;

        auto pick_min_degree = [&](const std::vector<uint32_t> &pool, bool lowest_index)
        {
            uint32_t best_deg = UINT32_MAX;
            for (uint32_t c : pool)
                best_deg = std::min(best_deg, degree[c]);
            candidates.clear();
            for (uint32_t c : pool)
                if (degree[c] == best_deg)
                    candidates.push_back(c);
            if (lowest_index)
                return *std::min_element(candidates.begin(), candidates.end());
            return candidates[rng() % candidates.size()];
        };
        // minimum-degree check among those NOT marked in the current epoch (scan is short: the marked set is small or,
        // when it is large, every unmarked check qualifies)
        std::vector<uint32_t> all_checks(m);
        std::iota(all_checks.begin(), all_checks.end(), 0u);
        auto pick_unreached = [&]()
        {
            uint32_t best_deg = UINT32_MAX;
            candidates.clear();
            for (uint32_t c = 0; c < m; ++c)
            {
                if (seen_epoch[c] == epoch)
                    continue;
                if (degree[c] < best_deg)
                {
                    best_deg = degree[c];
                    candidates.clear();
                }
                if (degree[c] == best_deg)
                    candidates.push_back(c);
            }
            return candidates[rng() % candidates.size()];
        };

        // For big codes a full O(m) scan per edge is too slow: keep checks bucketed by degree and probe the lowest buckets.
        const bool big = m > 20000;
        std::vector<std::vector<uint32_t>> bucket; // bucket[d]: checks of degree d (lazy: entries may be stale)
        if (big)
        {
            bucket.resize(1);
            bucket[0] = all_checks;
        }
        auto pick_unreached_big = [&]()
        {
            for (size_t d = 0; d < bucket.size(); ++d)
            {
                auto &b = bucket[d];
                // drop stale entries from the back, then probe random positions
                for (int attempt = 0; attempt < 64 && !b.empty(); ++attempt)
                {
                    const size_t pos = rng() % b.size();
                    const uint32_t c = b[pos];
                    if (degree[c] != d)
                    {
                        b[pos] = b.back();
                        b.pop_back();
                        continue;
                    }
                    if (seen_epoch[c] != epoch)
                        return c;
                }
                for (size_t pos = 0; pos < b.size();)
                {
                    const uint32_t c = b[pos];
                    if (degree[c] != d)
                    {
                        b[pos] = b.back();
                        b.pop_back();
                        continue;
                    }
                    if (seen_epoch[c] != epoch)
                        return c;
                    ++pos;
                }
            }
            throw std::runtime_error("PEG: no unreached check left");
        };
        auto connect = [&](uint32_t bit, uint32_t chk)
        {
            bit_adj[bit].push_back(chk);
            chk_adj[chk].push_back(bit);
            ++degree[chk];
            if (big)
            {
                if (bucket.size() <= degree[chk])
                    bucket.resize(degree[chk] + 1);
                bucket[degree[chk]].push_back(chk);
            }
        };

        uint32_t min_degree_cursor = 0; // lowest-index check of minimum degree, advanced monotonically per degree round
        uint32_t current_min = 0;
        for (uint32_t j = 0; j < n; ++j)
        {
            for (size_t k = 0; k < dv; ++k)
            {
                uint32_t chosen;
                if (k == 0)
                {
                    // lowest-index check of minimum degree
                    for (;;)
                    {
                        while (min_degree_cursor < m && degree[min_degree_cursor] != current_min)
                            ++min_degree_cursor;
                        if (min_degree_cursor < m)
                            break;
                        min_degree_cursor = 0;
                        current_min = *std::min_element(degree.begin(), degree.end());
                    }
                    chosen = min_degree_cursor;
                }
                else
                {
                    ++epoch;
                    size_t reached = 0;
                    frontier.clear();
                    bit_epoch[j] = epoch;
                    for (uint32_t c : bit_adj[j])
                    {
                        seen_epoch[c] = epoch;
                        frontier.push_back(c);
                        ++reached;
                    }
                    last_level = frontier;
                    bool choose_from_last = false;
                    for (;;)
                    {
                        next_frontier.clear();
                        for (uint32_t c : frontier)
                            for (uint32_t b : chk_adj[c])
                            {
                                if (bit_epoch[b] == epoch)
                                    continue;
                                bit_epoch[b] = epoch;
                                for (uint32_t c2 : bit_adj[b])
                                    if (seen_epoch[c2] != epoch)
                                    {
                                        seen_epoch[c2] = epoch;
                                        next_frontier.push_back(c2);
                                    }
                            }
                        if (next_frontier.empty())
                            break; // the tree stopped growing: take an unreached check
                        if (reached + next_frontier.size() == m)
                        {
                            last_level = next_frontier; // about to cover everything: the farthest checks are the new ones
                            choose_from_last = true;
                            break;
                        }
                        reached += next_frontier.size();
                        frontier.swap(next_frontier);
                        if (reached > bfs_limit)
                            break; // depth-limited search (large codes)
                    }
                    if (choose_from_last)
                        chosen = pick_min_degree(last_level, false);
                    else
                        chosen = big ? pick_unreached_big() : pick_unreached();
                }
                connect(j, chosen);
            }
        }

        // alist: header, weights, 1-based sorted lists padded with zeros to the maximum weight
        size_t max_cw = 0;
        for (auto &row : chk_adj)
        {
            std::sort(row.begin(), row.end());
            max_cw = std::max(max_cw, row.size());
        }
        for (auto &col : bit_adj)
            std::sort(col.begin(), col.end());
        fs::create_directories(out_path.parent_path());
        std::ofstream out(out_path);
        if (!out.is_open())
            throw std::runtime_error("Failed to open file: " + out_path.string());
        std::string text;
        text.reserve(n * dv * 16);
        auto put = [&](size_t v)
        {
            text += std::to_string(v);
            text += ' ';
        };
        text += std::to_string(n) + " " + std::to_string(m) + "\n" + std::to_string(dv) + " " + std::to_string(max_cw) + "\n";
        for (size_t i = 0; i < n; ++i)
            put(bit_adj[i].size());
        text += "\n";
        for (size_t c = 0; c < m; ++c)
            put(chk_adj[c].size());
        text += "\n";
        for (size_t i = 0; i < n; ++i)
        {
            for (uint32_t c : bit_adj[i])
                put(c + 1);
            for (size_t pad = bit_adj[i].size(); pad < dv; ++pad)
                put(0);
            text += "\n";
        }
        for (size_t c = 0; c < m; ++c)
        {
            for (uint32_t b : chk_adj[c])
                put(b + 1);
            for (size_t pad = chk_adj[c].size(); pad < max_cw; ++pad)
                put(0);
            text += "\n";
        }
        out << text;
    }
}
