// Host-side preparation of a parity-check matrix for the decode kernels.
//
// The reference keeps H as two jagged adjacency halves (struct H_matrix,
// /root/reference/src/array_and_matrix_operations.hpp:16-27) and routes messages POSITIONALLY:
// a check appends its outgoing message to the next free slot of the receiving bit and vice versa
// (/root/reference/src/qkd_ldpc_algorithm.cpp:228-243, 300-311). The per-node operation order that
// results (product over a check's list left to right; sum over a bit's arrivals in check-scan order)
// is what has to be reproduced for bit-exact fp64 messages. CodeLayout replays those arrival counters
// once on the host and turns them into flat tables:
//
//   * physical message slots: checks are sorted by weight (descending, stable) and the k-th edge of the
//     check at sorted position p lives at slot base[k] + p. Consecutive threads (consecutive p) touch
//     consecutive addresses, so the check pass is bank-conflict free in shared memory and coalesced in
//     global memory, and no per-check offset table is needed: weight(p) = #{k : cnt[k] > p}.
//   * bit_slots[a][i]: slot holding the a-th message of bit i in the reference's summation order.
//
// One in-place message array is enough because a check reads and overwrites only its own slots and a
// bit reads and overwrites only the slots of its own edges.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

namespace qlb
{
    constexpr int kMaxCheckWeight = 128; // cnt[]/base[] travel in kernel parameters (2 x 512 B)
    constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

    struct CodeLayout
    {
        int32_t n = 0, m = 0, e = 0;
        int32_t max_check_w = 0, max_bit_w = 0;
        int32_t uniform_bit_w = 0; // common bit weight, 0 when bits differ
        int32_t slots = 0;         // physical message slots (>= e: every edge-position row starts on a 32-slot boundary)
        double gather_wavefronts_naive = 0, gather_wavefronts_opt = 0; // mean shared-memory wavefronts per bit-pass gather
        int32_t words_n = 0, words_m = 0;
        std::vector<int32_t> row_ptr, col_idx;   // check half as given (syndrome kernel)
        std::vector<uint32_t> cnt, base;         // [max_check_w]: checks with weight > k; slot offset of edge position k
        std::vector<uint32_t> check_order;       // [m] original check index at sorted position p
        std::vector<uint32_t> check_pos;         // [m] sorted position of original check j
        std::vector<uint32_t> slot_of_edge;      // [e] physical slot of CSR position
        std::vector<uint32_t> bit_slots;         // [max_bit_w * n], kNoSlot padded
        std::vector<uint32_t> col_of_slot;       // [e] bit index of the edge stored at a slot

        // Bank-aware placement. In the bit pass the 32 lanes of a warp (32 consecutive bits) gather one message each
        // through bit_slots[a][.]; with rows aligned to 32 slots the bank of a message is its check's sorted position mod
        // 32 (4-byte messages) and, for 8-byte messages -- which the hardware serves half-warp by half-warp over 16 pairs of
        // banks -- its position mod 16 within the half-warp. A random placement costs ~3.3 wavefronts per 4-byte gather (balls
        // in bins); re-ordering the checks inside each weight class so that the checks met by one warp at one list position
        // fall into distinct banks brings that close to 1. Greedy placement followed by pairwise swaps that reduce
        // sum(count^2) over (warp, position, bank mod 32) PLUS (warp, half-warp, position, bank mod 16): one order serves both
        // message widths (lanes of one half-warp take one of {b, b + 16} for every b, the other half the complements).
        //
        // Inside a weight class the checks are first cut into kLastBitBuckets runs by the bit index of their LAST edge (the
        // largest bit of the check for sorted adjacency lists), ascending, and the bank placement works inside a run. The fp64
        // SM-resident kernel keeps the end of the last row of message slots outside shared memory: with this order those
        // slots belong to bits of high index only, so that most 32-bit groups of its bit pass never leave shared memory.
        static constexpr int kLastBitBuckets = 16;
        void spread_banks(const int32_t *rp, const int32_t *ci, const int32_t *cp, const std::vector<int32_t> &edge_of_bitslot)
        {
            const int32_t warps = (n + 31) / 32;
            const int32_t groups = warps * max_bit_w;
            // Only codes that can live in shared memory profit (slots < 65535 is the resident kernels' limit); for larger codes
            // the natural order keeps the first edges of consecutive bits on consecutive message rows (DRAM locality).
            if (m < 64 || groups <= 0 || e >= 65535)
                return;
            // groups a check belongs to (one per edge): (warp of the bit, position of this edge in the bit's list)
            std::vector<std::vector<int32_t>> groups_of(m);
            for (int32_t i = 0; i < n; ++i)
                for (int32_t q = cp[i]; q < cp[i + 1]; ++q)
                {
                    const int32_t edge = edge_of_bitslot[q];
                    const int32_t j = static_cast<int32_t>(std::upper_bound(rp, rp + m + 1, edge) - rp) - 1;
                    groups_of[j].push_back(((i / 32) * max_bit_w + (q - cp[i])) * 2 + ((i % 32) / 16));
                }
            // histograms per group: [32 banks] for 4-byte messages, then [2 half-warps][16 bank pairs] for 8-byte messages
            std::vector<uint8_t> count(static_cast<size_t>(groups) * 64, 0);
            auto idx32 = [](int32_t gh, int b) { return static_cast<size_t>(gh >> 1) * 64 + b; };
            auto idx16 = [](int32_t gh, int b) { return static_cast<size_t>(gh >> 1) * 64 + 32 + (gh & 1) * 16 + (b & 15); };
            auto true_cost = [&](const std::vector<int32_t> &bank)
            {
                std::vector<uint8_t> c(static_cast<size_t>(groups) * 32, 0);
                for (int32_t j = 0; j < m; ++j)
                    for (int32_t g : groups_of[j])
                        ++c[static_cast<size_t>(g >> 1) * 32 + bank[j]];
                double tot = 0;
                int64_t used = 0;
                for (int32_t g = 0; g < groups; ++g)
                {
                    int mx = 0;
                    for (int b = 0; b < 32; ++b)
                        mx = std::max<int>(mx, c[static_cast<size_t>(g) * 32 + b]);
                    if (mx > 0)
                    {
                        tot += mx;
                        ++used;
                    }
                }
                return used ? tot / static_cast<double>(used) : 0.0;
            };
            // weight classes are contiguous runs of sorted positions
            std::vector<int32_t> bank(m, 0);
            {
                std::vector<int32_t> naive(m);
                for (int32_t p = 0; p < m; ++p)
                    naive[check_order[p]] = p % 32;
                gather_wavefronts_naive = true_cost(naive);
            }
            uint64_t rng = 0x9E3779B97F4A7C15ull;
            auto next = [&]()
            {
                rng ^= rng << 13;
                rng ^= rng >> 7;
                rng ^= rng << 17;
                return rng;
            };
            int32_t lo = 0;
            std::vector<uint32_t> new_order(m);
            while (lo < m)
            {
                // the next run: a kLastBitBuckets-th of a weight class (check_order is already sorted by (weight, last bit) here)
                int32_t hi = lo, class_hi = lo;
                const int32_t w = rp[check_order[lo] + 1] - rp[check_order[lo]];
                while (class_hi < m && rp[check_order[class_hi] + 1] - rp[check_order[class_hi]] == w)
                    ++class_hi;
                {
                    int32_t class_lo = lo;
                    while (class_lo > 0 && rp[check_order[class_lo - 1] + 1] - rp[check_order[class_lo - 1]] == w)
                        --class_lo;
                    const int32_t run = std::max<int32_t>(64, (class_hi - class_lo + kLastBitBuckets - 1) / kLastBitBuckets);
                    hi = std::min(class_hi, class_lo + ((lo - class_lo) / run + 1) * run);
                    if (class_hi - hi < 64)
                        hi = class_hi;
                }
                std::vector<int32_t> cap(32, 0);
                for (int32_t p = lo; p < hi; ++p)
                    ++cap[p % 32];
                std::vector<int32_t> members(check_order.begin() + lo, check_order.begin() + hi);
                // greedy
                for (int32_t j : members)
                {
                    int best = -1;
                    long best_cost = 0;
                    for (int b = 0; b < 32; ++b)
                    {
                        if (cap[b] == 0)
                            continue;
                        long c = 0;
                        for (int32_t g : groups_of[j])
                            c += count[idx32(g, b)] + count[idx16(g, b)];
                        if (best < 0 || c < best_cost)
                        {
                            best = b;
                            best_cost = c;
                        }
                    }
                    bank[j] = best;
                    --cap[best];
                    for (int32_t g : groups_of[j])
                    {
                        ++count[idx32(g, best)];
                        ++count[idx16(g, best)];
                    }
                }
                // pairwise swaps
                const size_t sz = members.size();
                if (sz >= 2)
                {
                    auto delta_move = [&](int32_t j, int from, int to)
                    {
                        long d = 0;
                        for (int32_t g : groups_of[j])
                        {
                            const long cf = count[idx32(g, from)], ct = count[idx32(g, to)];
                            d += (2 * ct + 1) - (2 * cf - 1); // (ct+1)^2 - ct^2 + (cf-1)^2 - cf^2
                            if (((from ^ to) & 15) != 0)
                            {
                                const long hf = count[idx16(g, from)], ht = count[idx16(g, to)];
                                d += (2 * ht + 1) - (2 * hf - 1);
                            }
                        }
                        return d;
                    };
                    auto apply_move = [&](int32_t j, int from, int to)
                    {
                        for (int32_t g : groups_of[j])
                        {
                            --count[idx32(g, from)];
                            ++count[idx32(g, to)];
                            --count[idx16(g, from)];
                            ++count[idx16(g, to)];
                        }
                        bank[j] = to;
                    };
                    const size_t trials = sz * 400;
                    for (size_t t = 0; t < trials; ++t)
                    {
                        const int32_t j1 = members[next() % sz], j2 = members[next() % sz];
                        const int b1 = bank[j1], b2 = bank[j2];
                        if (b1 == b2)
                            continue;
                        const long d1 = delta_move(j1, b1, b2);
                        apply_move(j1, b1, b2);
                        const long d2 = delta_move(j2, b2, b1);
                        if (d1 + d2 < 0)
                            apply_move(j2, b2, b1);
                        else
                            apply_move(j1, b2, b1); // undo
                    }
                }
                // positions lo..hi-1: position p takes the next check assigned to bank p % 32 (original order inside a bank)
                std::vector<std::vector<int32_t>> by_bank(32);
                std::vector<int32_t> sorted_members = members;
                std::sort(sorted_members.begin(), sorted_members.end());
                for (int32_t j : sorted_members)
                    by_bank[bank[j]].push_back(j);
                std::vector<size_t> taken(32, 0);
                for (int32_t p = lo; p < hi; ++p)
                    new_order[p] = static_cast<uint32_t>(by_bank[p % 32][taken[p % 32]++]);
                lo = hi;
            }
            check_order = new_order;
            gather_wavefronts_opt = true_cost(bank);
        }

        // Returns an empty string on success, else the reason the matrix is rejected.
        std::string build(int32_t n_bits, int32_t n_checks, const int32_t *rp, const int32_t *ci, const int32_t *cp,
                          const int32_t *ri)
        {
            if (n_bits <= 0 || n_checks <= 0 || !rp || !ci || !cp || !ri)
                return "matrix dimensions must be positive and all four index arrays non-null";
            n = n_bits;
            m = n_checks;
            if (rp[0] != 0 || cp[0] != 0)
                return "row_ptr[0] and col_ptr[0] must be 0";
            for (int32_t j = 0; j < m; ++j)
                if (rp[j + 1] < rp[j])
                    return "row_ptr is not non-decreasing";
            for (int32_t i = 0; i < n; ++i)
                if (cp[i + 1] < cp[i])
                    return "col_ptr is not non-decreasing";
            if (rp[m] != cp[n])
                return "the two adjacency halves hold different numbers of edges (" + std::to_string(rp[m]) + " vs " +
                       std::to_string(cp[n]) + ")";
            e = rp[m];
            if (e <= 0)
                return "matrix has no edges";
            for (int32_t p = 0; p < e; ++p)
            {
                if (ci[p] < 0 || ci[p] >= n)
                    return "bit index out of range in the check half";
                if (ri[p] < 0 || ri[p] >= m)
                    return "check index out of range in the bit half";
            }

            // Replay the reference's arrival counters (qkd_ldpc_algorithm.cpp:228-243): scanning the check half in
            // order, the message for bit b lands in that bit's next free slot.
            std::vector<int32_t> arrivals(n, 0);
            std::vector<int32_t> edge_of_bitslot(e, -1); // bit-side slot -> check-side edge that fills it
            for (int32_t j = 0; j < m; ++j)
                for (int32_t p = rp[j]; p < rp[j + 1]; ++p)
                {
                    const int32_t b = ci[p];
                    const int32_t w = cp[b + 1] - cp[b];
                    if (arrivals[b] >= w)
                        return "bit " + std::to_string(b) + " receives more messages than its weight " + std::to_string(w) +
                               " (the reference would write past the end of its row)";
                    edge_of_bitslot[cp[b] + arrivals[b]++] = p;
                }
            for (int32_t i = 0; i < n; ++i)
                if (arrivals[i] != cp[i + 1] - cp[i])
                    return "bit " + std::to_string(i) + " receives fewer messages than its weight (the reference would read "
                           "uninitialised memory)";
            // ... and the reverse direction (:300-311): the message bit i derives from its a-th arrival goes to check
            // bit_nodes[i][a] and lands in that check's next free slot. The in-place message array requires it to land
            // exactly on the edge it came from, i.e. both halves list the same edges in a mutually consistent order
            // (true for the reference's dense loader by construction and for sorted alist files).
            std::vector<int32_t> carrivals(m, 0);
            for (int32_t i = 0; i < n; ++i)
                for (int32_t q = cp[i]; q < cp[i + 1]; ++q)
                {
                    const int32_t c = ri[q];
                    const int32_t w = rp[c + 1] - rp[c];
                    if (carrivals[c] >= w)
                        return "check " + std::to_string(c) + " receives more messages than its weight";
                    const int32_t landing = rp[c] + carrivals[c]++;
                    if (landing != edge_of_bitslot[q])
                        return "the bit half and the check half of the matrix are not mutually consistent at bit " +
                               std::to_string(i) + " (unsorted adjacency lists or different edge sets): the reference's "
                               "positional routing would misroute messages on this input";
                }

            row_ptr.assign(rp, rp + m + 1);
            col_idx.assign(ci, ci + e);
            words_n = (n + 31) / 32;
            words_m = (m + 31) / 32;
            max_check_w = max_bit_w = 0;
            for (int32_t j = 0; j < m; ++j)
                max_check_w = std::max(max_check_w, rp[j + 1] - rp[j]);
            for (int32_t i = 0; i < n; ++i)
                max_bit_w = std::max(max_bit_w, cp[i + 1] - cp[i]);
            uniform_bit_w = max_bit_w;
            for (int32_t i = 0; i < n; ++i)
                if (cp[i + 1] - cp[i] != max_bit_w)
                    uniform_bit_w = 0;
            if (max_check_w > kMaxCheckWeight)
                return "check weight " + std::to_string(max_check_w) + " exceeds the supported maximum " +
                       std::to_string(kMaxCheckWeight);

            check_order.resize(m);
            std::iota(check_order.begin(), check_order.end(), 0u);
            // (large codes keep the file order inside a weight class: the streaming decoder wants the first edges of consecutive
            // bits on consecutive message rows, and no SM-resident kernel takes them)
            const bool by_last_bit = e < 65535;
            auto last_bit = [&](uint32_t j) { return by_last_bit && rp[j + 1] > rp[j] ? ci[rp[j + 1] - 1] : 0; };
            std::stable_sort(check_order.begin(), check_order.end(), [&](uint32_t a, uint32_t b)
                             {
                                 const int32_t wa = rp[a + 1] - rp[a], wb = rp[b + 1] - rp[b];
                                 return wa != wb ? wa > wb : last_bit(a) < last_bit(b);
                             });
            spread_banks(rp, ci, cp, edge_of_bitslot);
            check_pos.resize(m);
            for (int32_t p = 0; p < m; ++p)
                check_pos[check_order[p]] = static_cast<uint32_t>(p);
            cnt.assign(max_check_w, 0);
            base.assign(max_check_w, 0);
            for (int32_t j = 0; j < m; ++j)
                for (int32_t k = 0; k < rp[j + 1] - rp[j]; ++k)
                    ++cnt[k];
            // every row of slots starts on a multiple of 32, so the shared-memory bank of a slot is (sorted position % 32)
            for (int32_t k = 1; k < max_check_w; ++k)
                base[k] = base[k - 1] + (cnt[k - 1] + 31u) / 32u * 32u;
            slots = static_cast<int32_t>(base[max_check_w - 1] + cnt[max_check_w - 1]);

            slot_of_edge.resize(e);
            col_of_slot.assign(slots, kNoSlot);
            for (int32_t j = 0; j < m; ++j)
                for (int32_t p = rp[j]; p < rp[j + 1]; ++p)
                {
                    const uint32_t s = base[p - rp[j]] + check_pos[j];
                    slot_of_edge[p] = s;
                    col_of_slot[s] = static_cast<uint32_t>(ci[p]);
                }
            bit_slots.assign(static_cast<size_t>(max_bit_w) * n, kNoSlot);
            for (int32_t i = 0; i < n; ++i)
                for (int32_t q = cp[i]; q < cp[i + 1]; ++q)
                    bit_slots[static_cast<size_t>(q - cp[i]) * n + i] = slot_of_edge[edge_of_bitslot[q]];
            return std::string();
        }
    };
}
