// Parity-check matrix files -> H_matrix (flat CSR / CSC arrays behind the reference's row pointers), and the key helpers that
// stay on the host.
// Mirrors (same signatures, validation and messages): reference src/array_and_matrix_operations.cpp
//   read_sparse_alist_matrix :109-292, read_dense_matrix :295-421, free_matrix_H :88-94, arrays_equal :96-106,
//   generate_random_bit_array :424-431, introduce_errors :434-460.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <set>
#include <stdexcept>

#include "device_bridge.hpp"
#include "qkd_ldpc.hpp"

namespace
{
    // ---- text scanning ------------------------------------------------------------------------------------------------------
    // The reference reads a file line by line and each line with `iss >> int` (src/array_and_matrix_operations.cpp:119-141):
    // integers separated by white space, the first token that is not an integer ends the row. Here the file is read once into
    // one buffer and scanned in place; rows are delivered value by value to a callback, so nothing is held per line -- the
    // adjacency lists go straight into two flat arrays (CSR for the checks, CSC for the bits).
    struct text_file
    {
        std::string data;
        size_t lines = 0; // as std::getline counts them: a final line without '\n' counts, an empty tail after the last '\n' does not
        explicit text_file(const fs::path &path)
        {
            std::ifstream file(path, std::ios::binary);
            if (!file.is_open())
                throw std::runtime_error("Failed to open file: " + path.string());
            file.seekg(0, std::ios::end);
            const std::streamoff size = file.tellg();
            file.seekg(0, std::ios::beg);
            data.resize(size > 0 ? static_cast<size_t>(size) : 0);
            if (!data.empty())
                file.read(data.data(), static_cast<std::streamsize>(data.size()));
            lines = static_cast<size_t>(std::count(data.begin(), data.end(), '\n')) + ((!data.empty() && data.back() != '\n') ? 1 : 0);
            if (lines == 0)
                throw std::runtime_error("File is empty or cannot be read properly: " + path.string());
        }
    };

    struct line_cursor
    {
        const char *p, *end;
        explicit line_cursor(const text_file &f) : p(f.data.data()), end(f.data.data() + f.data.size()) {}
        // Scans the next line, calling on_value(v) for every integer `iss >> int` would extract from it; returns how many.
        template <typename F>
        size_t next(F &&on_value)
        {
            const char *eol = static_cast<const char *>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
            const char *stop = eol ? eol : end;
            size_t count = 0;
            const char *q = p;
            for (;;)
            {
                while (q < stop && (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\v' || *q == '\f'))
                    ++q;
                if (q >= stop)
                    break;
                bool negative = false;
                if (*q == '+' || *q == '-')
                {
                    negative = *q == '-';
                    ++q;
                }
                if (q >= stop || *q < '0' || *q > '9')
                    break; // not an integer: the extraction fails and the row ends here
                long long v = 0;
                bool overflow = false;
                while (q < stop && *q >= '0' && *q <= '9')
                {
                    v = v * 10 + (*q - '0');
                    if (v > 2147483648LL)
                        overflow = true, v = 2147483648LL;
                    ++q;
                }
                if (negative)
                    v = -v;
                if (overflow || v > 2147483647LL || v < -2147483648LL)
                    break; // out of int's range: failbit, the row ends
                on_value(static_cast<int>(v));
                ++count;
            }
            p = eol ? eol + 1 : end;
            return count;
        }
    };

    // ---- storage --------------------------------------------------------------------------------------------------------------
    // H_matrix keeps the reference's shape (int ** rows), but the rows of a loaded matrix are slices of ONE array per direction:
    // check_nodes[j] = csr + row_ptr[j], bit_nodes[i] = csc + col_ptr[i]. device_bridge hands those arrays to qlb_code_create
    // without copying. The arenas are remembered here so that free_matrix_H knows a loaded matrix from one whose rows a caller
    // allocated one by one (the reference's own layout, still accepted everywhere).
    std::mutex g_arena_mu;
    std::set<const int *> g_arenas;

    int **slice_rows(int *arena, const int *weights, size_t count)
    {
        int **rows = new int *[count];
        size_t at = 0;
        for (size_t i = 0; i < count; ++i)
        {
            rows[i] = arena + at;
            at += static_cast<size_t>(weights[i]);
        }
        std::lock_guard<std::mutex> lk(g_arena_mu);
        g_arenas.insert(arena);
        return rows;
    }

    void release_rows(int **rows, size_t count)
    {
        if (!rows)
            return;
        bool arena = false;
        if (count > 0)
        {
            std::lock_guard<std::mutex> lk(g_arena_mu);
            arena = g_arenas.erase(rows[0]) > 0;
        }
        if (arena)
            delete[] rows[0];
        else
            for (size_t i = 0; i < count; ++i)
                delete[] rows[i];
        delete[] rows;
    }

    bool all_equal(const int *w, size_t n)
    {
        return std::all_of(w, w + n, [&](int v) { return v == w[0]; });
    }

    // one array for `count` lists of the given weights (at least one element, so that the arena has an address of its own)
    int *new_arena(const int *weights, size_t count)
    {
        size_t total = 0;
        for (size_t i = 0; i < count; ++i)
            total += static_cast<size_t>(std::max(weights[i], 0));
        return new int[std::max<size_t>(total, 1)];
    }
}

void free_matrix_H(H_matrix &matrix)
{
    qkd_b200::forget_matrix(matrix); // drop the device copy keyed on these arrays before they are recycled
    release_rows(matrix.bit_nodes, matrix.num_bit_nodes);
    release_rows(matrix.check_nodes, matrix.num_check_nodes);
    delete[] matrix.bit_nodes_weight;
    delete[] matrix.check_nodes_weight;
    matrix = H_matrix{};
}

bool arrays_equal(const int *const array1, const int *const array2, const size_t &array_length)
{
    return std::equal(array1, array1 + array_length, array2);
}

void read_sparse_alist_matrix(const fs::path &matrix_path, H_matrix &matrix_out)
{
    const text_file file(matrix_path);
    const std::string where = matrix_path.string();
    if (file.lines < 4)
        throw std::runtime_error("Insufficient data in the file: " + where);
    line_cursor cur(file);
    int dims[2] = {0, 0}, max_w[2] = {0, 0};
    size_t k = 0;
    const size_t dims_count = cur.next([&](int v) { if (k < 2) dims[k] = v; ++k; });
    k = 0;
    const size_t maxw_count = cur.next([&](int v) { if (k < 2) max_w[k] = v; ++k; });
    if (dims_count != 2 || maxw_count != 2)
        throw std::runtime_error("File format does not match the alist format: " + where);
    std::vector<int> bit_w, check_w; // third and fourth line
    cur.next([&](int v) { bit_w.push_back(v); });
    cur.next([&](int v) { check_w.push_back(v); });

    const size_t n = static_cast<size_t>(dims[0]), m = static_cast<size_t>(dims[1]);
    const size_t n_listed = bit_w.size(), m_listed = check_w.size();
    if (file.lines < 4 + n_listed + m_listed)
        throw std::runtime_error("Insufficient data in the file: " + where);
    if (n != n_listed)
        throw std::runtime_error("Number of columns '" + std::to_string(n) + "' is not the same as the length of the third line '" +
                                 std::to_string(n_listed) + "'. File: " + where);
    if (m != m_listed)
        throw std::runtime_error("Number of rows '" + std::to_string(m) + "' is not the same as the length of the fourth line '" +
                                 std::to_string(m_listed) + "'. File: " + where);

    // One pass over the list lines: the first `weight` entries of a line (1-based, trailing zero padding never looked at) go
    // straight to their place in the flat array, while the line's non-zero entries are counted against the declared weight.
    std::unique_ptr<int[]> csc(new_arena(bit_w.data(), n)), csr(new_arena(check_w.data(), m));
    auto read_lists = [&](const std::vector<int> &weights, int *arena, size_t first_line, const char *which)
    {
        size_t at = 0;
        for (size_t i = 0; i < weights.size(); ++i)
        {
            const int w = weights[i];
            size_t nz = 0;
            int taken = 0;
            cur.next([&](int v)
                     {
                         nz += v != 0;
                         if (taken < w)
                             arena[at + taken++] = v - 1;
                     });
            if (nz != static_cast<size_t>(w))
                throw std::runtime_error("Number of non-zero elements '" + std::to_string(nz) + "' in the line '" +
                                         std::to_string(first_line + i + 1) + "' does not match the weight in the " + which + " line '" +
                                         std::to_string(w) + "'. File: " + where);
            at += static_cast<size_t>(w);
        }
    };
    read_lists(bit_w, csc.get(), 4, "third");
    read_lists(check_w, csr.get(), 4 + n, "fourth");

    H_matrix h;
    h.num_bit_nodes = n;
    h.num_check_nodes = m;
    h.max_bit_nodes_weight = max_w[0];
    h.max_check_nodes_weight = max_w[1];
    h.bit_nodes_weight = new int[n];
    h.check_nodes_weight = new int[m];
    std::copy(bit_w.begin(), bit_w.end(), h.bit_nodes_weight);
    std::copy(check_w.begin(), check_w.end(), h.check_nodes_weight);
    h.is_regular = all_equal(h.bit_nodes_weight, n) && all_equal(h.check_nodes_weight, m);
    h.bit_nodes = slice_rows(csc.release(), h.bit_nodes_weight, n);
    h.check_nodes = slice_rows(csr.release(), h.check_nodes_weight, m);
    matrix_out = h;
}

void read_dense_matrix(const fs::path &matrix_path, H_matrix &matrix_out)
{
    const text_file file(matrix_path);
    const std::string where = matrix_path.string();
    // One pass: the positions of the ones of each row ARE the CSR half; column weights are counted on the way.
    std::vector<int> csr, row_w, col_w;
    csr.reserve(file.data.size() / 8);
    size_t n = 0;
    bool ragged = false;
    line_cursor cur(file);
    for (size_t j = 0; j < file.lines; ++j)
    {
        int column = 0, ones = 0;
        cur.next([&](int v)
                 {
                     if (v != 0 && v != 1)
                         throw std::runtime_error("Parity check matrix can only take values 0 or 1.");
                     if (v == 1)
                     {
                         csr.push_back(column);
                         ++ones;
                         if (static_cast<size_t>(column) >= col_w.size())
                             col_w.resize(static_cast<size_t>(column) + 1, 0);
                         ++col_w[static_cast<size_t>(column)];
                     }
                     ++column;
                 });
        if (j == 0)
            n = static_cast<size_t>(column);
        ragged |= static_cast<size_t>(column) != n;
        row_w.push_back(ones);
    }
    if (ragged)
        throw std::runtime_error("Different lengths of rows in a matrix. File: " + where);
    const size_t m = row_w.size();
    col_w.resize(n, 0);
    for (size_t i = 0; i < n; ++i)
        if (col_w[i] <= 0)
            throw std::runtime_error("Column '" + std::to_string(i + 1) + "' weight cannot be equal to or less than zero. File: " + where);
    for (size_t j = 0; j < m; ++j)
        if (row_w[j] <= 0)
            throw std::runtime_error("Row '" + std::to_string(j + 1) + "' weight cannot be equal to or less than zero. File: " + where);

    H_matrix h;
    h.num_bit_nodes = n;
    h.num_check_nodes = m;
    h.bit_nodes_weight = new int[n];
    h.check_nodes_weight = new int[m];
    std::copy(col_w.begin(), col_w.end(), h.bit_nodes_weight);
    std::copy(row_w.begin(), row_w.end(), h.check_nodes_weight);
    h.max_bit_nodes_weight = n ? *std::max_element(col_w.begin(), col_w.end()) : 0;
    h.max_check_nodes_weight = m ? *std::max_element(row_w.begin(), row_w.end()) : 0;
    h.is_regular = all_equal(h.bit_nodes_weight, n) && all_equal(h.check_nodes_weight, m);

    // CSC by a counting sort of the CSR entries: rows are visited in ascending order, so every bit's list comes out ascending
    int *csr_arena = new_arena(h.check_nodes_weight, m), *csc_arena = new_arena(h.bit_nodes_weight, n);
    std::copy(csr.begin(), csr.end(), csr_arena);
    std::vector<size_t> fill(n + 1, 0);
    for (size_t i = 0; i < n; ++i)
        fill[i + 1] = fill[i] + static_cast<size_t>(col_w[i]);
    size_t at = 0;
    for (size_t j = 0; j < m; ++j)
        for (int k2 = 0; k2 < row_w[j]; ++k2)
            csc_arena[fill[static_cast<size_t>(csr[at++])]++] = static_cast<int>(j);
    h.bit_nodes = slice_rows(csc_arena, h.bit_nodes_weight, n);
    h.check_nodes = slice_rows(csr_arena, h.check_nodes_weight, m);
    matrix_out = h;
}

// Alice's key: one draw per bit from uniform_int_distribution<int>(0, 1) -- the same library call as the reference, so the
// same libstdc++ produces the same keys from the same seed.
void generate_random_bit_array(XoshiroCpp::Xoshiro256PlusPlus &prng, size_t length, int *const random_bit_array_out)
{
    std::uniform_int_distribution<int> coin(0, 1);
    std::generate(random_bit_array_out, random_bit_array_out + length, [&] { return coin(prng); });
}

// Bob's key: Alice's with exactly floor(N * q) positions flipped, the positions being the head of a std::shuffle of 0..N-1.
double introduce_errors(XoshiroCpp::Xoshiro256PlusPlus &prng, const int *const bit_array, size_t array_length, double error_probability,
                        int *const bit_array_with_errors_out)
{
    const size_t flips = static_cast<size_t>(array_length * error_probability);
    std::copy(bit_array, bit_array + array_length, bit_array_with_errors_out);
    if (flips != 0)
    {
        std::vector<size_t> order(array_length);
        std::iota(order.begin(), order.end(), size_t{0});
        std::shuffle(order.begin(), order.end(), prng);
        for (size_t k = 0; k < flips; ++k)
            bit_array_with_errors_out[order[k]] ^= 1;
    }
    return static_cast<double>(flips) / array_length;
}
