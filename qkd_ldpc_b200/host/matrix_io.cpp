// Parity-check matrix files -> H_matrix, and the key helpers that stay on the host.
// Mirrors (same signatures, validation and messages): reference src/array_and_matrix_operations.cpp
//   read_sparse_alist_matrix :109-292, read_dense_matrix :295-421, free_matrix_H :88-94, arrays_equal :96-106,
//   generate_random_bit_array :424-431, introduce_errors :434-460.
#include <algorithm>
#include <fstream>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>

#include "device_bridge.hpp"
#include "qkd_ldpc.hpp"

namespace
{
    using table = std::vector<std::vector<int>>;

    // Every line of the file as a row of integers (tokens that are not integers end the row, as `iss >> int` does).
    table read_int_rows(const fs::path &path, bool binary_only)
    {
        std::ifstream file(path);
        if (!file.is_open())
            throw std::runtime_error("Failed to open file: " + path.string());
        table rows;
        std::string line;
        while (std::getline(file, line))
        {
            std::istringstream iss(line);
            std::vector<int> row;
            for (int v; iss >> v;)
            {
                if (binary_only && v != 0 && v != 1)
                    throw std::runtime_error("Parity check matrix can only take values 0 or 1.");
                row.push_back(v);
            }
            rows.push_back(std::move(row));
        }
        if (rows.empty())
            throw std::runtime_error("File is empty or cannot be read properly: " + path.string());
        return rows;
    }

    void release_rows(int **rows, size_t count)
    {
        if (!rows)
            return;
        for (size_t i = 0; i < count; ++i)
            delete[] rows[i];
        delete[] rows;
    }

    bool all_equal(const int *w, size_t n)
    {
        return std::all_of(w, w + n, [&](int v) { return v == w[0]; });
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
    const table t = read_int_rows(matrix_path, false);
    const std::string where = matrix_path.string();
    if (t.size() < 4)
        throw std::runtime_error("Insufficient data in the file: " + where);
    if (t[0].size() != 2 || t[1].size() != 2)
        throw std::runtime_error("File format does not match the alist format: " + where);

    const size_t n = t[0][0], m = t[0][1];
    const size_t n_listed = t[2].size(), m_listed = t[3].size();
    if (t.size() < 4 + n_listed + m_listed)
        throw std::runtime_error("Insufficient data in the file: " + where);
    if (n != n_listed)
        throw std::runtime_error("Number of columns '" + std::to_string(n) + "' is not the same as the length of the third line '" +
                                 std::to_string(n_listed) + "'. File: " + where);
    if (m != m_listed)
        throw std::runtime_error("Number of rows '" + std::to_string(m) + "' is not the same as the length of the fourth line '" +
                                 std::to_string(m_listed) + "'. File: " + where);

    // the number of non-zero entries of every list must equal the declared weight
    auto check_weights = [&](size_t first_line, const std::vector<int> &weights, const char *which)
    {
        for (size_t i = 0; i < weights.size(); ++i)
        {
            const auto &row = t[first_line + i];
            const size_t nz = std::count_if(row.begin(), row.end(), [](int v) { return v != 0; });
            if (nz != static_cast<size_t>(weights[i]))
                throw std::runtime_error("Number of non-zero elements '" + std::to_string(nz) + "' in the line '" +
                                         std::to_string(first_line + i + 1) + "' does not match the weight in the " + which + " line '" +
                                         std::to_string(weights[i]) + "'. File: " + where);
        }
    };
    check_weights(4, t[2], "third");
    check_weights(4 + n, t[3], "fourth");

    H_matrix h;
    h.num_bit_nodes = n;
    h.num_check_nodes = m;
    h.max_bit_nodes_weight = t[1][0];
    h.max_check_nodes_weight = t[1][1];
    h.bit_nodes_weight = new int[n];
    h.check_nodes_weight = new int[m];
    std::copy(t[2].begin(), t[2].end(), h.bit_nodes_weight);
    std::copy(t[3].begin(), t[3].end(), h.check_nodes_weight);
    h.is_regular = all_equal(h.bit_nodes_weight, n) && all_equal(h.check_nodes_weight, m);

    // lists are 1-based in the file and read up to the node's weight (trailing zero padding is never looked at)
    auto fill = [&](size_t first_line, const int *weights, size_t count)
    {
        int **rows = new int *[count];
        for (size_t i = 0; i < count; ++i)
        {
            rows[i] = new int[weights[i]];
            for (int k = 0; k < weights[i]; ++k)
                rows[i][k] = t[first_line + i][k] - 1;
        }
        return rows;
    };
    h.bit_nodes = fill(4, h.bit_nodes_weight, n);
    h.check_nodes = fill(4 + n, h.check_nodes_weight, m);
    matrix_out = h;
}

void read_dense_matrix(const fs::path &matrix_path, H_matrix &matrix_out)
{
    const table t = read_int_rows(matrix_path, true);
    const std::string where = matrix_path.string();
    for (const auto &row : t)
        if (row.size() != t[0].size())
            throw std::runtime_error("Different lengths of rows in a matrix. File: " + where);

    const size_t n = t[0].size(), m = t.size();
    std::vector<int> col_w(n, 0), row_w(m, 0);
    for (size_t j = 0; j < m; ++j)
        for (size_t i = 0; i < n; ++i)
        {
            col_w[i] += t[j][i];
            row_w[j] += t[j][i];
        }
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
    h.max_bit_nodes_weight = *std::max_element(col_w.begin(), col_w.end());
    h.max_check_nodes_weight = *std::max_element(row_w.begin(), row_w.end());
    h.is_regular = all_equal(h.bit_nodes_weight, n) && all_equal(h.check_nodes_weight, m);

    // ascending neighbour lists in both directions
    h.bit_nodes = new int *[n];
    for (size_t i = 0; i < n; ++i)
    {
        h.bit_nodes[i] = new int[col_w[i]];
        int k = 0;
        for (size_t j = 0; j < m; ++j)
            if (t[j][i] == 1)
                h.bit_nodes[i][k++] = static_cast<int>(j);
    }
    h.check_nodes = new int *[m];
    for (size_t j = 0; j < m; ++j)
    {
        h.check_nodes[j] = new int[row_w[j]];
        int k = 0;
        for (size_t i = 0; i < n; ++i)
            if (t[j][i] == 1)
                h.check_nodes[j][k++] = static_cast<int>(i);
    }
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
