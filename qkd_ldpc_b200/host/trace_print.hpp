// Console helpers of the reference's src/utils.hpp:15-46 and the colour convention of its fmt::print(fg(...)) calls, without
// the fmt dependency: a coloured print is ESC[38;2;R;G;Bm <text> ESC[0m, "{:.4}" of a double is printf's "%.4g" (checked
// equal on 4 M values), "{}" of a double is its shortest round-trip form with fmt's fixed/exponent switch at 1e-4 / 1e16.
#pragma once
#include <cstddef>
#include <cstdio>
#include <string>

namespace qkd_b200
{
    enum class colour { blue, green, purple, red };
    void print_coloured(colour c, const std::string &text, std::FILE *to = stdout);
    std::string format_g4(double v);       // fmt "{:.4}"
    std::string format_shortest(double v); // fmt "{}"
}

void print_array(const int *const array, size_t array_length);    // "{} " per element, blue (src/utils.cpp:3-9)
void print_array(const double *const array, size_t array_length); // "{:.4} " per element, blue (src/utils.cpp:11-17)

// A matrix whose rows all have `cols_number` entries (src/utils.hpp:21-32): one row per line.
template <typename T>
void print_regular_matrix(const T *const *matrix, size_t rows_number, size_t cols_number)
{
    for (size_t i = 0; i < rows_number; i++)
    {
        print_array(matrix[i], cols_number);
        std::fputs("\n", stdout);
    }
}

// A matrix whose row i has rows_length[i] entries (src/utils.hpp:35-46).
template <typename T>
void print_irregular_matrix(const T *const *matrix, size_t rows_number, const int *const rows_length)
{
    for (size_t i = 0; i < rows_number; i++)
    {
        print_array(matrix[i], static_cast<size_t>(rows_length[i]));
        std::fputs("\n", stdout);
    }
}
