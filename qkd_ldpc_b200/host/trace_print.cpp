#include "trace_print.hpp"

#include <charconv>
#include <cmath>
#include <cstring>

namespace qkd_b200
{
    void print_coloured(colour c, const std::string &text, std::FILE *to)
    {
        static const char *const kCode[] = {"\x1b[38;2;000;000;255m", "\x1b[38;2;000;128;000m", "\x1b[38;2;128;000;128m", "\x1b[38;2;255;000;000m"};
        std::fputs(kCode[static_cast<int>(c)], to);
        std::fwrite(text.data(), 1, text.size(), to);
        std::fputs("\x1b[0m", to);
    }

    std::string format_g4(double v)
    {
        char buf[40];
        std::snprintf(buf, sizeof buf, "%.4g", v);
        return buf;
    }

    std::string format_shortest(double v)
    {
        if (std::isnan(v))
            return std::signbit(v) ? "-nan" : "nan";
        if (std::isinf(v))
            return v < 0 ? "-inf" : "inf";
        if (v == 0)
            return std::signbit(v) ? "-0" : "0";
        // shortest digits that round-trip, as d.ddd e±x
        char sci[48];
        const auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);
        std::string s(sci, r.ptr);
        std::string out;
        size_t pos = 0;
        if (s[0] == '-')
        {
            out = "-";
            pos = 1;
        }
        const size_t epos = s.find('e');
        std::string digits;
        for (size_t i = pos; i < epos; ++i)
            if (s[i] != '.')
                digits += s[i];
        const int exp10 = std::atoi(s.c_str() + epos + 1); // value = d.ddd * 10^exp10
        const int nd = static_cast<int>(digits.size());
        if (exp10 < -4 || exp10 >= 16)
        {
            out += digits[0];
            if (nd > 1)
                out += "." + digits.substr(1);
            char e[16];
            std::snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
            return out + e;
        }
        if (exp10 >= nd - 1) // integer: pad with zeros
            return out + digits + std::string(static_cast<size_t>(exp10 - (nd - 1)), '0');
        if (exp10 >= 0)
            return out + digits.substr(0, static_cast<size_t>(exp10) + 1) + "." + digits.substr(static_cast<size_t>(exp10) + 1);
        return out + "0." + std::string(static_cast<size_t>(-exp10 - 1), '0') + digits;
    }
}

void print_array(const int *const array, size_t array_length)
{
    for (size_t i = 0; i < array_length; i++)
        qkd_b200::print_coloured(qkd_b200::colour::blue, std::to_string(array[i]) + " ");
}

void print_array(const double *const array, size_t array_length)
{
    for (size_t i = 0; i < array_length; i++)
        qkd_b200::print_coloured(qkd_b200::colour::blue, qkd_b200::format_g4(array[i]) + " ");
}
