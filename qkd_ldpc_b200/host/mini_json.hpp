// A small JSON reader (objects, arrays, strings, numbers, booleans, null) -- enough for config.json. The reference
// parses its config with nlohmann/json (CMakeLists.txt:39-44), which is not vendored; this keeps the host build free of
// third-party code.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json
{
    struct value
    {
        enum kind_t { null_k, bool_k, number_k, string_k, array_k, object_k } kind = null_k;
        bool b = false;
        double num = 0;
        std::string text; // string payload, or the literal spelling of a number
        std::vector<value> items;
        std::map<std::string, value> members;

        bool contains(const std::string &key) const { return kind == object_k && members.count(key) != 0; }
        const value &at(const std::string &key) const
        {
            auto it = members.find(key);
            if (kind != object_k || it == members.end())
                throw std::runtime_error("key '" + key + "' not found");
            return it->second;
        }
        bool as_bool() const
        {
            if (kind != bool_k)
                throw std::runtime_error("type must be boolean");
            return b;
        }
        double as_double() const
        {
            if (kind != number_k)
                throw std::runtime_error("type must be number");
            return num;
        }
        size_t as_size() const
        {
            if (kind != number_k || num < 0)
                throw std::runtime_error("type must be a non-negative number");
            if (text.find_first_of(".eE") == std::string::npos)
                return static_cast<size_t>(std::strtoull(text.c_str(), nullptr, 10)); // exact for 64-bit seeds
            return static_cast<size_t>(num);
        }
        const std::string &as_string() const
        {
            if (kind != string_k)
                throw std::runtime_error("type must be string");
            return text;
        }
    };

    class parser
    {
    public:
        explicit parser(const std::string &s) : s_(s) {}
        value parse()
        {
            value v = element();
            skip();
            if (pos_ != s_.size())
                fail("trailing characters");
            return v;
        }

    private:
        const std::string &s_;
        size_t pos_ = 0;
        [[noreturn]] void fail(const std::string &what) const { throw std::runtime_error("JSON parse error at offset " + std::to_string(pos_) + ": " + what); }
        void skip()
        {
            // whitespace and a UTF-8 byte-order mark
            if (pos_ == 0 && s_.compare(0, 3, "\xEF\xBB\xBF") == 0)
                pos_ = 3;
            while (pos_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[pos_])))
                ++pos_;
        }
        bool eat(char c)
        {
            skip();
            if (pos_ < s_.size() && s_[pos_] == c)
            {
                ++pos_;
                return true;
            }
            return false;
        }
        value element()
        {
            skip();
            if (pos_ >= s_.size())
                fail("unexpected end");
            const char c = s_[pos_];
            if (c == '{')
                return object();
            if (c == '[')
                return array();
            if (c == '"')
            {
                value v;
                v.kind = value::string_k;
                v.text = string();
                return v;
            }
            if (s_.compare(pos_, 4, "true") == 0 || s_.compare(pos_, 5, "false") == 0)
            {
                value v;
                v.kind = value::bool_k;
                v.b = c == 't';
                pos_ += v.b ? 4 : 5;
                return v;
            }
            if (s_.compare(pos_, 4, "null") == 0)
            {
                pos_ += 4;
                return value{};
            }
            return number();
        }
        value number()
        {
            const size_t start = pos_;
            while (pos_ < s_.size() && (std::isdigit(static_cast<unsigned char>(s_[pos_])) || std::string("+-.eE").find(s_[pos_]) != std::string::npos))
                ++pos_;
            if (start == pos_)
                fail("unexpected character");
            value v;
            v.kind = value::number_k;
            v.text = s_.substr(start, pos_ - start);
            char *end = nullptr;
            v.num = std::strtod(v.text.c_str(), &end);
            if (end == v.text.c_str())
                fail("bad number");
            return v;
        }
        std::string string()
        {
            std::string out;
            ++pos_; // opening quote
            while (pos_ < s_.size() && s_[pos_] != '"')
            {
                char c = s_[pos_++];
                if (c == '\\' && pos_ < s_.size())
                {
                    const char e = s_[pos_++];
                    switch (e)
                    {
                    case 'n': c = '\n'; break;
                    case 't': c = '\t'; break;
                    case 'r': c = '\r'; break;
                    case 'b': c = '\b'; break;
                    case 'f': c = '\f'; break;
                    default: c = e; break; // \" \\ \/ (and \uXXXX left as-is: not needed for config.json)
                    }
                }
                out.push_back(c);
            }
            if (pos_ >= s_.size())
                fail("unterminated string");
            ++pos_;
            return out;
        }
        value array()
        {
            value v;
            v.kind = value::array_k;
            ++pos_;
            if (eat(']'))
                return v;
            do
                v.items.push_back(element());
            while (eat(','));
            if (!eat(']'))
                fail("expected ']'");
            return v;
        }
        value object()
        {
            value v;
            v.kind = value::object_k;
            ++pos_;
            if (eat('}'))
                return v;
            do
            {
                skip();
                if (pos_ >= s_.size() || s_[pos_] != '"')
                    fail("expected a member name");
                std::string key = string();
                if (!eat(':'))
                    fail("expected ':'");
                v.members[key] = element();
            } while (eat(','));
            if (!eat('}'))
                fail("expected '}'");
            return v;
        }
    };

    inline value parse(const std::string &text) { return parser(text).parse(); }
}
