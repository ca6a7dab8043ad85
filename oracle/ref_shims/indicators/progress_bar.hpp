// ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// No-op stand-in for the un-vendored third-party module p-ranav/indicators v2.3
// (pinned at /root/reference/CMakeLists.txt:28-31). The reference only draws a
// console progress bar with it (/root/reference/src/simulation.cpp:194-215,239,248);
// nothing numeric depends on it.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace indicators
{
    enum class Color { grey, red, green, yellow, blue, magenta, cyan, white, unspecified };
    enum class FontStyle { bold, dark, italic, underline, blink, reverse, concealed, crossed };

    namespace option
    {
        template <typename T>
        struct Setting
        {
            T value{};
            Setting() = default;
            Setting(T v) : value(std::move(v)) {}
        };
        struct BarWidth : Setting<std::size_t> { using Setting::Setting; };
        struct Start : Setting<std::string> { using Setting::Setting; };
        struct Fill : Setting<std::string> { using Setting::Setting; };
        struct Lead : Setting<std::string> { using Setting::Setting; };
        struct Remainder : Setting<std::string> { using Setting::Setting; };
        struct End : Setting<std::string> { using Setting::Setting; };
        struct PrefixText : Setting<std::string> { using Setting::Setting; };
        struct PostfixText : Setting<std::string> { using Setting::Setting; };
        struct ForegroundColor : Setting<Color> { using Setting::Setting; };
        struct ShowElapsedTime : Setting<bool> { using Setting::Setting; };
        struct ShowRemainingTime : Setting<bool> { using Setting::Setting; };
        struct ShowPercentage : Setting<bool> { using Setting::Setting; };
        struct FontStyles : Setting<std::vector<FontStyle>> { using Setting::Setting; };
        struct MaxProgress : Setting<std::size_t> { using Setting::Setting; };
    }

    class ProgressBar
    {
    public:
        template <typename... Args>
        explicit ProgressBar(Args &&...) {}
        template <typename T>
        void set_option(T &&) {}
        void tick() {}
        void set_progress(std::size_t) {}
        void mark_as_completed() {}
        bool is_completed() const { return false; }
    };
}
