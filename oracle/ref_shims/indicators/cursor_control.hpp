// ORACLE / TEST INFRASTRUCTURE ONLY -- no-op stand-in for p-ranav/indicators v2.3
// cursor control (/root/reference/src/simulation.cpp:202).
#pragma once
namespace indicators
{
    inline void show_console_cursor(bool) {}
}
