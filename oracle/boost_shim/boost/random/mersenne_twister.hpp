// Minimal stand-in (oracle/_ref build only). The parity runs never use random init:
// weights are always handed in explicitly, so stream equality with Boost is not needed.
#pragma once
#include <random>
namespace boost {
    typedef std::mt19937 mt19937;
    namespace random { typedef std::mt19937 mt19937; }
}
