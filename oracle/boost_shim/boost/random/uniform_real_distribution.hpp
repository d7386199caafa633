#pragma once
#include <random>
namespace boost {
    namespace random { template <typename T = double> using uniform_real_distribution = std::uniform_real_distribution<T>; }
}
