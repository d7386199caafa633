#pragma once
#include <random>
namespace boost {
    namespace random { template <typename T = double> using normal_distribution = std::normal_distribution<T>; }
    template <typename T = double> using normal_distribution = std::normal_distribution<T>;
}
