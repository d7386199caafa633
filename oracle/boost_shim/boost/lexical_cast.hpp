// Minimal stand-in for <boost/lexical_cast.hpp> (oracle/_ref build only).
#pragma once
#include <sstream>
#include <string>
#include <stdexcept>
namespace boost {
    template <typename Target, typename Source>
    Target lexical_cast(const Source &src) {
        std::stringstream ss;
        ss << src;
        Target t;
        if (!(ss >> t)) throw std::runtime_error("bad lexical cast");
        return t;
    }
}
