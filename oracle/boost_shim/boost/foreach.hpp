// Minimal stand-in for <boost/foreach.hpp> (oracle/_ref build only).
#pragma once
#include <iterator>
namespace boost_shim {
    template <typename C>
    struct reversed_range {
        C &c;
        auto begin() -> decltype(c.rbegin()) { return c.rbegin(); }
        auto end()   -> decltype(c.rend())   { return c.rend(); }
    };
    template <typename C>
    reversed_range<C> reversed(C &c) { return reversed_range<C>{c}; }
}
#define BOOST_FOREACH(decl, container)         for (decl : container)
#define BOOST_REVERSE_FOREACH(decl, container) for (decl : ::boost_shim::reversed(container))
