// Minimal stand-in for <boost/noncopyable.hpp> (oracle/_ref build only).
#pragma once
namespace boost {
    class noncopyable {
    protected:
        noncopyable() {}
        ~noncopyable() {}
    private:
        noncopyable(const noncopyable&);
        noncopyable& operator=(const noncopyable&);
    };
}
