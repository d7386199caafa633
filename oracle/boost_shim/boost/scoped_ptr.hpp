// Minimal stand-in for <boost/scoped_ptr.hpp> (oracle/_ref build only).
#pragma once
#include <cstddef>
namespace boost {
    template <typename T>
    class scoped_ptr {
        T *m_p;
        scoped_ptr(const scoped_ptr&);
        scoped_ptr& operator=(const scoped_ptr&);
    public:
        explicit scoped_ptr(T *p = NULL) : m_p(p) {}
        ~scoped_ptr() { delete m_p; }
        void reset(T *p = NULL) { if (p != m_p) { delete m_p; m_p = p; } }
        T* get() const { return m_p; }
        T& operator*() const { return *m_p; }
        T* operator->() const { return m_p; }
        operator bool() const { return m_p != NULL; }
    };
}
