// Basic types of the host layer: the reference's real_t / pattern types (Types.hpp:30-67) and the device-side
// vector that takes the place of thrust::device_vector, backed by the C ABI (bl_malloc / bl_memcpy_*).
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>
#include "blstm_b200.h"

#define PATTYPE_NONE   0
#define PATTYPE_FIRST  1
#define PATTYPE_NORMAL 2
#define PATTYPE_LAST   3

typedef float real_t;

namespace device {

// every non-zero status from the C ABI becomes the reference's error convention: std::runtime_error
inline void check(bl_ctx *ctx, int rc)
{
    if (rc) throw std::runtime_error(bl_last_error(ctx));
}

inline int paddedLd(int size) { return (size + 3) / 4 * 4; }     // rows padded to 16 bytes (TMA-legal strides)

template <typename T>
class Vector {
public:
    Vector() : m_ctx(nullptr), m_ptr(nullptr), m_size(0) {}
    Vector(bl_ctx *ctx, size_t n, bool zero = true) : m_ctx(nullptr), m_ptr(nullptr), m_size(0) { allocate(ctx, n, zero); }
    Vector(const Vector &) = delete;
    Vector &operator=(const Vector &) = delete;
    ~Vector() { release(); }

    void allocate(bl_ctx *ctx, size_t n, bool zero = true)
    {
        release();
        m_ctx = ctx; m_size = n;
        if (n) {
            check(ctx, bl_malloc(ctx, (void **)&m_ptr, n * sizeof(T)));
            if (zero) check(ctx, bl_memset(ctx, m_ptr, 0, n * sizeof(T)));
        }
    }
    void release() { if (m_ptr) bl_free(m_ctx, m_ptr); m_ptr = nullptr; m_size = 0; }
    void swap(Vector &o) { std::swap(m_ctx, o.m_ctx); std::swap(m_ptr, o.m_ptr); std::swap(m_size, o.m_size); }

    T *data() { return m_ptr; }
    const T *data() const { return m_ptr; }
    size_t size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    bl_ctx *ctx() const { return m_ctx; }

    void fromHost(const T *src, size_t n, size_t offset = 0)
    {
        if (offset + n > m_size) throw std::runtime_error("device::Vector::fromHost out of range");
        check(m_ctx, bl_memcpy_h2d(m_ctx, m_ptr + offset, src, n * sizeof(T)));
    }
    void toHost(T *dst, size_t n, size_t offset = 0) const
    {
        if (offset + n > m_size) throw std::runtime_error("device::Vector::toHost out of range");
        check(m_ctx, bl_memcpy_d2h(m_ctx, dst, m_ptr + offset, n * sizeof(T)));
        check(m_ctx, bl_sync(m_ctx));
    }
    std::vector<T> toHost() const { std::vector<T> v(m_size); if (m_size) toHost(v.data(), m_size); return v; }

private:
    bl_ctx *m_ctx;
    T *m_ptr;
    size_t m_size;
};

typedef Vector<real_t> real_vector;
typedef Vector<int>    int_vector;
typedef Vector<char>   pattype_vector;

} // namespace device
