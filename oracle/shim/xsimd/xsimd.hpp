// ORACLE - TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in for the part of the xsimd 14.3.0 API that the reference's CPU spreader uses
// (include/finufft/simd.hpp, spread.hpp, interp.hpp; the library itself is a network fetch of
// the reference's build, CMakeLists.txt:72, cmake/setupXSIMD.cmake, and is absent here).  It
// lets oracle/build.py compile the reference's OWN sources where they lie, so that the
// restatement in oracle/finufft_oracle.cpp and the GPU library can be checked against what the
// reference's code computes.  Written from the published xsimd interface, not from its sources:
// batches are GCC vector-extension values of 16 or 32 bytes ("sse2" / "avx2" architectures),
// lane semantics as documented by xsimd:
//   swizzle(x, mask)[i]    = x[mask[i]]
//   shuffle(x, y, mask)[i] = mask[i] < N ? x[mask[i]] : y[mask[i] - N]
//   fma(x, y, z) = x*y + z,  fnma(x, y, z) = -(x*y) + z   (fused)
//   to_int = truncation toward zero.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <new>
#include <type_traits>
#include <utility>

#include "config/xsimd_config.hpp"

namespace xsimd {

struct sse2 {
  static constexpr std::size_t bytes = 16;
  static constexpr std::size_t alignment() noexcept { return 16; }
  static constexpr const char *name() noexcept { return "sse2(stand-in)"; }
};
struct avx2 {
  static constexpr std::size_t bytes = 32;
  static constexpr std::size_t alignment() noexcept { return 32; }
  static constexpr const char *name() noexcept { return "avx2(stand-in)"; }
};
using best_arch    = avx2;
using default_arch = avx2;

template<class T> struct as_unsigned_integer {
  using type = std::make_unsigned_t<T>;
};
template<> struct as_unsigned_integer<float> {
  using type = uint32_t;
};
template<> struct as_unsigned_integer<double> {
  using type = uint64_t;
};
template<class T> using as_unsigned_integer_t = typename as_unsigned_integer<T>::type;

template<class T> struct as_integer {
  using type = T;
};
template<> struct as_integer<float> {
  using type = int32_t;
};
template<> struct as_integer<double> {
  using type = int64_t;
};

// GCC applies vector_size only to non-dependent types: spell the eight vector types out
template<class T, std::size_t Bytes> struct vec_of;
#define XSIMD_SHIM_VEC(T, B)                                                      \
  template<> struct vec_of<T, B> {                                                \
    typedef T type __attribute__((vector_size(B)));                               \
    typedef T utype __attribute__((vector_size(B), aligned(alignof(T))));         \
  };
XSIMD_SHIM_VEC(float, 16)
XSIMD_SHIM_VEC(float, 32)
XSIMD_SHIM_VEC(double, 16)
XSIMD_SHIM_VEC(double, 32)
XSIMD_SHIM_VEC(int32_t, 16)
XSIMD_SHIM_VEC(int32_t, 32)
XSIMD_SHIM_VEC(int64_t, 16)
XSIMD_SHIM_VEC(int64_t, 32)
#undef XSIMD_SHIM_VEC

template<class T, class A = default_arch> class batch {
 public:
  using value_type                  = T;
  using arch_type                   = A;
  static constexpr std::size_t size = A::bytes / sizeof(T);
  using vec_t = typename vec_of<T, A::bytes>::type;
  using vec_u = typename vec_of<T, A::bytes>::utype;
  vec_t v;

  batch() = default;
  batch(T s) noexcept {  // broadcast
    for (std::size_t i = 0; i < size; ++i) v[i] = s;
  }
  template<class... S, class = std::enable_if_t<sizeof...(S) + 2 == size>>
  batch(T a, T b, S... rest) noexcept : v{a, b, static_cast<T>(rest)...} {}
  explicit batch(vec_t x) noexcept : v(x) {}

  static batch load_aligned(const T *p) noexcept {
    return batch(*reinterpret_cast<const vec_t *>(p));
  }
  static batch load_unaligned(const T *p) noexcept {
    return batch((vec_t) * reinterpret_cast<const vec_u *>(p));
  }
  void store_aligned(T *p) const noexcept { *reinterpret_cast<vec_t *>(p) = v; }
  void store_unaligned(T *p) const noexcept { *reinterpret_cast<vec_u *>(p) = (vec_u)v; }
  T get(std::size_t i) const noexcept { return v[i]; }

  batch &operator+=(const batch &o) noexcept {
    v += o.v;
    return *this;
  }
  batch &operator-=(const batch &o) noexcept {
    v -= o.v;
    return *this;
  }
  batch &operator*=(const batch &o) noexcept {
    v *= o.v;
    return *this;
  }
};

#define XSIMD_SHIM_BINOP(op)                                                                     \
  template<class T, class A> inline batch<T, A> operator op(const batch<T, A> &a,               \
                                                            const batch<T, A> &b) noexcept {    \
    return batch<T, A>(a.v op b.v);                                                              \
  }                                                                                              \
  template<class T, class A, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>         \
  inline batch<T, A> operator op(const batch<T, A> &a, S s) noexcept {                           \
    return a op batch<T, A>(static_cast<T>(s));                                                  \
  }                                                                                              \
  template<class T, class A, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>         \
  inline batch<T, A> operator op(S s, const batch<T, A> &b) noexcept {                           \
    return batch<T, A>(static_cast<T>(s)) op b;                                                  \
  }
XSIMD_SHIM_BINOP(+)
XSIMD_SHIM_BINOP(-)
XSIMD_SHIM_BINOP(*)
XSIMD_SHIM_BINOP(/)
#undef XSIMD_SHIM_BINOP
template<class T, class A> inline batch<T, A> operator-(const batch<T, A> &a) noexcept {
  return batch<T, A>(-a.v);
}

// fused multiply-add per lane (hardware FMA under -mfma; std::fma is correctly rounded anyway)
template<class T, class A>
inline batch<T, A> fma(const batch<T, A> &x, const batch<T, A> &y, const batch<T, A> &z) noexcept {
  batch<T, A> r;
  for (std::size_t i = 0; i < batch<T, A>::size; ++i) r.v[i] = std::fma(x.v[i], y.v[i], z.v[i]);
  return r;
}
template<class T, class A>
inline batch<T, A> fnma(const batch<T, A> &x, const batch<T, A> &y, const batch<T, A> &z) noexcept {
  batch<T, A> r;
  for (std::size_t i = 0; i < batch<T, A>::size; ++i) r.v[i] = std::fma(-x.v[i], y.v[i], z.v[i]);
  return r;
}
template<class T, class A> inline batch<T, A> floor(const batch<T, A> &x) noexcept {
  batch<T, A> r;
  for (std::size_t i = 0; i < batch<T, A>::size; ++i) r.v[i] = std::floor(x.v[i]);
  return r;
}
template<class T, class A>
inline batch<T, A> min(const batch<T, A> &a, const batch<T, A> &b) noexcept {
  return batch<T, A>(a.v < b.v ? a.v : b.v);
}
template<class T, class A>
inline batch<T, A> max(const batch<T, A> &a, const batch<T, A> &b) noexcept {
  return batch<T, A>(a.v > b.v ? a.v : b.v);
}
template<class T, class A> inline T reduce_min(const batch<T, A> &a) noexcept {
  T r = a.v[0];
  for (std::size_t i = 1; i < batch<T, A>::size; ++i) r = a.v[i] < r ? a.v[i] : r;
  return r;
}
template<class T, class A> inline T reduce_max(const batch<T, A> &a) noexcept {
  T r = a.v[0];
  for (std::size_t i = 1; i < batch<T, A>::size; ++i) r = a.v[i] > r ? a.v[i] : r;
  return r;
}
template<class T, class A>
inline batch<typename as_integer<T>::type, A> to_int(const batch<T, A> &x) noexcept {
  using I = typename as_integer<T>::type;
  return batch<I, A>(__builtin_convertvector(x.v, typename batch<I, A>::vec_t));
}

// compile-time lane constants
template<class U, class A, U... Vs> struct batch_constant {
  using value_type                  = U;
  using arch_type                   = A;
  static constexpr std::size_t size = sizeof...(Vs);
};
namespace detail {
template<class U, class G, class A, std::size_t... I>
constexpr auto make_constant(std::index_sequence<I...>) noexcept {
  return batch_constant<U, A, static_cast<U>(G::get(static_cast<unsigned>(I),
                                                    static_cast<unsigned>(sizeof...(I))))...>{};
}
}  // namespace detail
template<class U, class G, class A = default_arch> constexpr auto make_batch_constant() noexcept {
  return detail::make_constant<U, G, A>(std::make_index_sequence<A::bytes / sizeof(U)>{});
}

template<class T, class A, class U, U... Vs>
inline batch<T, A> swizzle(const batch<T, A> &x, batch_constant<U, A, Vs...>) noexcept {
  static_assert(sizeof...(Vs) == batch<T, A>::size, "mask width");
  using I    = std::make_signed_t<U>;
  using ivec = typename vec_of<I, A::bytes>::type;
  return batch<T, A>(__builtin_shuffle(x.v, ivec{static_cast<I>(Vs)...}));
}
template<class T, class A, class U, U... Vs>
inline batch<T, A> shuffle(const batch<T, A> &x, const batch<T, A> &y,
                           batch_constant<U, A, Vs...>) noexcept {
  static_assert(sizeof...(Vs) == batch<T, A>::size, "mask width");
  using I    = std::make_signed_t<U>;
  using ivec = typename vec_of<I, A::bytes>::type;
  return batch<T, A>(__builtin_shuffle(x.v, y.v, ivec{static_cast<I>(Vs)...}));
}

// batch of exactly N lanes, or void when no architecture offers it
template<class T, std::size_t N> struct make_sized_batch {
  using type = std::conditional_t<
      N * sizeof(T) == sse2::bytes, batch<T, sse2>,
      std::conditional_t<N * sizeof(T) == avx2::bytes, batch<T, avx2>, void>>;
};
template<class T, std::size_t N> using make_sized_batch_t = typename make_sized_batch<T, N>::type;

template<class T, std::size_t Align> struct aligned_allocator {
  using value_type = T;
  template<class U> struct rebind {
    using other = aligned_allocator<U, Align>;
  };
  aligned_allocator() noexcept = default;
  template<class U> aligned_allocator(const aligned_allocator<U, Align> &) noexcept {}
  T *allocate(std::size_t n) {
    if (n == 0) return nullptr;
    const std::size_t bytes = (n * sizeof(T) + Align - 1) / Align * Align;
    void *p                 = std::aligned_alloc(Align, bytes);
    if (!p) throw std::bad_alloc();
    return static_cast<T *>(p);
  }
  void deallocate(T *p, std::size_t) noexcept { std::free(p); }
  template<class U> bool operator==(const aligned_allocator<U, Align> &) const noexcept {
    return true;
  }
  template<class U> bool operator!=(const aligned_allocator<U, Align> &) const noexcept {
    return false;
  }
};

}  // namespace xsimd
