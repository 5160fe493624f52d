// ORACLE - TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the three names of POET v0.0.1 (third-party compile-time dispatch library, a
// network fetch of the reference's build: cmake/setupPOET.cmake) that the reference's
// spreader uses (include/finufft/spread.hpp:951-954, interp.hpp:613-617):
//   poet::dispatch(functor, std::make_tuple(dispatch_param<inclusive_range<A0,A1>>{a},
//                                           dispatch_param<inclusive_range<B0,B1>>{b}))
// calls functor.template operator()<a, b>() for run-time a, b inside the ranges.  No arithmetic
// lives here.  Written from that call contract; values outside the ranges return R{}.
#pragma once
#include <tuple>
#include <utility>

namespace poet {

template<int Lo, int Hi> struct inclusive_range {
  static constexpr int lo = Lo, hi = Hi, count = Hi - Lo + 1;
};
template<class Range> struct dispatch_param {
  int value;
};

namespace detail {
template<class F, int A, int B> auto invoke(F &f) { return f.template operator()<A, B>(); }

template<class F, class R1, class R2, std::size_t... K>
auto dispatch2(F &f, int a, int b, std::index_sequence<K...>) {
  using Ret   = decltype(invoke<F, R1::lo, R2::lo>(f));
  using Fn    = Ret (*)(F &);
  static constexpr Fn table[] = {
      &invoke<F, R1::lo + (int)(K / R2::count), R2::lo + (int)(K % R2::count)>...};
  if (a < R1::lo || a > R1::hi || b < R2::lo || b > R2::hi) return Ret{};
  return table[(a - R1::lo) * R2::count + (b - R2::lo)](f);
}
}  // namespace detail

template<class F, class R1, class R2>
auto dispatch(F &&f, std::tuple<dispatch_param<R1>, dispatch_param<R2>> p) {
  return detail::dispatch2<std::remove_reference_t<F>, R1, R2>(
      f, std::get<0>(p).value, std::get<1>(p).value,
      std::make_index_sequence<(std::size_t)R1::count * R2::count>{});
}

}  // namespace poet
