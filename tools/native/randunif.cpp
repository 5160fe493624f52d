// Synthetic inputs of the benches: the data definition of the reference's perftest generator
// (perftest/randunif.h:29-73, perftest/perftest.cpp:172-202), so that a number measured here is
// measured on the same points perftest / cuperftest would use:
//   value i of a stream = scale * (shift + u_i),  u_i the (i mod 65536)-th draw of
//   std::uniform_real_distribution<T>(-1, 1) on std::mt19937_64 seeded with stream + i / 65536.
// Built with the same libstdc++ (g++ 13), so the float distribution is bit-identical.
// C ABI, called from bench.py through ctypes; blocks are independent, so threads only change
// the wall time.
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

namespace {
constexpr int64_t kBlock = 1 << 16;

// values [i0, i0 + n) of the stream into out[0, n); blocks b0 + first, b0 + first + step, ...
template<class T>
void fill_blocks(T *out, int64_t i0, int64_t n, uint64_t stream, T scale, T shift, int64_t first,
                 int64_t step) {
  const int64_t b0 = i0 / kBlock, b1 = (i0 + n + kBlock - 1) / kBlock;
  for (int64_t b = b0 + first; b < b1; b += step) {
    std::mt19937_64 gen(stream + (uint64_t)b);
    std::uniform_real_distribution<T> u(T(-1), T(1));
    const int64_t end = std::min<int64_t>(i0 + n, (b + 1) * kBlock);
    for (int64_t i = b * kBlock; i < end; ++i) {
      const T v = scale * (shift + u(gen));
      if (i >= i0) out[i - i0] = v;
    }
  }
}

template<class T>
void fill(T *out, int64_t i0, int64_t n, uint64_t stream, T scale, T shift, int nthreads) {
  const int64_t nblocks = (i0 + n + kBlock - 1) / kBlock - i0 / kBlock;
  int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  if (nt > nblocks) nt = (int)nblocks;
  if (nt <= 1) {
    fill_blocks(out, i0, n, stream, scale, shift, 0, 1);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([=] { fill_blocks(out, i0, n, stream, scale, shift, t, nt); });
  for (auto &th : pool) th.join();
}
}  // namespace

extern "C" {
void b200_randunif_f32(float *out, int64_t first, int64_t n, uint64_t stream, float scale,
                       float shift, int nthreads) {
  fill<float>(out, first, n, stream, scale, shift, nthreads);
}
void b200_randunif_f64(double *out, int64_t first, int64_t n, uint64_t stream, double scale,
                       double shift, int nthreads) {
  fill<double>(out, first, n, stream, scale, shift, nthreads);
}
}
