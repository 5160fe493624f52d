// ORACLE - TEST INFRASTRUCTURE ONLY.
//
// Unnormalised in-place complex DFT with exponent sign `sign`, sizes 2,3,5-smooth (any size
// works, O(n*p) per prime factor p), OpenMP over blocks of lines.  Stands in for the
// reference's third-party FFT (FFTW 3.3.10 / ducc0, src/fft.cpp:109-123,266-371; dims
// slowest-first, :253-261) in two places: the CPU restatement oracle/finufft_oracle.cpp and
// the FFTW stand-in oracle/shim/fftw_standin.cpp that lets the reference's own sources link
// here.  Decimation-in-time mixed radix (4, 2, 3, 5, then any odd factor); kLanes lines are
// transformed side by side in split re/im form so the butterflies vectorise across lines.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

namespace orc {
using i64 = int64_t;
static constexpr double FFT_PI = 3.141592653589793238462643383279502884;
static constexpr int kLanes    = 16;

template<class T> struct Fft1d {
  i64 n;
  int sign;
  std::vector<T> twr, twi;  // exp(sign*2*pi*i*k/n)
  Fft1d(i64 n_, int sign_) : n(n_), sign(sign_), twr(n_), twi(n_) {
    for (i64 k = 0; k < n; ++k) {
      const double ang = sign * 2.0 * FFT_PI * (double)k / (double)n;
      twr[k] = (T)std::cos(ang);
      twi[k] = (T)std::sin(ang);
    }
  }
  // Arrays hold kLanes values per element: element e of lane b at [e*kLanes + b].
  // out[0..m) = DFT of in[0], in[stride], ... (length m, m | n); tmp = scratch of m elements.
  void rec(i64 m, const T *inr, const T *ini, i64 stride, T *outr, T *outi, T *tmpr,
           T *tmpi) const {
    constexpr int B = kLanes;
    if (m == 1) {
      for (int b = 0; b < B; ++b) outr[b] = inr[b], outi[b] = ini[b];
      return;
    }
    int p = 0;
    for (int cand : {4, 2, 3, 5})
      if (m % cand == 0) {
        p = cand;
        break;
      }
    if (!p)
      for (i64 cand = 7; cand <= m; cand += 2)
        if (m % cand == 0) {
          p = (int)cand;
          break;
        }
    const i64 q = m / p;
    for (int r = 0; r < p; ++r)
      rec(q, inr + r * stride * B, ini + r * stride * B, stride * p, tmpr + r * q * B,
          tmpi + r * q * B, outr + r * q * B, outi + r * q * B);
    const i64 tstep = n / m;
    if (p == 2) {
      for (i64 k = 0; k < q; ++k) {
        const T wr = twr[k * tstep], wi = twi[k * tstep];
        const T *ar = tmpr + k * B, *ai = tmpi + k * B, *br = tmpr + (q + k) * B,
                *bi = tmpi + (q + k) * B;
        T *o0r = outr + k * B, *o0i = outi + k * B, *o1r = outr + (k + q) * B,
          *o1i = outi + (k + q) * B;
        for (int b = 0; b < B; ++b) {
          const T xr = br[b] * wr - bi[b] * wi, xi = br[b] * wi + bi[b] * wr;
          o0r[b] = ar[b] + xr, o0i[b] = ai[b] + xi;
          o1r[b] = ar[b] - xr, o1i[b] = ai[b] - xi;
        }
      }
    } else if (p == 4) {
      const T sg = (T)sign;
      for (i64 k = 0; k < q; ++k) {
        const T w1r = twr[k * tstep], w1i = twi[k * tstep];
        const T w2r = twr[2 * k * tstep], w2i = twi[2 * k * tstep];
        const T w3r = twr[3 * k * tstep], w3i = twi[3 * k * tstep];
        const T *ar = tmpr + k * B, *ai = tmpi + k * B;
        const T *br = tmpr + (q + k) * B, *bi = tmpi + (q + k) * B;
        const T *cr = tmpr + (2 * q + k) * B, *ci = tmpi + (2 * q + k) * B;
        const T *dr = tmpr + (3 * q + k) * B, *di = tmpi + (3 * q + k) * B;
        T *o0r = outr + k * B, *o0i = outi + k * B;
        T *o1r = outr + (k + q) * B, *o1i = outi + (k + q) * B;
        T *o2r = outr + (k + 2 * q) * B, *o2i = outi + (k + 2 * q) * B;
        T *o3r = outr + (k + 3 * q) * B, *o3i = outi + (k + 3 * q) * B;
        for (int b = 0; b < B; ++b) {
          const T xbr = br[b] * w1r - bi[b] * w1i, xbi = br[b] * w1i + bi[b] * w1r;
          const T xcr = cr[b] * w2r - ci[b] * w2i, xci = cr[b] * w2i + ci[b] * w2r;
          const T xdr = dr[b] * w3r - di[b] * w3i, xdi = dr[b] * w3i + di[b] * w3r;
          const T s0r = ar[b] + xcr, s0i = ai[b] + xci, s1r = ar[b] - xcr, s1i = ai[b] - xci;
          const T s2r = xbr + xdr, s2i = xbi + xdi;
          // s3 = (sign*i) * (xb - xd)
          const T s3r = -sg * (xbi - xdi), s3i = sg * (xbr - xdr);
          o0r[b] = s0r + s2r, o0i[b] = s0i + s2i;
          o1r[b] = s1r + s3r, o1i[b] = s1i + s3i;
          o2r[b] = s0r - s2r, o2i[b] = s0i - s2i;
          o3r[b] = s1r - s3r, o3i[b] = s1i - s3i;
        }
      }
    } else {
      for (i64 k = 0; k < q; ++k)
        for (int j = 0; j < p; ++j) {
          const i64 kk = k + j * q;
          T *orr = outr + kk * B, *oi = outi + kk * B;
          for (int b = 0; b < B; ++b) orr[b] = tmpr[k * B + b], oi[b] = tmpi[k * B + b];
          for (int r = 1; r < p; ++r) {
            const i64 t = ((r * kk) % m) * tstep;
            const T wr = twr[t], wi = twi[t];
            const T *xr = tmpr + (r * q + k) * B, *xi = tmpi + (r * q + k) * B;
            for (int b = 0; b < B; ++b) {
              orr[b] += xr[b] * wr - xi[b] * wi;
              oi[b] += xr[b] * wi + xi[b] * wr;
            }
          }
        }
    }
  }
};

// dims fastest-first: nf[0] is the contiguous axis
template<class T> void fft_nd(int dim, const i64 *nf, int sign, T *data, int nthr) {
  constexpr int B = kLanes;
  auto *a        = reinterpret_cast<std::complex<T> *>(data);
  const i64 n[3] = {nf[0], dim > 1 ? nf[1] : 1, dim > 2 ? nf[2] : 1};
  if (nthr < 1) nthr = 1;
  for (int ax = 0; ax < dim; ++ax) {
    const i64 len = n[ax];
    if (len == 1) continue;
    Fft1d<T> plan(len, sign);
    // lines of this axis: element k of line (o, i) lives at o*len*inner + k*inner + i
    // (ax = 0: inner = 1 and lines are whole rows, so lanes are neighbouring rows instead)
    const i64 inner  = (ax == 0) ? 1 : (ax == 1 ? n[0] : n[0] * n[1]);
    const i64 outer  = n[0] * n[1] * n[2] / (len * inner);
    const i64 nlanes = (ax == 0) ? outer : inner;          // lines that can sit side by side
    const i64 ngroup = (ax == 0) ? 1 : outer;              // independent groups of such lines
    const i64 nblk   = (nlanes + B - 1) / B;
#pragma omp parallel num_threads(nthr)
    {
      std::vector<T> br(len * B), bi(len * B), orr(len * B), oi(len * B), tr(len * B),
          ti(len * B);
#pragma omp for schedule(static) collapse(2)
      for (i64 g = 0; g < ngroup; ++g)
        for (i64 blk = 0; blk < nblk; ++blk) {
          const i64 l0 = blk * B;
          const int nb = (int)((nlanes - l0) < B ? (nlanes - l0) : B);
          // lane b, element k
          auto at = [&](i64 k, int b) -> std::complex<T> & {
            return ax == 0 ? a[(l0 + b) * len + k] : a[g * len * inner + k * inner + l0 + b];
          };
          for (i64 k = 0; k < len; ++k) {
            for (int b = 0; b < nb; ++b) {
              const std::complex<T> v = at(k, b);
              br[k * B + b] = v.real(), bi[k * B + b] = v.imag();
            }
            for (int b = nb; b < B; ++b) br[k * B + b] = 0, bi[k * B + b] = 0;
          }
          plan.rec(len, br.data(), bi.data(), 1, orr.data(), oi.data(), tr.data(), ti.data());
          for (i64 k = 0; k < len; ++k)
            for (int b = 0; b < nb; ++b)
              at(k, b) = std::complex<T>(orr[k * B + b], oi[k * B + b]);
        }
    }
  }
}

}  // namespace orc
