// Row-sweep spread / interp kernels for 3D float, "quad split" register layout (sweep3d.cuh).
//
// Same sweep as sweep3d.cu (a warp per row of bins, 8-column window in the lanes, staging tile,
// vector reductions), different ownership inside a window column: lane (la, by, bz) holds the
// tile rows y = by + 2m, z = bz + 2n (m, n < 6): 36 complex accumulators.  A point whose y / z
// stencils start at jb / k0 touches, in every lane, the slots m in [jb/2, jb/2+4) and
// n in [k0/2, k0/2+4): 16 packed FFMA2 per point and lane instead of 21 (x 7/8, y 7/8, z 7/8 of
// the lanes / slots carry non-zero weights; the 4-row z split of sweep3d.cu leaves z at 7/12).
// Runs are keyed by (window position, jb/2, k0/2): 9 unrolled bodies.  The weights of a point
// are stored per lane parity (wy[by][4], wz[bz][4], zero where the slot lies outside the
// stencil) and the x window comes pre-multiplied by the strength (spread).
#include "sweep3d.cuh"

#include "sweepmath.cuh"

#include <limits.h>
#include <stdlib.h>

#include <type_traits>

namespace b200 {

namespace {

template<int OFF> __device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "f"(v.x), "f"(v.y)
               : "memory");
}
template<int OFF> __device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(OFF));
  return v;
}

template<int NS> struct QCfg {
  static constexpr int HL   = NS / 2;        // cells left of a bin a stencil can reach
  static constexpr int W    = 8;             // x columns of the register window
  static constexpr int YR   = kBinY + NS;    // y rows of the tile
  static constexpr int ZT   = kBinZ + NS;    // z rows of the tile
  static constexpr int MS   = 6;             // slots per lane and direction: row = parity + 2*slot
  static constexpr int NSL  = NS / 2 + 1;    // slots a point touches per direction
  static constexpr int NCL  = 3;             // slot offsets jb/2, k0/2 in 0..2
  static constexpr int XB   = 4;             // window positions are (i0 + XB) >> 1
  static constexpr int SX   = 4;             // x columns collected per flush / fill
  static constexpr int NRED = (ZT * YR * (SX / 2) + 31) / 32;  // 16-byte pieces per lane
  static constexpr int CH   = 32;            // points per chunk: one per lane
  static constexpr int RECW = 36;            // words per point record
  // record: [4*by + d] y weight of slot jb/2 + d for lanes of parity by   [8 + 4*bz + d] same in z
  //         spread: [16 + 2a, +1] strength times the x weight of window column a (mod 8)
  //         interp: [16 + a] x weight of window column a
  static constexpr size_t STAGE_BYTES = (size_t)NRED * 32 * sizeof(float4);
  static constexpr size_t REC_BYTES   = (size_t)(CH + 2) * RECW * sizeof(float);
  static constexpr int PP             = 17;  // pitch of a point's 16 partial sums (interp)
  static constexpr size_t PART_BYTES  = (size_t)CH * PP * sizeof(float2);
  static_assert(NS + 1 <= W && HL <= XB && NS <= 7 && NSL == 4, "layout");
  static_assert(NCL - 1 + NSL <= MS && YR <= 2 * MS && ZT <= 2 * MS, "slots");
};

// Horner table padded to 8 columns (pairs of panels), as in sweep3d.cu
template<int NS> struct alignas(16) PairTable8 {
  float c[TableRows<NS>::value * 8];
};
template<int NS>
__device__ __forceinline__ void eval_window8(const PairTable8<NS> &tab, float x1, float (&out)[8]) {
  const float z = fma_rn(2.0f, x1, (float)(NS - 1));
  float2 r[4];
#pragma unroll
  for (int p = 0; p < (NS + 1) / 2; ++p) {
    r[p] = *reinterpret_cast<const float2 *>(&tab.c[2 * p]);
#pragma unroll
    for (int k = 1; k < TableRows<NS>::value; ++k)
      r[p] = ffma2_s(z, r[p], *reinterpret_cast<const float2 *>(&tab.c[k * 8 + 2 * p]));
    out[2 * p]     = r[p].x;
    out[2 * p + 1] = r[p].y;
  }
}

template<int NS> struct QArgs {
  const float *xs, *ys, *zs;   // coordinates, refined bin order
  const uint32_t *sidx;        // position -> user index
  const SweepItem *items;
  GridGeom<float> g;
  PairTable8<NS> tab;
  const float2 *c_in;
  float2 *c_out;
  float2 *fw;
};

struct RawPoint {
  float x, y, z;
  uint32_t j;
};

// weight of stencil cell t (0 outside the stencil)
template<int NS, int T> __device__ __forceinline__ float wsel(const float (&kv)[8]) {
  if constexpr (T >= 0 && T < NS) return kv[T];
  else return 0.f;
}
// the four slot weights of both lane parities for a stencil that starts at row `start`
template<int NS>
__device__ __forceinline__ void slot_weights(const float (&kv)[8], int start, float *dst) {
  const bool odd = start & 1;
  float w[8];
  static_for<0, 2>([&](auto pc) {
    constexpr int par = decltype(pc)::value;
    static_for<0, 4>([&](auto dc) {
      constexpr int d = decltype(dc)::value;
      // row = par + 2*(start/2 + d); stencil cell t = row - start = par + 2d - (start & 1)
      w[4 * par + d] = odd ? wsel<NS, par + 2 * d - 1>(kv) : wsel<NS, par + 2 * d>(kv);
    });
  });
  *reinterpret_cast<float4 *>(dst)     = make_float4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<float4 *>(dst + 4) = make_float4(w[4], w[5], w[6], w[7]);
}

// Thread-per-point preparation: fold, stencil starts, windows (+ strength) -> record; returns
// the run key (window position << 4 | 3*(jb/2) + k0/2).
template<int NS, bool SPREAD>
__device__ __forceinline__ int make_record_q(const QArgs<NS> &a, const RawPoint &pt, float2 c,
                                             float *rec, int i2, int i3) {
  using CF = QCfg<NS>;
  int i0;
  float x1;
  float kv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  stencil_start<float, NS>(fold_rescale<float>(pt.x, a.g.nf_t[0]), i0, x1);
  eval_window8<NS>(a.tab, x1, kv);
  const int gpos = (i0 + CF::XB) >> 1;  // window start 2*gpos - XB <= i0 <= that + 1
  // rotate: the weight of stencil cell t belongs to the window column x = i0 + t (mod 8)
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float w = t < NS ? kv[t] : 0.f;
    const int col = (i0 + t) & 7;
    if constexpr (SPREAD) *reinterpret_cast<float2 *>(rec + 16 + 2 * col) = fmul2_s(w, c);
    else rec[16 + col] = w;
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) kv[s] = 0.f;
  stencil_start<float, NS>(fold_rescale<float>(pt.y, a.g.nf_t[1]), i0, x1);
  eval_window8<NS>(a.tab, x1, kv);
  const int jb = min(max(i0 - (kBinY * i2 - CF::HL), 0), kBinY);
  slot_weights<NS>(kv, jb, rec);
#pragma unroll
  for (int s = 0; s < 8; ++s) kv[s] = 0.f;
  stencil_start<float, NS>(fold_rescale<float>(pt.z, a.g.nf_t[2]), i0, x1);
  eval_window8<NS>(a.tab, x1, kv);
  const int k0 = min(max(i0 - (kBinZ * i3 - CF::HL), 0), kBinZ);
  slot_weights<NS>(kv, k0, rec + 8);
  return (gpos << 4) | (3 * (jb >> 1) + (k0 >> 1));
}

// what one lane needs of one record
struct LaneRecQ {
  float4 wy, wz;
  float2 cw;  // spread: strength * x weight;  interp: .x = x weight
};

template<int NS, int JC, int KC>
__device__ __forceinline__ void spread_update_q(float2 (&acc)[QCfg<NS>::MS][QCfg<NS>::MS],
                                                const LaneRecQ &r) {
  const float wy[4] = {r.wy.x, r.wy.y, r.wy.z, r.wy.w};
  const float wz[4] = {r.wz.x, r.wz.y, r.wz.z, r.wz.w};
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const float2 w = fmul2_s(wz[n], r.cw);
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[KC + n][JC + m] = ffma2_s(wy[m], w, acc[KC + n][JC + m]);
  }
}
template<int NS, int JC, int KC>
__device__ __forceinline__ float2 interp_gather_q(
    const float2 (&gv)[QCfg<NS>::MS][QCfg<NS>::MS], const LaneRecQ &r) {
  const float wy[4] = {r.wy.x, r.wy.y, r.wy.z, r.wy.w};
  const float wz[4] = {r.wz.x, r.wz.y, r.wz.z, r.wz.w};
  float2 tot = make_float2(0.f, 0.f);
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    float2 s = fmul2_s(wy[0], gv[KC + n][JC]);
#pragma unroll
    for (int m = 1; m < 4; ++m) s = ffma2_s(wy[m], gv[KC + n][JC + m], s);
    tot = ffma2_s(wz[n], s, tot);
  }
  return fmul2_s(r.cw.x, tot);
}

template<int NS, bool SPREAD>
__global__ void __launch_bounds__(32, 16) k_sweep3q(const QArgs<NS> a) {
  using CF = QCfg<NS>;
  constexpr int PP = CF::PP, MS = CF::MS, YR = CF::YR, ZT = CF::ZT, SX = CF::SX;
  extern __shared__ __align__(16) unsigned char smem[];
  float4 *stage4 = reinterpret_cast<float4 *>(smem);
  float *rec     = reinterpret_cast<float *>(smem + CF::STAGE_BYTES);
  float2 *part   = reinterpret_cast<float2 *>(smem + CF::STAGE_BYTES + CF::REC_BYTES);

  const int lane = threadIdx.x;
  const int la = lane >> 2, by = (lane >> 1) & 1, bz = lane & 1;
  const SweepItem item = a.items[blockIdx.x];
  const int nb2 = a.g.nb[1];
  const int i2 = item.row % nb2, i3 = (item.row / nb2) % a.g.nb[2];  // rows are group-major
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];

  // fine-grid offset of the (z, y) line each flush / fill iteration of this lane serves
  uint32_t lineoff[CF::NRED];
#pragma unroll
  for (int k = 0; k < CF::NRED; ++k) {
    const int yz = (lane + 32 * k) >> 1, z = yz / YR, yc = yz - z * YR;
    const int gz = wrap_index(kBinZ * i3 - CF::HL + z, nf3),
              gy = wrap_index(kBinY * i2 - CF::HL + yc, nf2);
    lineoff[k] = z < ZT ? ((uint32_t)gz * (uint32_t)nf2 + (uint32_t)gy) * (uint32_t)nf1
                        : 0xffffffffu;
  }

  // register window: x columns [jw, jw+8), this lane's column is x = la (mod 8);
  // acc[n][m] = tile row z = bz + 2n, y = by + 2m
  float2 acc[MS][MS];
#pragma unroll
  for (int n = 0; n < MS; ++n)
#pragma unroll
    for (int m = 0; m < MS; ++m) acc[n][m] = float2{0.f, 0.f};
  constexpr int NONE = INT_MIN;
  int jw    = NONE;  // first x column of the window (even); NONE = window empty
  int sbase = 0;     // x of staging column 0 (multiple of 4)
  int stlo  = 0;     // spread: first staging column that holds data

  // staging cell (z, yc, xs) is float2 index (z*YR + yc)*SX + xs
  const uint32_t my_stage = (uint32_t)__cvta_generic_to_shared(smem) +
                            (uint32_t)((bz * YR + by) * SX * sizeof(float2));
  // does slot (n, m) of this lane exist in the tile?
  auto in_tile = [&](int n, int m) { return bz + 2 * n < ZT && by + 2 * m < YR; };

  auto flush_stage = [&](int hi) {
    __syncwarp();
    const int pr = lane & 1;
    if (2 * pr >= stlo && 2 * pr < hi) {
      const uint32_t gx = (uint32_t)wrap_index(sbase + 2 * pr, nf1);
#pragma unroll
      for (int k = 0; k < CF::NRED; ++k) {
        if (lineoff[k] != 0xffffffffu) {
          const float4 v = stage4[lane + 32 * k];
          if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
            atomicAdd(reinterpret_cast<float4 *>(a.fw + lineoff[k] + gx), v);
        }
      }
    }
    __syncwarp();
  };
  auto fill_stage = [&]() {
    __syncwarp();
    const uint32_t gx = (uint32_t)wrap_index(sbase + 2 * (lane & 1), nf1);
#pragma unroll
    for (int k = 0; k < CF::NRED; ++k)
      if (lineoff[k] != 0xffffffffu)
        stage4[lane + 32 * k] = __ldg(reinterpret_cast<const float4 *>(a.fw + lineoff[k] + gx));
    __syncwarp();
  };
  // spread: the window's first two columns leave, their owners park them in staging.
  // interp: two new columns enter at the far end, their owners pick them up from staging.
  auto slide = [&]() {
    if (SPREAD) {
      const int rel   = (la - jw) & 7;  // this lane's column is jw + rel
      const int leave = rel < 2;
      if (leave) {
        const uint32_t dst = my_stage + (uint32_t)((jw - sbase + rel) * sizeof(float2));
        static_for<0, MS>([&](auto nc_) {
          constexpr int n = decltype(nc_)::value;
          static_for<0, MS>([&](auto mc_) {
            constexpr int m = decltype(mc_)::value;
            if constexpr (2 * n < ZT && 2 * m < YR) {
              if ((2 * n + 1 < ZT && 2 * m + 1 < YR) || in_tile(n, m))
                sts64<((2 * n * YR + 2 * m) * SX) * (int)sizeof(float2)>(dst, acc[n][m]);
            }
          });
        });
      }
      const float keep = leave ? 0.f : 1.f;  // clear the columns that left (see sweep3d.cu)
#pragma unroll
      for (int n = 0; n < MS; ++n)
#pragma unroll
        for (int m = 0; m < MS; ++m) acc[n][m] = fmul2_s(keep, acc[n][m]);
      jw += 2;
      if (jw - sbase == SX) {
        flush_stage(SX);
        sbase += SX;
        stlo = 0;
      }
    } else {
      const int xn = jw + CF::W;  // columns jw+8, jw+9 enter; staging holds [sbase, sbase+4)
      if (xn - sbase == SX) {
        sbase += SX;
        fill_stage();
      }
      const int rel = (la - xn) & 7;
      if (rel < 2) {
        const uint32_t src = my_stage + (uint32_t)((xn - sbase + rel) * sizeof(float2));
        static_for<0, MS>([&](auto nc_) {
          constexpr int n = decltype(nc_)::value;
          static_for<0, MS>([&](auto mc_) {
            constexpr int m = decltype(mc_)::value;
            if constexpr (2 * n < ZT && 2 * m < YR) {
              if ((2 * n + 1 < ZT && 2 * m + 1 < YR) || in_tile(n, m))
                acc[n][m] = lds64<((2 * n * YR + 2 * m) * SX) * (int)sizeof(float2)>(src);
            }
          });
        });
      }
      jw += 2;
    }
  };
  auto empty_window = [&]() {
    if (jw == NONE) return;
    if (SPREAD) {
      for (int k = 0; k < CF::W / 2; ++k) slide();
      if (jw - sbase > stlo) flush_stage(jw - sbase);
    }
    jw = NONE;
  };
  auto advance_to = [&](int x) {
    if (x == jw) return;
    if (jw != NONE && (x < jw || x - jw >= CF::W)) empty_window();
    if (jw == NONE) {
      if (SPREAD) {
        jw    = x;
        sbase = x & ~(SX - 1);
        stlo  = x - sbase;
      } else {  // put the window 8 columns to the left and slide the real ones in
        jw    = x - CF::W;
        sbase = (x & ~(SX - 1)) - SX;  // so that the first slide fills the staging
        if ((x & (SX - 1)) != 0) {     // x sits in the middle of a column group
          sbase += SX;
          fill_stage();
        }
      }
    }
    while (jw < x) slide();
  };

  // ---- software pipeline over chunks of 32 points: raw data two chunks ahead, strength one
  auto load_raw = [&](uint32_t q) {
    RawPoint r{0.f, 0.f, 0.f, 0u};
    if (q < item.qb) {
      r.x = __ldcs(a.xs + q), r.y = __ldcs(a.ys + q), r.z = __ldcs(a.zs + q);
      r.j = __ldcs(a.sidx + q);
    }
    return r;
  };
  auto load_c = [&](uint32_t q, const RawPoint &r) {
    float2 c = make_float2(0.f, 0.f);
    if (SPREAD && q < item.qb) c = __ldg(a.c_in + r.j);
    return c;
  };
  RawPoint r1 = load_raw(item.qa + lane);
  RawPoint r2 = load_raw(item.qa + 32 + lane);
  float2 c1   = load_c(item.qa + lane, r1);

  for (uint32_t q0 = item.qa; q0 < item.qb; q0 += CF::CH) {
    const int nc       = (int)min((uint32_t)CF::CH, item.qb - q0);
    const RawPoint cur = r1;
    const float2 ccur  = c1;
    r1                 = r2;
    c1                 = load_c(q0 + 32 + lane, r1);
    r2                 = load_raw(q0 + 64 + lane);
    int key            = INT_MIN;
    if (lane < nc) key = make_record_q<NS, SPREAD>(a, cur, ccur, rec + lane * CF::RECW, i2, i3);
    // runs of equal key: bit l of heads = point l starts a run
    const int prev       = __shfl_up_sync(0xffffffffu, key, 1);
    const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    __syncwarp();
    auto load = [&](int p) {
      const float *rp = rec + p * CF::RECW;
      LaneRecQ r;
      r.wy = *reinterpret_cast<const float4 *>(rp + 4 * by);
      r.wz = *reinterpret_cast<const float4 *>(rp + 8 + 4 * bz);
      if (SPREAD) r.cw = *reinterpret_cast<const float2 *>(rp + 16 + 2 * la);
      else r.cw = make_float2(rp[16 + la], 0.f);
      return r;
    };
    // add lane pairs and park the 16 partial sums of point p (summed thread-per-point below)
    auto put_part = [&](int p, float2 v) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 1);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
      if ((lane & 1) == 0) part[p * PP + (lane >> 1)] = v;
    };
    int p       = 0;
    LaneRecQ nx = load(0);
    while (p < nc) {
      const int kp        = __shfl_sync(0xffffffffu, key, p);
      const uint32_t rest = p < 31 ? heads >> (p + 1) : 0u;
      int e               = rest ? p + __ffs(rest) : 32;
      e                   = min(e, nc);
      advance_to(2 * (kp >> 4) - CF::XB);
#define B200_RUNQ(JC, KC)                                                   \
  case 3 * JC + KC:                                                         \
    for (int q = p; q < e; ++q) {                                           \
      const LaneRecQ n2 = load(q + 1);                                      \
      if (SPREAD) spread_update_q<NS, JC, KC>(acc, nx);                     \
      else put_part(q, interp_gather_q<NS, JC, KC>(acc, nx));               \
      nx = n2;                                                              \
    }                                                                       \
    break;
      switch (kp & 15) {
        B200_RUNQ(0, 0) B200_RUNQ(0, 1) B200_RUNQ(0, 2) B200_RUNQ(1, 0) B200_RUNQ(1, 1)
        B200_RUNQ(1, 2) B200_RUNQ(2, 0) B200_RUNQ(2, 1) B200_RUNQ(2, 2)
      default: break;
      }
#undef B200_RUNQ
      p = e;
    }
    __syncwarp();
    if (!SPREAD && lane < nc) {  // lane = point: add its 16 partial sums, scatter
      float2 s = part[lane * PP], s2 = part[lane * PP + 1];
#pragma unroll
      for (int r = 2; r < 16; r += 2) {
        const float2 v = part[lane * PP + r], w = part[lane * PP + r + 1];
        s.x += v.x, s.y += v.y;
        s2.x += w.x, s2.y += w.y;
      }
      a.c_out[cur.j] = float2{s.x + s2.x, s.y + s2.y};
    }
    __syncwarp();
  }
  empty_window();
}

template<int NS, bool SPREAD>
cudaError_t launch_q(const SweepPoints &pts, const GridGeom<float> &g, int nc, const float *coef,
                     const float2 *c_in, float2 *c_out, float2 *fw, cudaStream_t st) {
  using CF = QCfg<NS>;
  if (pts.nitems == 0) return cudaSuccess;
  QArgs<NS> a;
  a.xs = pts.xs, a.ys = pts.ys, a.zs = pts.zs, a.sidx = pts.sidx, a.items = pts.items;
  a.g = g;
  constexpr int rows = TableRows<NS>::value;
  for (int k = 0; k < rows; ++k)
    for (int j = 0; j < 8; ++j) {
      const int src      = k - (rows - nc);
      a.tab.c[k * 8 + j] = (src >= 0 && j < NS) ? coef[src * NS + j] : 0.f;
    }
  a.c_in  = c_in;
  a.c_out = c_out;
  a.fw    = fw;
  const size_t shbytes = CF::STAGE_BYTES + CF::REC_BYTES + (SPREAD ? 0 : CF::PART_BYTES);
  auto kern            = k_sweep3q<NS, SPREAD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)shbytes);
  if (e != cudaSuccess) return e;
  kern<<<pts.nitems, 32, shbytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_spread3_sweepq(int ns, const SweepPoints &pts, const GridGeom<float> &g, int nc,
                                  const float *coef, const float2 *c_in, float2 *fw,
                                  cudaStream_t st) {
  switch (ns) {
  case 6: return launch_q<6, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  case 7: return launch_q<7, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_interp3_sweepq(int ns, const SweepPoints &pts, const GridGeom<float> &g, int nc,
                                  const float *coef, float2 *c_out, const float2 *fw,
                                  cudaStream_t st) {
  switch (ns) {
  case 6:
    return launch_q<6, false>(pts, g, nc, coef, nullptr, c_out, const_cast<float2 *>(fw), st);
  case 7:
    return launch_q<7, false>(pts, g, nc, coef, nullptr, c_out, const_cast<float2 *>(fw), st);
  default: return cudaErrorInvalidValue;
  }
}

}  // namespace b200
