// Row-sweep spread / interp kernels for 3D float (see sweep3d.cuh for the design).
#include "sweep3d.cuh"

#include "sweepmath.cuh"

#include <limits.h>
#include <stdlib.h>

#include <type_traits>

namespace b200 {

template<int OFF> __device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "f"(v.x), "f"(v.y)
               : "memory");
}
template<int OFF> __device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(OFF));
  return v;
}
template<int NS> struct SweepCfg {
  static constexpr int HL   = NS / 2;               // cells left of a bin a stencil can reach
  static constexpr int W    = 8;                    // x rows of the register window
  static constexpr int BQ   = 4;                    // lane groups along z
  static constexpr int YR   = kBinY + NS;           // cells of one register row (along y)
  static constexpr int ZT   = kBinZ + NS;           // z rows of the tile
  static constexpr int RZ   = (ZT + BQ - 1) / BQ;   // z rows per lane: z = bq + BQ*m
  static constexpr int XB   = 4;                    // window positions are (i0 + XB) >> 1
  static constexpr int SX   = 4;                    // x columns collected per flush / fill
  static constexpr int NRED = (ZT * YR * (SX / 2) + 31) / 32;  // 16-byte pieces per lane
  static constexpr int CH   = 32;                   // points per chunk: one per lane
  static constexpr int RECW = 36;                   // words per point record
  // record: [0,NS) phi_y  [7] key (window position << 4 | jb)  [8,16) phi_x rotated so that
  //         word 8+a is the weight of the window row x = a (mod 8)  [16+4*bq+m] phi_z of tile
  //         row bq+4m  [32,33] strength (spread)
  static constexpr size_t STAGE_BYTES = (size_t)NRED * 32 * sizeof(float4);
  static constexpr size_t REC_BYTES   = (size_t)(CH + 2) * RECW * sizeof(float);
  static constexpr int PP = 17;  // pitch of a point's 16 partial sums (odd: conflict-free rows)
  static constexpr size_t PART_BYTES  = (size_t)CH * PP * sizeof(float2);  // interp partial sums
  static_assert(NS + 1 <= W, "two stencil starts must fit the window");
  static_assert(HL <= XB && NS <= 7 && RZ <= 3, "layout");
};

template<int NS> struct SweepArgs {
  const float *xs, *ys, *zs;   // coordinates, refined bin order
  const uint32_t *sidx;        // position -> user index
  const SweepItem *items;
  GridGeom<float> g;
  typename SweepTab<float, NS>::type tab;
  const float2 *c_in;
  float2 *c_out;
  float2 *fw;
#ifdef B200_EXPERIMENTS
  int dbg;  // experiment switches, compiled in only with -DB200_EXPERIMENTS (never in the library)
#endif
};
// experiment switches cost nothing and cannot change results unless compiled in
#ifdef B200_EXPERIMENTS
#define B200_DBG(a, bit) ((a).dbg & (bit))
#else
#define B200_DBG(a, bit) 0
#endif

// raw data of one point, as prefetched
struct RawPoint {
  float x, y, z;
  uint32_t j;
};

// Thread-per-point preparation: fold, stencil starts, windows (+ strength) -> record.
template<int NS>
__device__ __forceinline__ void make_record(const SweepArgs<NS> &a, const RawPoint &pt, float2 c,
                                            float *rec, int i2, int i3) {
  using CF = SweepCfg<NS>;
  int i0;
  float x1;
  float kv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  stencil_start<float, NS>(fold_rescale<float>(pt.x, a.g.nf_t[0]), i0, x1);
  eval_window_t(a.tab, x1, kv);
  const int gpos = (i0 + CF::XB) >> 1;  // window start 2*gpos - XB <= i0 <= that + 1
  // rotate: the weight of stencil cell t belongs to the window row x = i0 + t (mod 8)
#pragma unroll
  for (int t = 0; t < 8; ++t) rec[8 + ((i0 + t) & 7)] = t < NS ? kv[t] : 0.f;
#pragma unroll
  for (int s = 0; s < 8; ++s) kv[s] = 0.f;
  stencil_start<float, NS>(fold_rescale<float>(pt.y, a.g.nf_t[1]), i0, x1);
  eval_window_t(a.tab, x1, kv);
  const int jb = min(max(i0 - (kBinY * i2 - CF::HL), 0), kBinY);
  kv[7]        = __int_as_float((gpos << 4) | jb);
  *reinterpret_cast<float4 *>(rec)     = make_float4(kv[0], kv[1], kv[2], kv[3]);
  *reinterpret_cast<float4 *>(rec + 4) = make_float4(kv[4], kv[5], kv[6], kv[7]);
  stencil_start<float, NS>(fold_rescale<float>(pt.z, a.g.nf_t[2]), i0, x1);
  eval_window_t(a.tab, x1, kv);
  // z stencil start relative to the tile's first z row, in [0, kBinZ]
  const int k0 = min(max(i0 - (kBinZ * i3 - CF::HL), 0), kBinZ);
  // tile row z = k0 + t lives in word 16 + 4*(z & 3) + (z >> 2)
#pragma unroll
  for (int bq = 0; bq < CF::BQ; ++bq)
    *reinterpret_cast<float4 *>(rec + 16 + 4 * bq) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < NS; ++t) rec[16 + 4 * ((k0 + t) & 3) + ((k0 + t) >> 2)] = kv[t];
  *reinterpret_cast<float2 *>(rec + 32) = c;
}

// what one lane needs of one record
struct LaneRec {
  float4 k0, k1;  // phi_y[0..6], key
  float wx;       // phi_x of this lane's window row
  float4 fz;      // phi_z of this lane's tile rows
  float2 c;       // strength (spread only)
};

template<int NS, int JB>
__device__ __forceinline__ void spread_update(float2 (&acc)[SweepCfg<NS>::RZ][SweepCfg<NS>::YR],
                                              const LaneRec &r) {
  using CF = SweepCfg<NS>;
  const float ky[8] = {r.k0.x, r.k0.y, r.k0.z, r.k0.w, r.k1.x, r.k1.y, r.k1.z, 0.f};
  const float fz[4] = {r.fz.x, r.fz.y, r.fz.z, r.fz.w};
  const float2 cw   = fmul2_s(r.wx, r.c);
#pragma unroll
  for (int m = 0; m < CF::RZ; ++m) {
    const float2 w = fmul2_s(fz[m], cw);
#pragma unroll
    for (int t = 0; t < NS; ++t) acc[m][JB + t] = ffma2_s(ky[t], w, acc[m][JB + t]);
  }
}
// this lane's share of one point's interpolated value
template<int NS, int JB>
__device__ __forceinline__ float2 interp_gather(
    const float2 (&gv)[SweepCfg<NS>::RZ][SweepCfg<NS>::YR], const LaneRec &r) {
  using CF = SweepCfg<NS>;
  const float ky[8] = {r.k0.x, r.k0.y, r.k0.z, r.k0.w, r.k1.x, r.k1.y, r.k1.z, 0.f};
  const float fz[4] = {r.fz.x, r.fz.y, r.fz.z, r.fz.w};
  float2 tot = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < CF::RZ; ++m) {
    // one chain per row (splitting it in two was measured slower: the extra packed add per row
    // costs more FMA-pipe time than the shorter dependency chain saves)
    float2 s = fmul2_s(ky[0], gv[m][JB]);
#pragma unroll
    for (int t = 1; t < NS; ++t) s = ffma2_s(ky[t], gv[m][JB + t], s);
    tot = ffma2_s(fz[m], s, tot);
  }
  return fmul2_s(r.wx, tot);
}

// One warp per work item: a run of consecutive points of one row of bins (i2, i3), which the
// refined bin order keeps sorted by x window position.  SPREAD: accumulate into the register
// window and reduce leaving rows into the fine grid; else: fill entering rows from the fine grid
// and gather.
// 16 resident one-warp blocks per SM (128 registers each).  Measured on B200 at C3: asking for
// 20 blocks (96 registers, a few spilled words) takes 10.3 ms instead of 8.5, 24 blocks (80
// registers) 13.6 ms: the kernel is bound by the FMA pipe, not by latency, so extra warps only
// add spill traffic.
template<int NS, bool SPREAD>
__global__ void __launch_bounds__(32, 16) k_sweep3(const SweepArgs<NS> a) {
  constexpr int PP = SweepCfg<NS>::PP;
  using CF = SweepCfg<NS>;
  extern __shared__ __align__(16) unsigned char smem[];
  float4 *stage4 = reinterpret_cast<float4 *>(smem);
  float *rec     = reinterpret_cast<float *>(smem + CF::STAGE_BYTES);
  float2 *part   = reinterpret_cast<float2 *>(smem + CF::STAGE_BYTES + CF::REC_BYTES);

  const int lane = threadIdx.x;
  const int la = lane >> 2, bq = lane & 3;
  const SweepItem item = a.items[blockIdx.x];
  const int nb2 = a.g.nb[1];
  const int i2 = item.row % nb2, i3 = (item.row / nb2) % a.g.nb[2];  // rows are group-major
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];

  // fine-grid offset of the (z, y) line each flush / fill iteration of this lane serves
  uint32_t lineoff[CF::NRED];
#pragma unroll
  for (int k = 0; k < CF::NRED; ++k) {
    const int yz = (lane + 32 * k) >> 1, z = yz / CF::YR, yc = yz - z * CF::YR;
    const int gz = grid_plane(a.g, wrap_index(kBinZ * i3 - CF::HL + z, nf3)),
              gy = wrap_index(kBinY * i2 - CF::HL + yc, nf2);
    // planes outside a z window (sort.cuh) are never touched by this plan's points
    lineoff[k] = z < CF::ZT && gz >= 0
                     ? ((uint32_t)gz * (uint32_t)nf2 + (uint32_t)gy) * (uint32_t)nf1
                     : 0xffffffffu;
  }

  // register window: x rows [jw, jw+8), this lane's row is x = la (mod 8); tile rows z = bq+4m
  float2 acc[CF::RZ][CF::YR];
#pragma unroll
  for (int m = 0; m < CF::RZ; ++m)
#pragma unroll
    for (int c = 0; c < CF::YR; ++c) acc[m][c] = float2{0.f, 0.f};
  constexpr int NONE = INT_MIN;
  int jw    = NONE;  // first x row of the window (even); NONE = window empty
  int sbase = 0;     // x of staging column 0 (multiple of 4)
  int stlo  = 0;     // spread: first staging column that holds data

  // staging cell (z, yc, xs) is float2 index (z*YR + yc)*SX + xs
  const uint32_t my_stage =
      (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(bq * CF::YR * CF::SX * sizeof(float2));

  // ---- spread: add staging columns [stlo, hi) to the fine grid, two x cells per reduction
  auto flush_stage = [&](int hi) {
    __syncwarp();
    const int pr = lane & 1;
    if (2 * pr >= stlo && 2 * pr < hi) {
      const uint32_t gx = (uint32_t)wrap_index(sbase + 2 * pr, nf1);
#pragma unroll
      for (int k = 0; k < CF::NRED; ++k) {
        if (lineoff[k] != 0xffffffffu) {
          const float4 v = stage4[lane + 32 * k];
          // skip all-zero pieces (sparse rows); one integer test instead of four float compares
          const uint32_t bits = __float_as_uint(v.x) | __float_as_uint(v.y) |
                                __float_as_uint(v.z) | __float_as_uint(v.w);
          if ((bits << 1) != 0u && !B200_DBG(a, 1))
            atomicAdd(reinterpret_cast<float4 *>(a.fw + lineoff[k] + gx), v);
        }
      }
    }
    __syncwarp();
  };
  // ---- interp: load fine-grid columns [sbase, sbase+4) of the tile into staging
  auto fill_stage = [&]() {
    __syncwarp();
    const uint32_t gx = (uint32_t)wrap_index(sbase + 2 * (lane & 1), nf1);
#pragma unroll
    for (int k = 0; k < CF::NRED; ++k)
      if (lineoff[k] != 0xffffffffu)
        stage4[lane + 32 * k] = __ldg(reinterpret_cast<const float4 *>(a.fw + lineoff[k] + gx));
      else if (lane + 32 * k < CF::ZT * CF::YR * (CF::SX / 2))
        stage4[lane + 32 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  };
  // spread: the window's first two x rows leave, their owners park them in staging.
  // interp: two new x rows enter at the far end, their owners pick them up from staging.
  auto slide = [&]() {
    if (SPREAD) {
      const int rel   = (la - jw) & 7;  // this lane's row is jw + rel
      const int leave = rel < 2;
      if (leave) {
        const uint32_t dst = my_stage + (uint32_t)((jw - sbase + rel) * sizeof(float2));
        static_for<0, CF::RZ>([&](auto mc) {
          constexpr int m = decltype(mc)::value;
          static_for<0, CF::YR>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            if ((m + 1) * CF::BQ <= CF::ZT || bq + m * CF::BQ < CF::ZT)
              sts64<(m * CF::BQ * CF::YR + c) * CF::SX * (int)sizeof(float2)>(dst, acc[m][c]);
          });
        });
      }
      // clear the rows that left: one packed multiply per accumulator, no branch, so the
      // accumulators stay in place (a select costs two instructions per accumulator)
      const float keep = leave ? 0.f : 1.f;
#pragma unroll
      for (int m = 0; m < CF::RZ; ++m)
#pragma unroll
        for (int c = 0; c < CF::YR; ++c) acc[m][c] = fmul2_s(keep, acc[m][c]);
      jw += 2;
      if (jw - sbase == CF::SX) {
        flush_stage(CF::SX);
        sbase += CF::SX;
        stlo = 0;
      }
    } else {
      // rows jw+8, jw+9 enter; staging holds columns [sbase, sbase+4)
      const int xn = jw + CF::W;
      if (xn - sbase == CF::SX) {
        sbase += CF::SX;
        fill_stage();
      }
      const int r0 = xn & 7, r1 = (xn + 1) & 7;
      if (la == r0 || la == r1) {
        const uint32_t src =
            my_stage + (uint32_t)((xn - sbase + (la == r1 ? 1 : 0)) * sizeof(float2));
        static_for<0, CF::RZ>([&](auto mc) {
          constexpr int m = decltype(mc)::value;
          static_for<0, CF::YR>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            if ((m + 1) * CF::BQ <= CF::ZT || bq + m * CF::BQ < CF::ZT)
              acc[m][c] = lds64<(m * CF::BQ * CF::YR + c) * CF::SX * (int)sizeof(float2)>(src);
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
        sbase = x & ~(CF::SX - 1);
        stlo  = x - sbase;
      } else {  // put the window 8 rows to the left and slide the real rows in
        jw    = x - CF::W;
        sbase = (x & ~(CF::SX - 1)) - CF::SX;  // so that the first slide fills the staging
        if ((x & (CF::SX - 1)) != 0) {         // x sits in the middle of a column group
          sbase += CF::SX;
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
    if (SPREAD && q < item.qb && !B200_DBG(a, 4)) c = __ldg(a.c_in + r.j);
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
    if (lane < nc) make_record<NS>(a, cur, ccur, rec + lane * CF::RECW, i2, i3);
    if (lane < 2) rec[(nc + lane) * CF::RECW + 7] = __int_as_float(-1);  // sentinels
    __syncwarp();
    auto load = [&](int p) {
      const float *rp = rec + p * CF::RECW;
      LaneRec r;
      r.k0 = *reinterpret_cast<const float4 *>(rp);
      r.k1 = *reinterpret_cast<const float4 *>(rp + 4);
      r.wx = rp[8 + la];
      r.fz = *reinterpret_cast<const float4 *>(rp + 16 + 4 * bq);
      if (SPREAD) r.c = *reinterpret_cast<const float2 *>(rp + 32);
      return r;
    };
    // add lane pairs and park the 16 partial sums of point p; the rest of the reduction runs
    // thread-per-point after the chunk (10.58 -> 10.08 ms at C3 against two shuffle stages;
    // parking all 32 partials with no shuffle at all measured the same as this)
    auto put_part = [&](int p, float2 v) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 1);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
      if ((lane & 1) == 0) part[p * PP + (lane >> 1)] = v;
    };
    int p      = 0;
    LaneRec nx = load(0);
    while (!B200_DBG(a, 2)) {
      const int key = __float_as_int(nx.k1.w);
      if (key < 0) break;
      advance_to(2 * (key >> 4) - CF::XB);
#define B200_RUN(J)                                                 \
  case J:                                                           \
    for (;;) {                                                      \
      LaneRec n2 = load(p + 1);                                     \
      if (SPREAD) spread_update<NS, J>(acc, nx);                    \
      else put_part(p, interp_gather<NS, J>(acc, nx));              \
      if (__float_as_int(n2.k1.w) != key) {                         \
        nx = n2;                                                    \
        p += 1;                                                     \
        break;                                                      \
      }                                                             \
      nx = load(p + 2);                                             \
      if (SPREAD) spread_update<NS, J>(acc, n2);                    \
      else put_part(p + 1, interp_gather<NS, J>(acc, n2));          \
      p += 2;                                                       \
      if (__float_as_int(nx.k1.w) != key) break;                    \
    }                                                               \
    break;
      switch (key & 15) {
        B200_RUN(0) B200_RUN(1) B200_RUN(2) B200_RUN(3) B200_RUN(4)
      default: __builtin_unreachable();
      }
#undef B200_RUN
    }
    __syncwarp();
    if (!SPREAD && lane < nc) {  // lane = point: add its 16 partial sums, scatter
      float2 s = part[lane * PP], s2 = part[lane * PP + 1];
#pragma unroll
      for (int r = 2; r < PP - 1; r += 2) {
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
static cudaError_t launch_ns(const SweepPoints &pts, const GridGeom<float> &g, int nc,
                             const float *coef, const float2 *c_in, float2 *c_out, float2 *fw,
                             cudaStream_t st) {
  using CF = SweepCfg<NS>;
  if (pts.nitems == 0) return cudaSuccess;
  SweepArgs<NS> a;
  a.xs = pts.xs, a.ys = pts.ys, a.zs = pts.zs, a.sidx = pts.sidx, a.items = pts.items;
  a.g = g;
  fill_table(a.tab, nc, coef);
  a.c_in  = c_in;
  a.c_out = c_out;
  a.fw    = fw;
#ifdef B200_EXPERIMENTS
  a.dbg = getenv("B200_SWEEP_DBG") ? atoi(getenv("B200_SWEEP_DBG")) : 0;
#endif
  const size_t shbytes = CF::STAGE_BYTES + CF::REC_BYTES + (SPREAD ? 0 : CF::PART_BYTES);
  auto kern            = k_sweep3<NS, SPREAD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)shbytes);
  if (e != cudaSuccess) return e;
  kern<<<pts.nitems, 32, shbytes, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_spread3_sweep(int ns, const SweepPoints &pts, const GridGeom<float> &g, int nc,
                                 const float *coef, const float2 *c_in, float2 *fw,
                                 cudaStream_t st) {
  switch (ns) {
  case 2: return launch_ns<2, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  case 3: return launch_ns<3, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  case 4: return launch_ns<4, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  case 5: return launch_ns<5, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  case 6: return launch_ns<6, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  case 7: return launch_ns<7, true>(pts, g, nc, coef, c_in, nullptr, fw, st);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_interp3_sweep(int ns, const SweepPoints &pts, const GridGeom<float> &g, int nc,
                                 const float *coef, float2 *c_out, const float2 *fw,
                                 cudaStream_t st) {
  float2 *g_ = const_cast<float2 *>(fw);
  switch (ns) {
  case 2: return launch_ns<2, false>(pts, g, nc, coef, nullptr, c_out, g_, st);
  case 3: return launch_ns<3, false>(pts, g, nc, coef, nullptr, c_out, g_, st);
  case 4: return launch_ns<4, false>(pts, g, nc, coef, nullptr, c_out, g_, st);
  case 5: return launch_ns<5, false>(pts, g, nc, coef, nullptr, c_out, g_, st);
  case 6: return launch_ns<6, false>(pts, g, nc, coef, nullptr, c_out, g_, st);
  case 7: return launch_ns<7, false>(pts, g, nc, coef, nullptr, c_out, g_, st);
  default: return cudaErrorInvalidValue;
  }
}

}  // namespace b200
