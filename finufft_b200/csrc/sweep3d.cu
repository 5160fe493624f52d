// Tube-sweep spread / interp kernels for 3D float (see sweep3d.cuh for the design).
#include "sweep3d.cuh"

#include <limits.h>
#include <stdlib.h>

namespace b200 {

// (re,im) += s * (wr,wi): one packed FFMA2 (the scalar operand is broadcast by the hardware)
__device__ __forceinline__ float2 ffma2_s(float s, float2 w, float2 acc) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n"
      " mov.b64 ra, {%2,%2};\n mov.b64 rb, {%3,%4};\n mov.b64 rc, {%5,%6};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(s), "f"(w.x), "f"(w.y), "f"(acc.x), "f"(acc.y));
  return d;
}
__device__ __forceinline__ float2 fmul2_s(float s, float2 w) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n"
      " mov.b64 ra, {%2,%2};\n mov.b64 rb, {%3,%4};\n"
      " mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(s), "f"(w.x), "f"(w.y));
  return d;
}

template<int NS> struct SweepCfg {
  static constexpr int HL   = NS / 2;               // cells left of a bin a stencil can reach
  static constexpr int W    = 8;                    // x rows of the register window
  static constexpr int BQ   = 4;                    // lane groups along z
  static constexpr int YR   = kBinY + NS;           // cells of one register row (along y)
  static constexpr int ZT   = kBinZ + NS;           // z rows of the tile
  static constexpr int RZ   = (ZT + BQ - 1) / BQ;   // z rows per lane: z = bq + BQ*m
  static constexpr int NG   = kBinX / 2 + 1;        // window positions per bin (steps of 2 cells)
  static constexpr int NJB  = kBinY + 1;            // y stencil starts inside one bin
  static constexpr int XB   = 4;                    // window origin of bin i1 is kBinX*i1 - XB
  static constexpr int SX   = 4;                    // x columns collected before a flush
  static constexpr int NRED = (ZT * YR * (SX / 2) + 31) / 32;  // flush iterations per lane
  static constexpr int CH   = 64;                   // points per chunk (records in flight)
  static constexpr int LCAP = 512;                  // points per sorted batch of one bin
  static constexpr int RECW = 36;                   // words per point record
  // record: [0,NS) phi_y  [7] key (g<<4 | jb)  [8,16) phi_x rotated so that word 8+a is the
  //         weight of the window row x = a (mod 8)  [16+4*bq+m] phi_z of tile row bq+4m
  //         [32,33] strength
  static constexpr size_t STAGE_BYTES = (size_t)NRED * 32 * sizeof(float4);
  static constexpr size_t REC_BYTES   = (size_t)(CH + 1) * RECW * sizeof(float);
  static constexpr size_t LIST_BYTES  = (size_t)LCAP * sizeof(uint16_t);
  static constexpr size_t CNT_BYTES   = 64 * sizeof(int);
  static constexpr size_t BYTES       = STAGE_BYTES + REC_BYTES + LIST_BYTES + CNT_BYTES;
  static_assert(NS + 1 <= W, "two stencil starts must fit the window");
  static_assert(HL <= XB && NS <= 7 && RZ <= 3 && NG * NJB <= 64, "layout");
};

// Horner table padded to 8 columns so that two neighbouring panels form one aligned pair:
// c[k*8 + j], k = 0 highest degree, columns j >= NS are zero.
template<int NS> struct alignas(16) PairTable {
  float c[TableRows<NS>::value * 8];
};
// NS window values by packed Horner (two panels per FFMA2); same roundings as eval_window.
template<int NS>
__device__ __forceinline__ void eval_window2(const PairTable<NS> &tab, float x1, float (&out)[8]) {
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

template<int NS> struct SweepArgs {
  PointSet<float> pts;
  GridGeom<float> g;
  PairTable<NS> tab;
  const float2 *c_in;
  float2 *c_out;
  float2 *fw;
  int nsplit;  // work items per row of bins
  int ypi;     // x bins per item
  int dbg;     // experiment switches (0 in production)
};

__device__ __forceinline__ int pmod(int v, int n) {
  int r = v % n;
  return r < 0 ? r + n : r;
}

// Sort key of a point inside bin (i1, i2): window position g and y stencil start jb.
template<int NS>
__device__ __forceinline__ int point_key(const SweepArgs<NS> &a, uint32_t q, int i1, int i2) {
  using CF = SweepCfg<NS>;
  int i0, j0;
  float t;
  stencil_start<float, NS>(fold_rescale<float>(a.pts.xs[q], a.g.nf_t[0]), i0, t);
  stencil_start<float, NS>(fold_rescale<float>(a.pts.ys[q], a.g.nf_t[1]), j0, t);
  const int g  = min(max((i0 - (kBinX * i1 - CF::XB)) >> 1, 0), CF::NG - 1);
  const int jb = min(max(j0 - (kBinY * i2 - CF::HL), 0), kBinY);
  return g * CF::NJB + jb;
}

// Thread-per-point preparation: fold, stencil starts, windows, strength -> record.
template<int NS, bool SPREAD>
__device__ __forceinline__ void make_record(const SweepArgs<NS> &a, uint32_t q, float *rec, int i1,
                                            int i2, int i3) {
  using CF = SweepCfg<NS>;
  float2 c = make_float2(0.f, 0.f);
  if (SPREAD) c = __ldcs(a.c_in + __ldcs(a.pts.sidx + q));
  int i0;
  float x1;
  float kv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  stencil_start<float, NS>(fold_rescale<float>(__ldcs(a.pts.xs + q), a.g.nf_t[0]), i0, x1);
  eval_window2<NS>(a.tab, x1, kv);
  i0 = min(max(i0, kBinX * i1 - CF::XB), kBinX * i1 - CF::XB + 2 * CF::NG - 1);
  const int g = (i0 - (kBinX * i1 - CF::XB)) >> 1;
  // rotate: the weight of stencil cell t belongs to the window row x = i0 + t (mod 8)
#pragma unroll
  for (int t = 0; t < 8; ++t) rec[8 + ((i0 + t) & 7)] = t < NS ? kv[t] : 0.f;
#pragma unroll
  for (int s = 0; s < 8; ++s) kv[s] = 0.f;
  stencil_start<float, NS>(fold_rescale<float>(__ldcs(a.pts.ys + q), a.g.nf_t[1]), i0, x1);
  eval_window2<NS>(a.tab, x1, kv);
  const int jb = min(max(i0 - (kBinY * i2 - CF::HL), 0), kBinY);
  kv[7]        = __int_as_float((g << 4) | jb);
  *reinterpret_cast<float4 *>(rec)     = make_float4(kv[0], kv[1], kv[2], kv[3]);
  *reinterpret_cast<float4 *>(rec + 4) = make_float4(kv[4], kv[5], kv[6], kv[7]);
  stencil_start<float, NS>(fold_rescale<float>(__ldcs(a.pts.zs + q), a.g.nf_t[2]), i0, x1);
  eval_window2<NS>(a.tab, x1, kv);
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

__device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// what one lane needs of one record
struct SpreadRec {
  float4 k0, k1;  // phi_y[0..6], key
  float wx;       // phi_x of this lane's window row
  float4 fz;      // phi_z of this lane's tile rows
  float2 c;
};

template<int NS, int JB>
__device__ __forceinline__ void spread_update(float2 (&acc)[SweepCfg<NS>::RZ][SweepCfg<NS>::YR],
                                              const SpreadRec &r) {
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

// One warp per work item: the row of bins (i2, i3), swept along x.
template<int NS>
__global__ void __launch_bounds__(32, 16) k_spread3_sweep(const SweepArgs<NS> a) {
  using CF = SweepCfg<NS>;
  extern __shared__ __align__(16) unsigned char smem[];
  float4 *stage4 = reinterpret_cast<float4 *>(smem);
  float *rec     = reinterpret_cast<float *>(smem + CF::STAGE_BYTES);
  uint16_t *list = reinterpret_cast<uint16_t *>(smem + CF::STAGE_BYTES + CF::REC_BYTES);
  int *cnt       = reinterpret_cast<int *>(smem + CF::STAGE_BYTES + CF::REC_BYTES + CF::LIST_BYTES);

  const int lane = threadIdx.x;
  const int la = lane >> 2, bq = lane & 3;
  const int nb1 = a.g.nb[0], nb2 = a.g.nb[1];
  const int row = blockIdx.x / a.nsplit, part = blockIdx.x % a.nsplit;
  const int i2 = row % nb2, i3 = row / nb2;
  const int xb0 = part * a.ypi, xb1 = min(nb1, xb0 + a.ypi);
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];

  // fine-grid offset of the (z, y) line each flush iteration of this lane serves
  uint32_t lineoff[CF::NRED];
#pragma unroll
  for (int k = 0; k < CF::NRED; ++k) {
    const int yz = (lane + 32 * k) >> 1, z = yz / CF::YR, yc = yz - z * CF::YR;
    const int gz = wrap_index(kBinZ * i3 - CF::HL + z, nf3),
              gy = wrap_index(kBinY * i2 - CF::HL + yc, nf2);
    lineoff[k] = z < CF::ZT ? ((uint32_t)gz * (uint32_t)nf2 + (uint32_t)gy) * (uint32_t)nf1
                            : 0xffffffffu;
  }

  float2 acc[CF::RZ][CF::YR];
#pragma unroll
  for (int m = 0; m < CF::RZ; ++m)
#pragma unroll
    for (int c = 0; c < CF::YR; ++c) acc[m][c] = float2{0.f, 0.f};
  constexpr int NONE = INT_MIN;
  int jw = NONE;  // first x row of the register window (even); NONE = window empty
  int sbase = 0;  // x of staging column 0 (multiple of 4)
  int stlo  = 0;  // first staging column that holds data

  // staging cell (z, yc, xs) is float2 index (z*YR + yc)*SX + xs
  const uint32_t my_stage =
      (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(bq * CF::YR * CF::SX * sizeof(float2));

  // add staging columns [stlo, hi) to the fine grid, 16 bytes (two x cells) per reduction
  auto flush_stage = [&](int hi) {
    __syncwarp();
    const int pr = lane & 1;
    if (2 * pr >= stlo && 2 * pr < hi) {
      const uint32_t gx = (uint32_t)wrap_index(sbase + 2 * pr, nf1);
#pragma unroll
      for (int k = 0; k < CF::NRED; ++k) {
        if (lineoff[k] != 0xffffffffu) {
          const float4 v = stage4[lane + 32 * k];
          if ((v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) && !(a.dbg & 1))
            atomicAdd(reinterpret_cast<float4 *>(a.fw + lineoff[k] + gx), v);
        }
      }
    }
    __syncwarp();
  };
  // the window's first two x rows leave: their owners park them in the staging tile
  auto slide = [&]() {
    const int r0 = jw & 7, r1 = (jw + 1) & 7;
    if (la == r0 || la == r1) {
      const uint32_t dst = my_stage + (uint32_t)((jw - sbase + (la == r1 ? 1 : 0)) * sizeof(float2));
#pragma unroll
      for (int m = 0; m < CF::RZ; ++m)
#pragma unroll
        for (int c = 0; c < CF::YR; ++c) {
          if (m * CF::BQ < CF::ZT - 3 || bq + m * CF::BQ < CF::ZT)
            sts64(dst + (uint32_t)(((m * CF::BQ * CF::YR) + c) * CF::SX * sizeof(float2)), acc[m][c]);
          acc[m][c] = float2{0.f, 0.f};
        }
    }
    jw += 2;
    if (jw - sbase == CF::SX) {
      flush_stage(CF::SX);
      sbase += CF::SX;
      stlo = 0;
    }
  };
  auto empty_window = [&]() {
    if (jw == NONE) return;
    for (int k = 0; k < CF::W / 2; ++k) slide();
    if (jw - sbase > stlo) flush_stage(jw - sbase);
    jw = NONE;
  };
  auto advance_to = [&](int x) {
    if (jw != NONE && (x < jw || x - jw >= CF::W)) empty_window();
    if (jw == NONE) {
      jw    = x;
      sbase = x & ~(CF::SX - 1);
      stlo  = x - sbase;
    }
    while (jw < x) slide();
  };

  for (int i1 = xb0; i1 < xb1; ++i1) {
    const uint32_t bin = (uint32_t)i1 + (uint32_t)nb1 * ((uint32_t)i2 + (uint32_t)nb2 * (uint32_t)i3);
    const uint32_t qs = a.pts.binstart[bin], qe = a.pts.binstart[bin + 1];
    const int xbase = kBinX * i1 - CF::XB;
    for (uint32_t q0 = qs; q0 < qe; q0 += CF::LCAP) {
      const int n = (int)min((uint32_t)CF::LCAP, qe - q0);
      // ---- counting sort of this batch by (window position, y stencil start)
      cnt[lane] = 0, cnt[lane + 32] = 0;
      __syncwarp();
      for (int k = lane; k < n; k += 32) atomicAdd(&cnt[point_key<NS>(a, q0 + k, i1, i2)], 1);
      __syncwarp();
      {
        const int v0 = cnt[2 * lane], v1 = cnt[2 * lane + 1];
        int incl = v0 + v1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += up;
        }
        __syncwarp();
        cnt[2 * lane]     = incl - v0 - v1;
        cnt[2 * lane + 1] = incl - v1;
      }
      __syncwarp();
      for (int k = lane; k < n; k += 32)
        list[atomicAdd(&cnt[point_key<NS>(a, q0 + k, i1, i2)], 1)] = (uint16_t)k;
      __syncwarp();
      for (int c0 = 0; c0 < n; c0 += CF::CH) {
        const int nc = min(CF::CH, n - c0);
        // ---- records of the next nc points in sorted order, then a sentinel
        for (int k = lane; k < nc; k += 32)
          make_record<NS, true>(a, q0 + list[c0 + k], rec + k * CF::RECW, i1, i2, i3);
        if (lane == 0) rec[nc * CF::RECW + 7] = __int_as_float(-1);
        __syncwarp();
        // ---- accumulate: runs of equal key share the window position and the y offset
        auto load = [&](int p) {
          const float *rp = rec + p * CF::RECW;
          SpreadRec r;
          r.k0 = *reinterpret_cast<const float4 *>(rp);
          r.k1 = *reinterpret_cast<const float4 *>(rp + 4);
          r.wx = rp[8 + la];
          r.fz = *reinterpret_cast<const float4 *>(rp + 16 + 4 * bq);
          r.c  = *reinterpret_cast<const float2 *>(rp + 32);
          return r;
        };
        int p        = 0;
        SpreadRec nx = load(0);
        while (!(a.dbg & 2)) {
          const int key = __float_as_int(nx.k1.w);
          if (key < 0) break;
          advance_to(xbase + 2 * (key >> 4));
#define B200_RUN(J)                                       \
  case J:                                                 \
    do {                                                  \
      const SpreadRec cu = nx;                            \
      nx                 = load(++p);                     \
      spread_update<NS, J>(acc, cu);                      \
    } while (__float_as_int(nx.k1.w) == key);             \
    break;
          switch (key & 15) {
            B200_RUN(0) B200_RUN(1) B200_RUN(2) B200_RUN(3) B200_RUN(4)
          default: __builtin_unreachable();
          }
#undef B200_RUN
        }
        __syncwarp();
      }
    }
  }
  empty_window();
}

template<int NS>
static cudaError_t launch_spread_ns(const PointSet<float> &pts, const GridGeom<float> &g, int nc,
                                    const float *coef, const float2 *c_in, float2 *fw,
                                    cudaStream_t st) {
  using CF = SweepCfg<NS>;
  SweepArgs<NS> a;
  a.pts = pts;
  a.g   = g;
  constexpr int rows = TableRows<NS>::value;
  for (int k = 0; k < rows; ++k)
    for (int j = 0; j < 8; ++j) {
      const int src       = k - (rows - nc);
      a.tab.c[k * 8 + j]  = (src >= 0 && j < NS) ? coef[src * NS + j] : 0.f;
    }
  a.c_in  = c_in;
  a.c_out = nullptr;
  a.fw    = fw;
  // one warp per row of bins (i2, i3); rows are cut along x only when there are too few of them
  const int nrows = g.nb[1] * g.nb[2];
  int nsplit     = 1;
  while (nrows * nsplit < 148 * 16 * 4 && g.nb[0] / (nsplit * 2) >= 2) nsplit *= 2;
  a.nsplit = nsplit;
  a.ypi    = (g.nb[0] + nsplit - 1) / nsplit;
  a.dbg    = getenv("B200_SWEEP_DBG") ? atoi(getenv("B200_SWEEP_DBG")) : 0;
  if (getenv("B200_SWEEP_NSPLIT")) {
    a.nsplit = atoi(getenv("B200_SWEEP_NSPLIT"));
    a.ypi    = (g.nb[0] + a.nsplit - 1) / a.nsplit;
  }
  auto kern = k_spread3_sweep<NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)CF::BYTES);
  if (e != cudaSuccess) return e;
  kern<<<nrows * a.nsplit, 32, CF::BYTES, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_spread3_sweep(int ns, const PointSet<float> &pts, const GridGeom<float> &g,
                                 int nc, const float *coef, const float2 *c_in, float2 *fw,
                                 cudaStream_t st) {
  switch (ns) {
  case 6: return launch_spread_ns<6>(pts, g, nc, coef, c_in, fw, st);
  case 7: return launch_spread_ns<7>(pts, g, nc, coef, c_in, fw, st);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_interp3_sweep(int, const PointSet<float> &, const GridGeom<float> &, int,
                                 const float *, float2 *, const float2 *, cudaStream_t) {
  return cudaErrorInvalidValue;
}

}  // namespace b200
