// Tube-sweep spread / interp kernels for 3D float (see sweep3d.cuh for the design).
#include "sweep3d.cuh"

#include <limits.h>

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
  static constexpr int HL    = NS / 2;                 // cells left of the bin a stencil can reach
  static constexpr int ROW   = kBinX + NS;             // cells of one register row
  static constexpr int PAD   = HL & 1;                 // tile x origin 16*i1-HL-PAD is even
  static constexpr int PITCH = (ROW + PAD + 1) & ~1;   // cells of one tile row (even)
  static constexpr int SP    = 26;                     // staging row pitch in cells (bank spread)
  static constexpr int ZR    = NS + 1;                 // z rows one z cell's stencils can touch
  static constexpr int RZ    = 2;                      // z rows per lane
  static constexpr int BQ    = (ZR + RZ - 1) / RZ;     // lane groups along z
  static constexpr int ZS    = RZ * BQ;                // z rows staged per warp (>= ZR)
  static constexpr int ZT    = kBinZ + NS;             // z rows of the block tile
  static constexpr int NW    = kBinZ;                  // warps per block = z cells per bin
  static constexpr int NT    = NW * 32;
  static constexpr int NSLOT = kBinY;                  // rows leaving the window per step
  static constexpr int NBUCK = NW * (kBinY + 1);       // (z cell, y stencil start) buckets
  static constexpr int CH    = 256;                    // points per chunk
  static constexpr int PPT   = CH / NT;
  static constexpr int RECW  = 28;                     // words per point record
  // record: [0,NS) phi_x  [7] x offset  [8,8+NS) phi_y  [16,16+ZS) phi_z window  [24,25] c
  static constexpr size_t STAGE_BYTES = (size_t)NW * NSLOT * ZS * SP * sizeof(float2);
  static constexpr size_t REC_BYTES   = (size_t)CH * RECW * sizeof(float);
  static constexpr size_t ORD_BYTES   = (size_t)CH * sizeof(uint16_t);
  static constexpr size_t BYTES       = STAGE_BYTES + REC_BYTES + ORD_BYTES;
  static_assert(PITCH <= SP, "staging pitch");
  static_assert(NS * BQ <= 32, "lane map");
  static_assert(NS <= 7 && ZS <= 8, "record layout");
};

template<int NS> struct SweepArgs {
  PointSet<float> pts;
  GridGeom<float> g;
  WindowTable<float, NS> tab;
  const float2 *c_in;
  float2 *c_out;
  float2 *fw;
  int nsplit;  // y ranges per tube
  int ypi;     // y bins per item
};

__device__ __forceinline__ int pmod(int v, int n) {
  int r = v % n;
  return r < 0 ? r + n : r;
}

// one point's contribution for a compile-time x offset: NS packed FMAs per owned row
template<int NS, int XO>
__device__ __forceinline__ void row_update(float2 (&acc)[SweepCfg<NS>::RZ][SweepCfg<NS>::ROW],
                                           const float (&kx)[8], const float2 (&w)[SweepCfg<NS>::RZ]) {
#pragma unroll
  for (int r = 0; r < SweepCfg<NS>::RZ; ++r)
#pragma unroll
    for (int t = 0; t < NS; ++t) acc[r][XO + t] = ffma2_s(kx[t], w[r], acc[r][XO + t]);
}

template<int NS>
__device__ __forceinline__ void row_update_switch(
    int xo, float2 (&acc)[SweepCfg<NS>::RZ][SweepCfg<NS>::ROW], const float (&kx)[8],
    const float2 (&w)[SweepCfg<NS>::RZ]) {
  switch (xo) {
#define B200_XO(k) case k: row_update<NS, k>(acc, kx, w); break;
    B200_XO(0) B200_XO(1) B200_XO(2) B200_XO(3) B200_XO(4) B200_XO(5) B200_XO(6) B200_XO(7)
    B200_XO(8) B200_XO(9) B200_XO(10) B200_XO(11) B200_XO(12) B200_XO(13) B200_XO(14)
    B200_XO(15) B200_XO(16)
#undef B200_XO
  default: break;
  }
}

// Phase A for one point: fold, stencil starts, windows, strength -> record; returns its bucket.
template<int NS>
__device__ __forceinline__ int make_record(const SweepArgs<NS> &a, uint32_t q, float *rec, int i1,
                                           int i2, int i3, bool spread) {
  using CF = SweepCfg<NS>;
  int i0;
  float x1;
  stencil_start<float, NS>(fold_rescale<float>(a.pts.xs[q], a.g.nf_t[0]), i0, x1);
  eval_window<float, NS>(a.tab, x1, rec);
  int xo = i0 - (kBinX * i1 - CF::HL);
  xo     = min(max(xo, 0), kBinX);
  rec[7] = __int_as_float(xo);
  stencil_start<float, NS>(fold_rescale<float>(a.pts.ys[q], a.g.nf_t[1]), i0, x1);
  eval_window<float, NS>(a.tab, x1, rec + 8);
  int jb = i0 - (kBinY * i2 - CF::HL);
  jb     = min(max(jb, 0), kBinY);
  const float Z = fold_rescale<float>(a.pts.zs[q], a.g.nf_t[2]);
  stencil_start<float, NS>(Z, i0, x1);
  float kz[NS];
  eval_window<float, NS>(a.tab, x1, kz);
  int wz = (int)floorf(Z) - kBinZ * i3;
  wz     = min(max(wz, 0), kBinZ - 1);
  int zsh = i0 - (kBinZ * i3 + wz - CF::HL);  // 0 or 1 (clamped: the rows must stay in the window)
  zsh     = min(max(zsh, 0), CF::ZS - NS);
#pragma unroll
  for (int r = 0; r < CF::ZS; ++r) {
    float v = 0.f;
#pragma unroll
    for (int s = 0; s <= CF::ZS - NS; ++s)
      if (r - s >= 0 && r - s < NS) v = (zsh == s) ? kz[r - s] : v;
    rec[16 + r] = v;
  }
  if (spread) {
    const float2 c = a.c_in[a.pts.sidx[q]];
    rec[24]        = c.x;
    rec[25]        = c.y;
  }
  return wz * (kBinY + 1) + jb;
}

template<int NS>
__global__ void __launch_bounds__(SweepCfg<NS>::NT, 4) k_spread3_sweep(const SweepArgs<NS> a) {
  using CF = SweepCfg<NS>;
  extern __shared__ __align__(16) unsigned char smem[];
  float2 *stage   = reinterpret_cast<float2 *>(smem);
  float *rec      = reinterpret_cast<float *>(smem + CF::STAGE_BYTES);
  uint16_t *order = reinterpret_cast<uint16_t *>(smem + CF::STAGE_BYTES + CF::REC_BYTES);
  __shared__ int s_cnt[32], s_base[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int la = lane / CF::BQ, bq = lane % CF::BQ;  // la >= NS: spare lane
  const bool live = la < NS;
  const int nb1 = a.g.nb[0], nb2 = a.g.nb[1];
  const int tube = blockIdx.x / a.nsplit, part = blockIdx.x % a.nsplit;
  const int i1 = tube % nb1, i3 = tube / nb1;
  const int yb0 = part * a.ypi, yb1 = min(nb2, yb0 + a.ypi);
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];
  const int x0t = kBinX * i1 - CF::HL - CF::PAD;  // tile origin (even)
  const int z0t = kBinZ * i3 - CF::HL;

  // staging starts zeroed: the pad cells of every row are never written afterwards
  for (int i = tid; i < (int)(CF::STAGE_BYTES / sizeof(float2)); i += CF::NT)
    stage[i] = float2{0.f, 0.f};

  float2 acc[CF::RZ][CF::ROW];
#pragma unroll
  for (int r = 0; r < CF::RZ; ++r)
#pragma unroll
    for (int c = 0; c < CF::ROW; ++c) acc[r][c] = float2{0.f, 0.f};
  int jw = INT_MIN;  // first y row of the register window (block-uniform); INT_MIN = none

  float2 *my_stage = stage + ((size_t)(warp * CF::NSLOT) * CF::ZS + bq * CF::RZ) * CF::SP + CF::PAD;

  // park this lane's rows in staging slot `slot` and clear them
  auto stage_rows = [&](int slot) {
    float2 *dst = my_stage + (size_t)slot * CF::ZS * CF::SP;
#pragma unroll
    for (int r = 0; r < CF::RZ; ++r)
#pragma unroll
      for (int c = 0; c < CF::ROW; ++c) {
        dst[r * CF::SP + c] = acc[r][c];
        acc[r][c]           = float2{0.f, 0.f};
      }
  };
  // sum the warps' staged rows [yfirst, yfirst+cnt) and add them to the fine grid
  auto merge = [&](int yfirst, int cnt) {
    constexpr int NXP = CF::PITCH / 2;
    for (int idx = tid; idx < cnt * CF::ZT * NXP; idx += CF::NT) {
      const int xp = idx % NXP, z = (idx / NXP) % CF::ZT, s = idx / (NXP * CF::ZT);
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < CF::NW; ++w) {
        const int zr = z - w;
        if (zr >= 0 && zr < CF::ZR) {
          const float4 v = *reinterpret_cast<const float4 *>(
              stage + ((size_t)(w * CF::NSLOT + s) * CF::ZS + zr) * CF::SP + 2 * xp);
          sum.x += v.x, sum.y += v.y, sum.z += v.z, sum.w += v.w;
        }
      }
      if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f) {
        const int gx = wrap_index(x0t + 2 * xp, nf1), gy = wrap_index(yfirst + s, nf2),
                  gz = wrap_index(z0t + z, nf3);
        atomicAdd(reinterpret_cast<float4 *>(a.fw + ((size_t)gz * nf2 + gy) * (size_t)nf1 + gx),
                  sum);
      }
    }
  };
  // write out every live row of the window (start of an irregular step, end of the item)
  auto flush_all = [&]() {
    if (jw == INT_MIN) return;
    const int rel = live ? pmod(la - jw, NS) : -1;
    for (int r0 = 0; r0 < NS; r0 += CF::NSLOT) {
      const int cnt = min(CF::NSLOT, NS - r0);
      if (rel >= r0 && rel < r0 + cnt) stage_rows(rel - r0);
      __syncthreads();
      merge(jw + r0, cnt);
      __syncthreads();
    }
    jw = INT_MIN;
  };

  __syncthreads();
  for (int i2 = yb0; i2 < yb1; ++i2) {
    const uint32_t bin = (uint32_t)i1 + (uint32_t)nb1 * ((uint32_t)i2 + (uint32_t)nb2 * (uint32_t)i3);
    const uint32_t qs = a.pts.binstart[bin], qe = a.pts.binstart[bin + 1];
    if (qs == qe && jw == INT_MIN) continue;  // nothing in flight, nothing to do
    const int target = kBinY * i2 - CF::HL;
    uint32_t qa = qs;
    do {
      const uint32_t qb = min(qe, qa + CF::CH);
      // ---------------- phase A: records + buckets
      if (tid < 32) s_cnt[tid] = 0;
      __syncthreads();
      int mybucket[CF::PPT], mypos[CF::PPT];
#pragma unroll
      for (int k = 0; k < CF::PPT; ++k) {
        const uint32_t slot = tid + k * CF::NT, q = qa + slot;
        mybucket[k]         = -1;
        if (q < qb) {
          mybucket[k] = make_record<NS>(a, q, rec + slot * CF::RECW, i1, i2, i3, true);
          mypos[k]    = atomicAdd(&s_cnt[mybucket[k]], 1);
        }
      }
      __syncthreads();
      if (tid < 32) {
        const int v = s_cnt[tid];
        int incl    = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += up;
        }
        s_base[tid] = incl - v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CF::PPT; ++k)
        if (mybucket[k] >= 0) order[s_base[mybucket[k]] + mypos[k]] = (uint16_t)(tid + k * CF::NT);
      // ---------------- window to the start of this bin (irregular only)
      if (jw != target) {
        flush_all();  // contains the barriers
        jw = target;
      }
      __syncthreads();
      // ---------------- phase B: my z cell's points, in y-stencil-start order
      for (int jb = 0; jb <= kBinY; ++jb) {
        const int j0 = target + jb;
        if (jb > 0 && la == pmod(j0 - 1, NS)) stage_rows(jb - 1);  // row j0-1 leaves the window
        const int ta  = live ? pmod(la - j0, NS) : 0;
        const int b   = warp * (kBinY + 1) + jb;
        const int pb  = s_base[b], pe = pb + s_cnt[b];
        for (int p = pb; p < pe; ++p) {
          const float *rp = rec + (int)order[p] * CF::RECW;
          const float4 k0 = *reinterpret_cast<const float4 *>(rp);
          const float4 k1 = *reinterpret_cast<const float4 *>(rp + 4);
          const float2 c  = *reinterpret_cast<const float2 *>(rp + 24);
          const float wy  = rp[8 + ta];
          const float2 fz = *reinterpret_cast<const float2 *>(rp + 16 + bq * CF::RZ);
          const float kx[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, 0.f};
          float2 w[CF::RZ];
          w[0] = fmul2_s(wy * fz.x, c);
          w[1] = fmul2_s(wy * fz.y, c);
          row_update_switch<NS>(__float_as_int(k1.w), acc, kx, w);
        }
      }
      jw = target + kBinY;
      __syncthreads();
      merge(target, CF::NSLOT);
      __syncthreads();
      qa = qb;
    } while (qa < qe);
  }
  flush_all();
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
    for (int j = 0; j < NS; ++j) {
      const int src       = k - (rows - nc);
      a.tab.c[k * NS + j] = src >= 0 ? coef[src * NS + j] : 0.f;
    }
  a.c_in  = c_in;
  a.c_out = nullptr;
  a.fw    = fw;
  // cut every tube into y ranges so that there are enough blocks to balance 148 SMs x 4
  const int tubes = g.nb[0] * g.nb[2];
  int nsplit      = 1;
  while (tubes * nsplit < 148 * 4 * 12 && g.nb[1] / (nsplit * 2) >= 8) nsplit *= 2;
  a.nsplit = nsplit;
  a.ypi    = (g.nb[1] + nsplit - 1) / nsplit;
  auto kern = k_spread3_sweep<NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)CF::BYTES);
  if (e != cudaSuccess) return e;
  kern<<<tubes * nsplit, CF::NT, CF::BYTES, st>>>(a);
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
