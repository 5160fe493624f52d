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
  static constexpr int HL    = NS / 2;                 // cells left of the bin a stencil can reach
  static constexpr int ROW   = kBinX + NS;             // cells of one register row
  static constexpr int PAD   = HL & 1;                 // tile x origin 16*i1-HL-PAD is even
  static constexpr int PITCH = (ROW + PAD + 1) & ~1;   // cells of one tile row (even)
  static constexpr int SP    = 26;                     // staging row pitch in cells (bank spread)
  static constexpr int ZT    = kBinZ + NS;             // z rows of the tube tile
  static constexpr int BQ    = 4;                      // lane groups along z
  static constexpr int RZ    = (ZT + BQ - 1) / BQ;     // z rows per lane: z = bq + BQ*m
  static constexpr int ZS    = RZ * BQ;                // z row slots (>= ZT)
  static constexpr int NJB   = kBinY + 1;              // y stencil starts inside one bin
  static constexpr int CH    = 64;                     // points per chunk (records in flight)
  static constexpr int LCAP  = 512;                    // points per sorted batch of one bin
  static constexpr int RECW  = 36;                     // words per point record
  // record: [0,NS) phi_x  [7] x offset | jb<<8  [8,8+NS) phi_y  [16+4*bq+m] phi_z of row slot
  //         (bq,m)  [32,33] strength
  static constexpr size_t STAGE_BYTES = (size_t)ZS * SP * sizeof(float2);
  static constexpr size_t REC_BYTES   = (size_t)CH * RECW * sizeof(float);
  static constexpr size_t LIST_BYTES  = (size_t)LCAP * sizeof(uint16_t);
  static constexpr size_t BYTES       = STAGE_BYTES + REC_BYTES + LIST_BYTES + 64;
  static_assert(PITCH <= SP, "staging pitch");
  static_assert(NS * BQ <= 32, "lane map");
  static_assert(NS <= 7 && RZ <= 3, "record layout");
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
  int nsplit;  // y ranges per tube
  int ypi;     // y bins per item
  int dbg;     // experiment switches (0 in production)
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
  default: __builtin_unreachable();
  }
}

// y stencil start of a point relative to the first one possible in y bin i2, in [0, kBinY]
template<int NS> __device__ __forceinline__ int y_bucket(float y, float nf2_t, int i2) {
  int j0;
  float y1;
  stencil_start<float, NS>(fold_rescale<float>(y, nf2_t), j0, y1);
  return min(max(j0 - (kBinY * i2 - SweepCfg<NS>::HL), 0), kBinY);
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
  stencil_start<float, NS>(fold_rescale<float>(__ldcs(a.pts.ys + q), a.g.nf_t[1]), i0, x1);
  eval_window2<NS>(a.tab, x1, kv);
  *reinterpret_cast<float4 *>(rec + 8)  = make_float4(kv[0], kv[1], kv[2], kv[3]);
  *reinterpret_cast<float4 *>(rec + 12) = make_float4(kv[4], kv[5], kv[6], 0.f);
  const int jb = min(max(i0 - (kBinY * i2 - CF::HL), 0), kBinY);
  stencil_start<float, NS>(fold_rescale<float>(__ldcs(a.pts.xs + q), a.g.nf_t[0]), i0, x1);
  eval_window2<NS>(a.tab, x1, kv);
  const int xo = min(max(i0 - (kBinX * i1 - CF::HL), 0), kBinX);
  kv[7]        = __int_as_float(xo | (jb << 8));
  *reinterpret_cast<float4 *>(rec)     = make_float4(kv[0], kv[1], kv[2], kv[3]);
  *reinterpret_cast<float4 *>(rec + 4) = make_float4(kv[4], kv[5], kv[6], kv[7]);
  stencil_start<float, NS>(fold_rescale<float>(__ldcs(a.pts.zs + q), a.g.nf_t[2]), i0, x1);
  eval_window2<NS>(a.tab, x1, kv);
  // z stencil start relative to the tile's first z row, in [0, kBinZ]
  const int k0 = min(max(i0 - (kBinZ * i3 - CF::HL), 0), kBinZ);
  // window value of tile row z = bq + BQ*m goes to word 16 + 4*bq + m
#pragma unroll
  for (int bq = 0; bq < CF::BQ; ++bq) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < CF::RZ; ++m) {
      const int z = bq + CF::BQ * m;
#pragma unroll
      for (int s = 0; s <= kBinZ; ++s)
        if (z - s >= 0 && z - s < NS) v[m] = (k0 == s) ? kv[z - s] : v[m];
    }
    *reinterpret_cast<float4 *>(rec + 16 + 4 * bq) = make_float4(v[0], v[1], v[2], v[3]);
  }
  *reinterpret_cast<float2 *>(rec + 32) = c;
}

__device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// One warp per work item: a tube (x bin i1, z bin i3) over the y bins [yb0, yb1).
template<int NS>
__global__ void __launch_bounds__(32, 12) k_spread3_sweep(const SweepArgs<NS> a) {
  using CF = SweepCfg<NS>;
  extern __shared__ __align__(16) unsigned char smem[];
  float2 *stage  = reinterpret_cast<float2 *>(smem);
  float *rec     = reinterpret_cast<float *>(smem + CF::STAGE_BYTES);
  uint16_t *list = reinterpret_cast<uint16_t *>(smem + CF::STAGE_BYTES + CF::REC_BYTES);
  int *cnt       = reinterpret_cast<int *>(smem + CF::STAGE_BYTES + CF::REC_BYTES + CF::LIST_BYTES);

  const int lane = threadIdx.x;
  const int la = lane / CF::BQ, bq = lane % CF::BQ;  // la >= NS: spare lane
  const int nb1 = a.g.nb[0], nb2 = a.g.nb[1];
  const int tube = blockIdx.x / a.nsplit, part = blockIdx.x % a.nsplit;
  const int i1 = tube % nb1, i3 = tube / nb1;
  const int yb0 = part * a.ypi, yb1 = min(nb2, yb0 + a.ypi);
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];
  const int x0t = kBinX * i1 - CF::HL - CF::PAD;  // tile origin (even)
  const int z0t = kBinZ * i3 - CF::HL;

  // staging starts zeroed: the pad cells of every row are never written afterwards
  for (int i = lane; i < (int)(CF::STAGE_BYTES / sizeof(float2)); i += 32)
    stage[i] = float2{0.f, 0.f};

  float2 acc[CF::RZ][CF::ROW];
#pragma unroll
  for (int r = 0; r < CF::RZ; ++r)
#pragma unroll
    for (int c = 0; c < CF::ROW; ++c) acc[r][c] = float2{0.f, 0.f};
  int jw = INT_MIN;  // first y row of the register window; INT_MIN = window empty

  const uint32_t my_stage =
      (uint32_t)__cvta_generic_to_shared(stage + (size_t)bq * CF::SP + CF::PAD);
  __syncwarp();

  // The window's first row leaves: its owners park it in the staging tile, then the whole warp
  // adds the tile row to the fine grid with 16-byte vector reductions.
  auto slide = [&]() {
    if (la == pmod(jw, NS)) {
#pragma unroll
      for (int m = 0; m < CF::RZ; ++m)
#pragma unroll
        for (int c = 0; c < CF::ROW; ++c) {
          sts64(my_stage + (uint32_t)((m * CF::BQ * CF::SP + c) * sizeof(float2)), acc[m][c]);
          acc[m][c] = float2{0.f, 0.f};
        }
    }
    __syncwarp();
    constexpr int NXP = CF::PITCH / 2;
    const int gy      = wrap_index(jw, nf2);
    for (int idx = lane; idx < CF::ZT * NXP; idx += 32) {
      const int z = idx / NXP, xp = idx - z * NXP;
      const float4 v = *reinterpret_cast<const float4 *>(stage + (size_t)z * CF::SP + 2 * xp);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        const int gx = wrap_index(x0t + 2 * xp, nf1), gz = wrap_index(z0t + z, nf3);
        if (!(a.dbg & 1))
          atomicAdd(reinterpret_cast<float4 *>(a.fw + ((size_t)gz * nf2 + gy) * (size_t)nf1 + gx), v);
      }
    }
    __syncwarp();
    ++jw;
  };
  // move the window so that it starts at row j0 (never backwards unless it is emptied first)
  auto advance_to = [&](int j0) {
    if (jw == INT_MIN) {
      jw = j0;
      return;
    }
    if (j0 < jw || j0 - jw >= NS) {  // rewind or long jump: empty the window
      for (int k = 0; k < NS; ++k) slide();
      jw = j0;
      return;
    }
    while (jw < j0) slide();
  };

  for (int i2 = yb0; i2 < yb1; ++i2) {
    const uint32_t bin = (uint32_t)i1 + (uint32_t)nb1 * ((uint32_t)i2 + (uint32_t)nb2 * (uint32_t)i3);
    const uint32_t qs = a.pts.binstart[bin], qe = a.pts.binstart[bin + 1];
    const int target = kBinY * i2 - CF::HL;
    for (uint32_t q0 = qs; q0 < qe; q0 += CF::LCAP) {
      const int n = (int)min((uint32_t)CF::LCAP, qe - q0);
      // ---- sort this batch of the bin by y stencil start (counting sort, 5 buckets)
      if (lane < 8) cnt[lane] = 0;
      __syncwarp();
      for (int k = lane; k < n; k += 32)
        atomicAdd(&cnt[y_bucket<NS>(a.pts.ys[q0 + k], a.g.nf_t[1], i2)], 1);
      __syncwarp();
      if (lane == 0) {
        int run = 0;
        for (int b = 0; b < CF::NJB; ++b) {
          const int c = cnt[b];
          cnt[b]      = run;
          run += c;
        }
      }
      __syncwarp();
      for (int k = lane; k < n; k += 32) {
        const int pos = atomicAdd(&cnt[y_bucket<NS>(a.pts.ys[q0 + k], a.g.nf_t[1], i2)], 1);
        list[pos]     = (uint16_t)k;
      }
      __syncwarp();
      for (int c0 = 0; c0 < n; c0 += CF::CH) {
        const int nc = min(CF::CH, n - c0);
        // ---- records for the next nc points in sorted order
        for (int k = lane; k < nc; k += 32)
          make_record<NS, true>(a, q0 + list[c0 + k], rec + k * CF::RECW, i1, i2, i3);
        __syncwarp();
        // ---- accumulate
        int cur_jb = -1, ta = 0;
        for (int p = 0; p < ((a.dbg & 2) ? 0 : nc); ++p) {
          const float *rp = rec + p * CF::RECW;
          const float4 k0 = *reinterpret_cast<const float4 *>(rp);
          const float4 k1 = *reinterpret_cast<const float4 *>(rp + 4);
          const int meta  = __float_as_int(k1.w);
          const int jb    = meta >> 8;
          if (jb != cur_jb) {  // warp-uniform
            cur_jb = jb;
            advance_to(target + jb);
            ta = la < NS ? pmod(la - (target + jb), NS) : 0;
          }
          const float wy  = rp[8 + ta];
          const float4 fz = *reinterpret_cast<const float4 *>(rp + 16 + 4 * bq);
          const float2 cw = fmul2_s(wy, *reinterpret_cast<const float2 *>(rp + 32));
          const float kx[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, 0.f};
          float2 w[CF::RZ];
          w[0] = fmul2_s(fz.x, cw);
          w[1] = fmul2_s(fz.y, cw);
          if (CF::RZ > 2) w[CF::RZ - 1] = fmul2_s(fz.z, cw);
          row_update_switch<NS>(meta & 0xff, acc, kx, w);
        }
        __syncwarp();
      }
    }
  }
  if (jw != INT_MIN)
    for (int k = 0; k < NS; ++k) slide();
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
  // cut every tube into y ranges so that there are enough blocks to balance 148 SMs x 4
  const int tubes = g.nb[0] * g.nb[2];
  int nsplit      = 1;
  while (tubes * nsplit < 148 * 12 * 8 && g.nb[1] / (nsplit * 2) >= 8) nsplit *= 2;
  a.nsplit = nsplit;
  a.ypi    = (g.nb[1] + nsplit - 1) / nsplit;
  a.dbg    = getenv("B200_SWEEP_DBG") ? atoi(getenv("B200_SWEEP_DBG")) : 0;
  if (getenv("B200_SWEEP_NSPLIT")) {
    a.nsplit = atoi(getenv("B200_SWEEP_NSPLIT"));
    a.ypi    = (g.nb[1] + a.nsplit - 1) / a.nsplit;
  }
  auto kern = k_spread3_sweep<NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)CF::BYTES);
  if (e != cudaSuccess) return e;
  kern<<<tubes * a.nsplit, 32, CF::BYTES, st>>>(a);
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
