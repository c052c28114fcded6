// Pi-block cell with 5x5 branch convolutions, hc = 16 (BUR1:142-178, LO1:142-171), fp32, 2-D.
//
// 4 800 FMA per cell: this variant is bound by the FP32 pipe, not by HBM.  Each thread owns 4
// x-adjacent cells; the six 5x5 convs are accumulated two output channels at a time as packed
// FFMA2 (accumulator pair = channels 2cp, 2cp+1; the input cell is the broadcast scalar operand,
// the weight pair comes from shared memory), so no operand ever needs re-alignment.
#pragma once
#include "kernels_prep.cuh"
#include "point_ops.cuh"

namespace percnn {


namespace k5 {

constexpr int CELLS = 4;
constexpr int BX = 8, BY = 16;
constexpr int THREADS = BX * BY;
constexpr int TILE_X = BX * CELLS, TILE_Y = BY;
constexpr int SM_W = TILE_X + 8;   // 4 halo columns each side (2 used) keeps rows float4-aligned
constexpr int SM_H = TILE_Y + 4;

__host__ __device__ inline size_t smem_bytes(int hc) { return size_t(2 * SM_H * SM_W + k5_total_floats(hc)) * 4; }

__device__ __forceinline__ float2 ffma2_bs(float s, float2 w, float2 acc) { return __ffma2_rn(make_float2(s, s), w, acc); }

__global__ void __launch_bounds__(THREADS) k_pi_k5_fwd(Geom g, int slot, int hc, const float* __restrict__ src,
                                                       float* __restrict__ dst, const float* __restrict__ k5w) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                       // [2][SM_H][SM_W]
  float* wsm = smem + 2 * SM_H * SM_W;      // repacked weights
  const float* P = c_prep[slot].f;
  const int ncp = hc / 2;
  const int x0 = blockIdx.x * TILE_X, y0 = blockIdx.y * TILE_Y;

  {  // weights: straight float4 copy
    const int n4 = k5_total_floats(hc) / 4;
    const float4* s4 = reinterpret_cast<const float4*>(k5w);
    float4* d4 = reinterpret_cast<float4*>(wsm);
    for (int i = threadIdx.x; i < n4; i += THREADS) d4[i] = __ldg(s4 + i);
    for (int i = n4 * 4 + threadIdx.x; i < k5_total_floats(hc); i += THREADS) wsm[i] = __ldg(k5w + i);
  }
  // state tile with periodic halo (rows: ghost-aware in slab mode)
  for (int e = threadIdx.x; e < 2 * SM_H * SM_W; e += THREADS) {
    const int c = e % SM_W;
    const int r = (e / SM_W) % SM_H;
    const int f = e / (SM_W * SM_H);
    int x = x0 + c - 4, y = y0 + r - 2;
    x %= g.W;
    if (x < 0) x += g.W;
    int64_t row;
    if (g.ghost) {
      row = min(max(y + g.ghost, 0), g.H + 2 * g.ghost - 1);
    } else {
      y %= g.H;
      if (y < 0) y += g.H;
      row = y;
    }
    tile[e] = __ldg(src + f * g.field + row * g.W + x);
  }
  __syncthreads();

  const int tx = threadIdx.x % BX, ty = threadIdx.x / BX;
  const float* bias = wsm + k5_weight_floats(hc);
  const float* w4 = bias + 2 * 3 * hc;
  float R[2][CELLS];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int j = 0; j < CELLS; ++j) R[q][j] = w4[2 * hc + q];
    for (int cp = 0; cp < ncp; ++cp) {
      float2 acc[3][CELLS];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float2 b = *reinterpret_cast<const float2*>(bias + (q * 3 + i) * hc + 2 * cp);
#pragma unroll
        for (int j = 0; j < CELLS; ++j) acc[i][j] = b;
      }
#pragma unroll
      for (int f = 0; f < 2; ++f) {
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
          const float4* trow = reinterpret_cast<const float4*>(tile + (f * SM_H + ty + dy) * SM_W + 4 * tx);
          const float4 d0 = trow[0], d1 = trow[1], d2 = trow[2];
          const float d[12] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w};
          const float4* wrow =
              reinterpret_cast<const float4*>(wsm + ((((q * ncp + cp) * 2 + f) * 5 + dy) * kK5RowFloats));
          float wr[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 t = wrow[v];
            wr[4 * v + 0] = t.x; wr[4 * v + 1] = t.y; wr[4 * v + 2] = t.z; wr[4 * v + 3] = t.w;
          }
#pragma unroll
          for (int dx = 0; dx < 5; ++dx)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float2 wp = make_float2(wr[(dx * 3 + i) * 2], wr[(dx * 3 + i) * 2 + 1]);
#pragma unroll
              for (int j = 0; j < CELLS; ++j) acc[i][j] = ffma2_bs(d[2 + dx + j], wp, acc[i][j]);
            }
        }
      }
      const float2 w4p = *reinterpret_cast<const float2*>(w4 + q * hc + 2 * cp);
#pragma unroll
      for (int j = 0; j < CELLS; ++j) {
        const float2 pr = __fmul2_rn(__fmul2_rn(acc[0][j], acc[1][j]), acc[2][j]);
        R[q][j] = fmaf(w4p.y, pr.y, fmaf(w4p.x, pr.x, R[q][j]));
      }
    }
  }
  const int y = y0 + ty;
  if (y >= g.H) return;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float* t = tile + (q * SM_H + ty + 2) * SM_W + 4 * tx + 4;
#pragma unroll
    for (int j = 0; j < CELLS; ++j) {
      const int x = x0 + 4 * tx + j;
      if (x >= g.W) continue;
      const float* c = t + j;
      float L = P[P_LAP_C0] * c[0];
      L = fmaf(P[P_LAP_AX + 0], c[-2 * SM_W], L);
      L = fmaf(P[P_LAP_AX + 1], c[-1 * SM_W], L);
      L = fmaf(P[P_LAP_AX + 2], c[1 * SM_W], L);
      L = fmaf(P[P_LAP_AX + 3], c[2 * SM_W], L);
      L = fmaf(P[P_LAP_AX + 4], c[-2], L);
      L = fmaf(P[P_LAP_AX + 5], c[-1], L);
      L = fmaf(P[P_LAP_AX + 6], c[1], L);
      L = fmaf(P[P_LAP_AX + 7], c[2], L);
      const float res = fmaf(P[P_ALPHA + q], L, R[q][j]);
      dst[q * g.field + int64_t(y + g.ghost) * g.W + x] = fmaf(res, P[P_DT], c[0]);
    }
  }
}

}  // namespace k5
}  // namespace percnn
