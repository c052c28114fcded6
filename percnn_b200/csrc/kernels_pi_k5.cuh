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
#ifndef K5_FWD_MIN_BLOCKS
#define K5_FWD_MIN_BLOCKS 4
#endif
constexpr int BX = 8, BY = 16;
constexpr int THREADS = BX * BY;
constexpr int TILE_X = BX * CELLS, TILE_Y = BY;
constexpr int SM_W = TILE_X + 8;   // 4 halo columns each side (2 used) keeps rows float4-aligned
constexpr int SM_H = TILE_Y + 4;

__host__ __device__ inline size_t smem_bytes(int hc) { return size_t(2 * SM_H * SM_W + k5_total_floats(hc) - k5_weight_floats(hc) / 2) * 4; }

__device__ __forceinline__ float2 ffma2_bs(float s, float2 w, float2 acc) { return __ffma2_rn(make_float2(s, s), w, acc); }

// One block = one 32 x 16 tile of ONE output field q = blockIdx.z: the two output fields share nothing but the input
// tile, and 2 x 512 half-size blocks spread over the 148 SMs far more evenly than 512 full ones (0.86 waves of 4
// resident blocks left a quarter of the SMs idle for a quarter of the kernel); a block also stages only its field's
// half of the weights.
__global__ void __launch_bounds__(THREADS, K5_FWD_MIN_BLOCKS) k_pi_k5_fwd(Geom g, int slot, int hc, const float* __restrict__ src,
                                                       float* __restrict__ dst, const float* __restrict__ k5w) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                       // [2][SM_H][SM_W]
  float* wsm = smem + 2 * SM_H * SM_W;      // this field's conv weights, then bias | w4 | b4 of both fields
  const float* P = c_prep[slot].f;
  const int ncp = hc / 2;
  const int x0 = blockIdx.x * TILE_X, y0 = blockIdx.y * TILE_Y;
  const int q = blockIdx.z;
  const int wq = k5_weight_floats(hc) / 2;  // conv weights of one field

  {  // weights: straight float4 copies (both regions are 16-byte aligned: wq and k5_weight_floats are multiples of 32)
    const float4* s4 = reinterpret_cast<const float4*>(k5w + q * wq);
    float4* d4 = reinterpret_cast<float4*>(wsm);
    for (int i = threadIdx.x; i < wq / 4; i += THREADS) d4[i] = __ldg(s4 + i);
    const int ntail = k5_total_floats(hc) - k5_weight_floats(hc);
    for (int i = threadIdx.x; i < ntail; i += THREADS) wsm[wq + i] = __ldg(k5w + k5_weight_floats(hc) + i);
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
  const float* bias = wsm + wq;
  const float* w4 = bias + 2 * 3 * hc;
  float R[CELLS];
#pragma unroll
  for (int j = 0; j < CELLS; ++j) R[j] = w4[2 * hc + q];
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
        const float4* wrow = reinterpret_cast<const float4*>(wsm + (((cp * 2 + f) * 5 + dy) * kK5RowFloats));
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
      R[j] = fmaf(w4p.y, pr.y, fmaf(w4p.x, pr.x, R[j]));
    }
  }
  const int y = y0 + ty;
  if (y >= g.H) return;
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
    const float res = fmaf(P[P_ALPHA + q], L, R[j]);
    dst[q * g.field + int64_t(y + g.ghost) * g.W + x] = fmaf(res, P[P_DT], c[0]);
  }
}

}  // namespace k5
}  // namespace percnn

namespace percnn {
// =====================================================================================================
// Adjoint of the 5x5 Pi-block step:
//   Gbar_i^q[c] = dt G^q W4^q[c] prod_{j != i} P_j^q[c]
//   g_f(x)      = g_add + G_f + dt alpha_f Lap^T(G_f) + sum_{q,i,c,a} W_i^q[c,f,a] Gbar_i^q[c](x - a + r)
//   dW_i^q[c,f,a] = sum_x Gbar_i^q[c](x) h_f(x + a - r) ; db_i^q[c] = sum_x Gbar_i^q[c]
//   dW4^q[c] = dt sum_x G^q (P1 P2 P3)^q[c] ; db4^q = dt sum_x G^q ; dalpha_q = dt sum_x Lap^T(G^q) q
// One block owns a 32 x 8 tile.  Per (field q, channel pair cp) stage:
//   phase 1  recompute P_1..3 on the tile + halo 2 (h staged with halo 4) and form the Gbar pairs in smem;
//   phase 2  transposed 5x5 conv: every thread gathers Gbar for 4 cells of one input field (FFMA2 over the pair);
//   phase 3  weight gradients: thread = (input field, tap row, tile row) slides along x with 15 FFMA2 per cell,
//            slices are summed in fixed order (deterministic, no float atomics).
// Block partial sums go to global memory in the raw parameter packing; k5_reduce_partials folds them in fp64.
// =====================================================================================================
namespace k5 {

constexpr int BT_X = 32, BT_Y = 8;
constexpr int R2_W = BT_X + 4, R2_H = BT_Y + 4;
constexpr int R4_W = BT_X + 8, R4_H = BT_Y + 8;
constexpr int BTHREADS = 128;
constexpr int R2_CELLS = R2_W * R2_H;
constexpr int SLICE_FLOATS = 2 * 5 * 5 * 3 * 2;   // (f, dy, dx, i, channel of the pair) = 300

__host__ __device__ inline size_t bwd_smem_floats(int hc, int nparams) {
  (void)nparams;   // every parameter sum is produced by exactly one (q, cp) stage, so the block writes it straight to its
                   // global partial vector: no block-wide accumulator array (it cost 19.7 KB and the 4th resident block)
  return size_t(2 * R4_H * R4_W + 2 * R2_H * R2_W + 3 * R2_CELLS * 2 + k5_total_floats(hc) + BT_Y * SLICE_FLOATS + 64);
}

__global__ void __launch_bounds__(BTHREADS) k_pi_k5_bwd(Geom g, int slot, int hc, int nparams, const float* __restrict__ h,
                                                        const float* __restrict__ gout, const float* __restrict__ gadd,
                                                        float* __restrict__ gin, const float* __restrict__ k5w,
                                                        float* __restrict__ partials) {
  extern __shared__ __align__(16) float smem[];
  float* sh = smem;                                   // [2][R4_H][R4_W]
  float* sg = sh + 2 * R4_H * R4_W;                   // [2][R2_H][R2_W]
  float2* sgbar = reinterpret_cast<float2*>(sg + 2 * R2_H * R2_W);   // [3][R2_CELLS]
  float* wsm = reinterpret_cast<float*>(sgbar + 3 * R2_CELLS);
  float* sslice = wsm + k5_total_floats(hc);          // [BT_Y][SLICE_FLOATS]
  float* sred = sslice + BT_Y * SLICE_FLOATS;         // [64] scratch for small reductions
  // this block's partial parameter sums, raw packing; each entry is written exactly once (by the stage that owns it)
  float* out = partials + size_t(blockIdx.y * gridDim.x + blockIdx.x) * nparams;
  const float* P = c_prep[slot].f;
  const PiPacking pk(2, 5, hc);
  const int ncp = hc / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * BT_X, y0 = blockIdx.y * BT_Y;
  const float dt = P[P_DT];

  for (int i = 2 + tid; i < 2 + 25; i += BTHREADS) out[i] = 0.f;   // the frozen Laplacian table gets no gradient
  {
    const int n4 = k5_total_floats(hc) / 4;
    const float4* s4 = reinterpret_cast<const float4*>(k5w);
    float4* d4 = reinterpret_cast<float4*>(wsm);
    for (int i = tid; i < n4; i += BTHREADS) d4[i] = __ldg(s4 + i);
    for (int i = n4 * 4 + tid; i < k5_total_floats(hc); i += BTHREADS) wsm[i] = __ldg(k5w + i);
  }
  auto wrap_row = [&](int y) -> int64_t {
    if (g.ghost) return min(max(y + g.ghost, 0), g.H + 2 * g.ghost - 1);
    y %= g.H;
    return y < 0 ? y + g.H : y;
  };
  auto wrap_col = [&](int x) -> int {
    x %= g.W;
    return x < 0 ? x + g.W : x;
  };
  for (int e = tid; e < 2 * R4_H * R4_W; e += BTHREADS) {
    const int c = e % R4_W, r = (e / R4_W) % R4_H, f = e / (R4_W * R4_H);
    sh[e] = __ldg(h + f * g.field + wrap_row(y0 + r - 4) * g.W + wrap_col(x0 + c - 4));
  }
  for (int e = tid; e < 2 * R2_H * R2_W; e += BTHREADS) {
    const int c = e % R2_W, r = (e / R2_W) % R2_H, f = e / (R2_W * R2_H);
    sg[e] = __ldg(gout + f * g.field + wrap_row(y0 + r - 2) * g.W + wrap_col(x0 + c - 2));
  }
  __syncthreads();

  // ---- roles ----
  // phase 1: quad of 4 cells in R2 (108 of 128 threads)
  const bool p1_active = tid < R2_H * (R2_W / 4);
  const int p1_row = tid / (R2_W / 4), p1_col = 4 * (tid % (R2_W / 4));
  bool p1_valid[CELLS];          // interior AND inside the domain: contributes to the parameter sums
#pragma unroll
  for (int j = 0; j < CELLS; ++j) {
    const int xl = p1_col + j - 2, yl = p1_row - 2;
    p1_valid[j] = p1_active && xl >= 0 && xl < BT_X && yl >= 0 && yl < BT_Y && (x0 + xl) < g.W && (y0 + yl) < g.H;
  }
  // phase 2: field f2, quad of interior cells
  const int f2 = tid >> 6, q2 = tid & 63;
  const int p2_row = q2 / (BT_X / 4), p2_col = 4 * (q2 % (BT_X / 4));
  float2 gacc[CELLS];
#pragma unroll
  for (int j = 0; j < CELLS; ++j) gacc[j] = make_float2(0.f, 0.f);
  // phase 3: (input field, tap row) x tile row
  const bool p3_active = tid < 10 * BT_Y;
  const int p3_role = tid % 10, p3_row = tid / 10;
  const int p3_f = p3_role / 5, p3_dy = p3_role % 5;

  const float* bias = wsm + k5_weight_floats(hc);
  const float* w4 = bias + 2 * 3 * hc;

  for (int q = 0; q < 2; ++q) {
    for (int cp = 0; cp < ncp; ++cp) {
      // ================= phase 1 =================
      float red8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) red8[i] = 0.f;
      if (p1_active) {
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
            const float4* trow = reinterpret_cast<const float4*>(sh + (f * R4_H + p1_row + dy) * R4_W + p1_col);
            const float4 d0 = trow[0], d1 = trow[1];
            const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            const float4* wrow = reinterpret_cast<const float4*>(wsm + ((((q * ncp + cp) * 2 + f) * 5 + dy) * kK5RowFloats));
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
                for (int j = 0; j < CELLS; ++j) acc[i][j] = ffma2_bs(d[dx + j], wp, acc[i][j]);
              }
          }
        }
        const float2 w4p = *reinterpret_cast<const float2*>(w4 + q * hc + 2 * cp);
#pragma unroll
        for (int j = 0; j < CELLS; ++j) {
          const float Gd = dt * sg[(q * R2_H + p1_row) * R2_W + p1_col + j];
          const float2 s = make_float2(Gd * w4p.x, Gd * w4p.y);
          const float2 p12 = __fmul2_rn(acc[0][j], acc[1][j]);
          const float2 g1 = __fmul2_rn(s, __fmul2_rn(acc[1][j], acc[2][j]));
          const float2 g2 = __fmul2_rn(s, __fmul2_rn(acc[0][j], acc[2][j]));
          const float2 g3 = __fmul2_rn(s, p12);
          const int cell = p1_row * R2_W + p1_col + j;
          sgbar[0 * R2_CELLS + cell] = g1;
          sgbar[1 * R2_CELLS + cell] = g2;
          sgbar[2 * R2_CELLS + cell] = g3;
          if (p1_valid[j]) {
            const float2 p123 = __fmul2_rn(p12, acc[2][j]);
            red8[0] = fmaf(Gd, p123.x, red8[0]);
            red8[1] = fmaf(Gd, p123.y, red8[1]);
            red8[2] += g1.x; red8[3] += g1.y;
            red8[4] += g2.x; red8[5] += g2.y;
            red8[6] += g3.x; red8[7] += g3.y;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = red8[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sred[warp * 8 + i] = v;
      }
      __syncthreads();   // Gbar pairs and the per-warp sums are visible
      if (tid < 8) {
        const float v = (sred[tid] + sred[8 + tid]) + (sred[16 + tid] + sred[24 + tid]);
        const int c = 2 * cp + (tid & 1);
        const int which = tid >> 1;   // 0: W4, 1..3: bias of conv which-1
        const int idx = which == 0 ? pk.w4(q) + c : pk.b(q, which - 1) + c;
        out[idx] = v;
      }
      // ================= phase 2: transposed conv into gacc (field f2) =================
      {
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
          const float4* wrow = reinterpret_cast<const float4*>(wsm + ((((q * ncp + cp) * 2 + f2) * 5 + dy) * kK5RowFloats));
          float wr[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 t = wrow[v];
            wr[4 * v + 0] = t.x; wr[4 * v + 1] = t.y; wr[4 * v + 2] = t.z; wr[4 * v + 3] = t.w;
          }
          // Gbar row of R2 that tap row dy reads: (p2_row + 2) - (dy - 2) = p2_row + 4 - dy
          const int r2 = p2_row + 4 - dy;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float4* grow = reinterpret_cast<const float4*>(sgbar + i * R2_CELLS + r2 * R2_W + p2_col);
            float2 gb[8];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const float4 t = grow[v];
              gb[2 * v] = make_float2(t.x, t.y);
              gb[2 * v + 1] = make_float2(t.z, t.w);
            }
            // cell j, tap dx reads R2 column (p2_col + j + 2) - (dx - 2) = p2_col + j + 4 - dx
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
              const float2 wp = make_float2(wr[(dx * 3 + i) * 2], wr[(dx * 3 + i) * 2 + 1]);
#pragma unroll
              for (int j = 0; j < CELLS; ++j) gacc[j] = __ffma2_rn(wp, gb[j + 4 - dx], gacc[j]);
            }
          }
        }
      }
      // ================= phase 3: weight gradients =================
      if (p3_active) {
        float2 wacc[3][5];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) wacc[i][dx] = make_float2(0.f, 0.f);
        const int gy = y0 + p3_row;
        if (gy < g.H) {
          // h_f(y + dy - 2, x + dx - 2): R4 row p3_row + 4 + dy - 2, R4 column x + 4 + dx - 2
          const float* hrow = sh + (p3_f * R4_H + p3_row + 2 + p3_dy) * R4_W + 2;
          float win[5] = {hrow[0], hrow[1], hrow[2], hrow[3], 0.f};
          const int nx = min(BT_X, g.W - x0);
          for (int x = 0; x < nx; ++x) {
            win[4] = hrow[x + 4];
            const int cell = (p3_row + 2) * R2_W + x + 2;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float2 gb = sgbar[i * R2_CELLS + cell];
#pragma unroll
              for (int dx = 0; dx < 5; ++dx) wacc[i][dx] = ffma2_bs(win[dx], gb, wacc[i][dx]);
            }
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) win[dx] = win[dx + 1];
          }
        }
        float* sl = sslice + p3_row * SLICE_FLOATS + ((p3_f * 5 + p3_dy) * 5) * 6;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            sl[(dx * 3 + i) * 2 + 0] = wacc[i][dx].x;
            sl[(dx * 3 + i) * 2 + 1] = wacc[i][dx].y;
          }
      }
      __syncthreads();   // slices complete; Gbar no longer needed
      for (int e = tid; e < SLICE_FLOATS; e += BTHREADS) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < BT_Y; ++r) v += sslice[r * SLICE_FLOATS + e];
        const int ch = e & 1, i = (e >> 1) % 3, dx = (e / 6) % 5, dy = (e / 30) % 5, f = e / 150;
        const int c = 2 * cp + ch;
        out[pk.w(q, i) + ((c * 2 + f) * 5 + dy) * 5 + dx] = v;
      }
      __syncthreads();
    }
  }
  // ---- epilogue: Laplacian adjoint, g_in, alpha and b4 sums ----
  float ra = 0.f, rb = 0.f;   // dt * sum Lap^T(G_f) h_f  and  dt * sum G_f   for field f2
  {
    const int gy = y0 + p2_row;
#pragma unroll
    for (int j = 0; j < CELLS; ++j) {
      const int gx = x0 + p2_col + j;
      if (gy < g.H && gx < g.W) {
        const float* c = sg + (f2 * R2_H + p2_row + 2) * R2_W + p2_col + j + 2;
        float LT = P[P_LAP_C0] * c[0];
        LT = fmaf(P[P_LAP_AX + 3], c[-2 * R2_W], LT);
        LT = fmaf(P[P_LAP_AX + 2], c[-1 * R2_W], LT);
        LT = fmaf(P[P_LAP_AX + 1], c[1 * R2_W], LT);
        LT = fmaf(P[P_LAP_AX + 0], c[2 * R2_W], LT);
        LT = fmaf(P[P_LAP_AX + 7], c[-2], LT);
        LT = fmaf(P[P_LAP_AX + 6], c[-1], LT);
        LT = fmaf(P[P_LAP_AX + 5], c[1], LT);
        LT = fmaf(P[P_LAP_AX + 4], c[2], LT);
        const float hval = sh[(f2 * R4_H + p2_row + 4) * R4_W + p2_col + j + 4];
        const int64_t o = f2 * g.field + int64_t(gy + g.ghost) * g.W + gx;
        float v = c[0] + fmaf(P[P_ALPHA + f2], dt * LT, gacc[j].x + gacc[j].y);
        if (gadd != nullptr) v += __ldg(gadd + o);
        gin[o] = v;
        ra = fmaf(dt * LT, hval, ra);
        rb = fmaf(dt, c[0], rb);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ra += __shfl_down_sync(0xffffffffu, ra, off);
    rb += __shfl_down_sync(0xffffffffu, rb, off);
  }
  if (lane == 0) {
    sred[32 + warp * 2] = ra;
    sred[32 + warp * 2 + 1] = rb;
  }
  __syncthreads();
  if (tid < 2) {   // warps 0,1 hold field 0; warps 2,3 field 1
    out[tid] = sred[32 + (2 * tid) * 2] + sred[32 + (2 * tid + 1) * 2];                    // alpha_f
    out[pk.w4(tid) + hc] = sred[32 + (2 * tid) * 2 + 1] + sred[32 + (2 * tid + 1) * 2 + 1];  // b4_f
  }
}

// acc[i] += sum over blocks (fixed order, fp64) of partials[b][i]
__global__ void k5_reduce_partials(const float* __restrict__ partials, int nblocks, int nparams, double* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nparams) return;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;   // four independent chains keep the loads in flight; order is fixed
  int b = 0;
  for (; b + 4 <= nblocks; b += 4) {
    s0 += double(__ldg(partials + size_t(b) * nparams + i));
    s1 += double(__ldg(partials + size_t(b + 1) * nparams + i));
    s2 += double(__ldg(partials + size_t(b + 2) * nparams + i));
    s3 += double(__ldg(partials + size_t(b + 3) * nparams + i));
  }
  for (; b < nblocks; ++b) s0 += double(__ldg(partials + size_t(b) * nparams + i));
  acc[i] += (s0 + s1) + (s2 + s3);
}

// raw-packing accumulators -> gradients (only CA/CB need the sigmoid chain rule; the frozen Laplacian gets 0)
__global__ void k5_finish(const float* __restrict__ raw, const double* __restrict__ acc, PrepDesc d, int nparams,
                          float* __restrict__ grads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nparams) return;
  double gval = acc[i];
  if (i < 2 && d.coef_mode == PERCNN_COEF_SIGMOID) {
    const double s = 1.0 / (1.0 + exp(-double(raw[i])));
    gval *= d.mu_up * s * (1.0 - s);
  }
  if (i >= 2 && i < 2 + 25) gval = 0.0;
  grads[i] = float(gval);
}

}  // namespace k5
}  // namespace percnn
