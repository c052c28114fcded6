// Shared definitions for the PeRCNN B200 kernels: geometry, the digested parameter block that lives
// in __constant__ memory, and small device helpers.  Everything here is sm_100a-only by design.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/percnn_b200.h"

namespace percnn {

// ---------------------------------------------------------------------------------------------
// Digested parameter block ("prep block").  percnn_params_load() runs prep kernels that turn the
// raw state_dict packing into this layout (in fp64, rounded once to the plan dtype) and copies it
// device-to-device into a __constant__ slot, so the step kernels read every coefficient through the
// uniform datapath (LDCU -> UR operands of FFMA/FFMA2) and spend no vector registers on weights.
// ---------------------------------------------------------------------------------------------
enum PrepIndex : int {
  P_ALPHA = 0,     // [2]  effective diffusion coefficient per field: DA | mu_up*sigmoid(CA)  (GS2D:115)
  P_DT = 2,        // [1]
  P_LAP_C0 = 3,    // [1]  centre tap of W_laplace.weight (already / dx^2, GS2D:66)
  P_LAP_AX = 4,    // [3][4] taps at offsets -2,-1,+1,+2 along axis a (a = 0 slowest); 2-D uses a = 0,1
  P_POLY = 16,     // [2][10] folded cubic of the 1x1 Pi-block: c00 c10 c01 c20 c11 c02 c30 c21 c12 c03
  P_DPOLY = 36,    // [4][6] d/du, d/dv of both cubics over monomials 1 u v u^2 uv v^2: (Ru,u) (Ru,v) (Rv,u) (Rv,v)
  P_PHYS = 16,     // Burgers: C1_u C2_u C1_v C2_v | dx taps[4] | dy taps[4]   (taps / dx, BUR3:78-80)
                   // LO:      C1..C5_u | C1..C5_v | C6_v
  P_BRANCH = 64,   // raw 1x1 weights for PERCNN_FLAG_EVAL_BRANCH: per field W1[hc][2] b1[hc] W2.. W3.. W4[hc] b4
  P_LAPT = 392,    // [13] Laplacian taps mirrored along every axis (the adjoint stencil): c0, then [3][4]
  P_SIZE = 408
};
constexpr int kMaxHidden = 16;
static_assert(P_BRANCH + 2 * (10 * kMaxHidden + 1) <= P_LAPT, "the EVAL_BRANCH weight copy must not reach the mirrored taps");
static_assert(P_LAPT + 13 <= P_SIZE, "prep block too small");
constexpr int kPrepSlots = 6;

struct PrepBlock {
  float f[P_SIZE];
  double d[P_SIZE];
};

// The library is several translation units compiled in parallel; each holds its own copy of the block
// (`static`: no cross-TU symbol) and percnn_params_load fills every copy (plan.h: *_load_prep).
static __constant__ PrepBlock c_prep[kPrepSlots];

// number of reduction quantities the adjoint kernels produce per step
constexpr int kRedPiK1 = 22;     // S_Lu S_Lv | M^u_ab[10] | M^v_ab[10]
constexpr int kRedBurgers = 6;   // nu_u nu_v C1_u C2_u C1_v C2_v
constexpr int kRedLO = 13;       // nu_u nu_v C1..C5_u C1..C5_v C6_v
constexpr int kRedMaxSmall = 24;

// ---------------------------------------------------------------------------------------------
// Geometry of one state buffer as a kernel sees it.
// ---------------------------------------------------------------------------------------------
struct Geom {
  int ndim;        // 2 | 3
  int D, H, W;     // interior extents (2-D: D = 1)
  int ghost;       // slab mode: number of ghost planes (rows in 2-D) on each side of the slowest axis (0 | 2)
  int64_t plane;   // H*W  (2-D: W)
  int64_t field;   // elements between field 0 and field 1 = (D + 2*ghost) * H * W   (2-D: (H+2*ghost)*W)
};

template <typename T>
struct PrepView;
template <>
struct PrepView<float> {
  static __device__ __forceinline__ const float* get(const PrepBlock& b) { return b.f; }
};
template <>
struct PrepView<double> {
  static __device__ __forceinline__ const double* get(const PrepBlock& b) { return b.d; }
};

__device__ __forceinline__ int wrap_idx(int i, int n) {
  // valid for -n <= i < 2n
  return i < 0 ? i + n : (i >= n ? i - n : i);
}

// Last-block fold of per-block partial sums: acc[i] += scale * sum_b partials[b * nred + i].
// Every warp of the calling block takes values i = warp, warp + nwarps, ...; its lanes stride over the blocks and
// a fixed shuffle tree combines them, so the result is deterministic for a given grid and the loads are
// independent (a serial loop over hundreds of blocks costs one L2 round trip per block).
__device__ __forceinline__ void fold_partials(const double* __restrict__ partials, unsigned nblocks, int nred,
                                              double* __restrict__ acc, double scale, int acc_offset = 0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp; i < nred; i += nwarps) {
    double s = 0;
    for (unsigned b = lane; b < nblocks; b += 32) s += __ldcg(partials + size_t(b) * nred + i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) acc[acc_offset + i] += scale * s;
  }
}

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// ---------------------------------------------------------------------------------------------
// Fused data loss (percnn_data_loss_t): the gradient of  sum (h[::s] - target)^2 / N  with respect to the state
// a step's adjoint produces is  coef * (h(x) - target[x / s])  at the cells whose every coordinate is a multiple
// of s, coef = dL/dloss * 2 / N.  The adjoint kernels hold h(x) already, so they add it on the fly.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Inject {
  const T* target;     // low-res frame [2][ld][lh][lw] of the state this step differentiates; nullptr = none
  const T* gscale;     // device scalar dL/dloss, nullptr = 1
  double two_over_n;   // 2 / N
  int s;               // spatial stride
  int lh, lw;          // low-res extents of the two fastest axes
  int64_t lfield;      // low-res elements per field
};
template <typename T>
__device__ __forceinline__ T inject_coef(const Inject<T>& j) {
  if (j.target == nullptr) return T(0);
  return T(j.two_over_n * (j.gscale != nullptr ? double(__ldg(j.gscale)) : 1.0));
}
// One field of one cell: the gradient contribution itself (0 off the sampling lattice).
template <typename T>
__device__ __forceinline__ T inject_field(const Inject<T>& j, T coef, int f, int z, int y, int x, T hval) {
  if ((x % j.s) | (y % j.s) | (z % j.s)) return T(0);
  const int64_t i = (int64_t(z / j.s) * j.lh + y / j.s) * j.lw + x / j.s;
  return coef * (hval - __ldg(j.target + f * j.lfield + i));
}
// One cell (z, y, x) of the interior grid; u, v = h at that cell.
template <typename T>
__device__ __forceinline__ void inject_cell(const Inject<T>& j, T coef, int z, int y, int x, T u, T v, T& gu, T& gv) {
  if ((x % j.s) | (y % j.s) | (z % j.s)) return;
  const int64_t i = (int64_t(z / j.s) * j.lh + y / j.s) * j.lw + x / j.s;
  gu = fma_t(coef, u - __ldg(j.target + i), gu);
  gv = fma_t(coef, v - __ldg(j.target + j.lfield + i), gv);
}

}  // namespace percnn
