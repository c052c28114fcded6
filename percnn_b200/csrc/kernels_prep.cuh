// Parameter digestion (raw state_dict packing -> prep block) and its transpose (accumulated reduction
// sums -> gradients w.r.t. the raw parameters).  Tiny single-block kernels, all arithmetic in fp64.
#pragma once
#include "common.cuh"

namespace percnn {

struct PrepDesc {
  int cell, ndim, k, hc, coef_mode, flags;
  double mu_up, dt, dx;
};

// offsets into the raw packing of a Pi cell (state_dict order: CA CB W_laplace Wh1_u.w Wh1_u.b ... Wh4_v.b)
struct PiPacking {
  int K;        // k^ndim
  int hc;
  int lap;      // offset of W_laplace.weight
  int nlap;     // 5^ndim
  int conv;     // scalars per branch conv: hc*2*K weights + hc biases
  int field;    // scalars per field: 3*conv + hc + 1
  int base;     // offset of Wh1_u.weight
  __host__ __device__ PiPacking(int ndim, int k, int hc_) {
    K = 1;
    nlap = 1;
    hc = hc_;
    for (int i = 0; i < ndim; ++i) { K *= k; nlap *= 5; }
    lap = 2;
    conv = hc * 2 * K + hc;
    field = 3 * conv + hc + 1;
    base = lap + nlap;
  }
  __host__ __device__ int w(int q, int i) const { return base + q * field + i * conv; }   // Wh{i+1}_q.weight
  __host__ __device__ int b(int q, int i) const { return w(q, i) + hc * 2 * K; }          // Wh{i+1}_q.bias
  __host__ __device__ int w4(int q) const { return base + q * field + 3 * conv; }         // Wh4_q.weight, then bias
  __host__ __device__ int total() const { return base + 2 * field; }
};

// k = 5 weights as the conv kernel wants them (see kernels_pi_k5.cuh):
//   W5[q][cp][f][dy][dx][i][2]  (cp = channel pair; innermost 2 = channels 2cp, 2cp+1), then
//   bias[q][i][c], w4[q][c], b4[q]
constexpr int kK5RowFloats = 32;  // 5 dx * 3 convs * 2 channels = 30, padded to 32 (8 x LDS.128)
__host__ __device__ inline int k5_weight_floats(int hc) { return 2 * (hc / 2) * 2 * 5 * kK5RowFloats; }
__host__ __device__ inline int k5_total_floats(int hc) { return k5_weight_floats(hc) + 2 * 3 * hc + 2 * hc + 2; }

__device__ __forceinline__ int mono_index(int a, int b) {
  const int deg = a + b;
  return deg * (deg + 1) / 2 + b;
}

template <typename T>
__device__ void extract_cross_taps(const T* __restrict__ dense, int ndim, double scale, double* __restrict__ out) {
  // out[P_LAP_C0], out[P_LAP_AX + a*4 + k]
  const int stride3[3] = {25, 5, 1}, stride2[2] = {5, 1};
  const int centre = ndim == 3 ? 62 : 12;
  const int offs[4] = {-2, -1, 1, 2};
  out[P_LAP_C0] = double(dense[centre]) * scale;
  for (int a = 0; a < ndim; ++a)
    for (int k = 0; k < 4; ++k) {
      const int st = ndim == 3 ? stride3[a] : stride2[a];
      out[P_LAP_AX + a * 4 + k] = double(dense[centre + offs[k] * st]) * scale;
    }
}

template <typename T>
__global__ void k_prep(const T* __restrict__ raw, PrepDesc d, PrepBlock* __restrict__ out, float* __restrict__ k5w) {
  __shared__ double v[P_SIZE];
  for (int i = threadIdx.x; i < P_SIZE; i += blockDim.x) v[i] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    v[P_DT] = d.dt;
    if (d.cell == PERCNN_CELL_PI) {
      const PiPacking pk(d.ndim, d.k, d.hc);
      for (int q = 0; q < 2; ++q) {
        const double c = double(raw[q]);
        v[P_ALPHA + q] = d.coef_mode == PERCNN_COEF_RAW ? c : d.mu_up / (1.0 + exp(-c));
      }
      extract_cross_taps<T>(raw + pk.lap, d.ndim, 1.0, v);
      if (d.k == 1) {
        for (int q = 0; q < 2; ++q) {
          double* c = v + P_POLY + 10 * q;
          const int hc = d.hc;
          for (int ch = 0; ch < hc; ++ch) {
            double A[3][3];
            for (int i = 0; i < 3; ++i) {
              A[i][0] = double(raw[pk.w(q, i) + 2 * ch + 0]);
              A[i][1] = double(raw[pk.w(q, i) + 2 * ch + 1]);
              A[i][2] = double(raw[pk.w(q, i) + 2 * hc + ch]);
            }
            const double w4 = double(raw[pk.w4(q) + ch]);
            for (int i = 0; i < 3; ++i)
              for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) {
                  const int a = (i == 0) + (j == 0) + (k == 0), b = (i == 1) + (j == 1) + (k == 1);
                  c[mono_index(a, b)] += w4 * A[0][i] * A[1][j] * A[2][k];
                }
          }
          c[0] += double(raw[pk.w4(q) + hc]);
          double* du = v + P_DPOLY + 12 * q;
          double* dv = du + 6;
          du[0] = c[1]; du[1] = 2 * c[3]; du[2] = c[4]; du[3] = 3 * c[6]; du[4] = 2 * c[7]; du[5] = c[8];
          dv[0] = c[2]; dv[1] = c[4]; dv[2] = 2 * c[5]; dv[3] = c[7]; dv[4] = 2 * c[8]; dv[5] = 3 * c[9];
          // raw copy for PERCNN_FLAG_EVAL_BRANCH: W1[hc][2] b1[hc] W2 b2 W3 b3 W4[hc] b4
          double* B = v + P_BRANCH + q * (10 * hc + 1);
          for (int i = 0; i < 3; ++i)
            for (int e = 0; e < 3 * hc; ++e) B[i * 3 * hc + e] = double(raw[pk.w(q, i) + e]);
          for (int e = 0; e <= hc; ++e) B[9 * hc + e] = double(raw[pk.w4(q) + e]);
        }
      }
    } else if (d.cell == PERCNN_CELL_BURGERS) {
      // nu_u nu_v C1_u C2_u C1_v C2_v | laplace_op.filter.weight | dx_op.filter.weight | dy_op.filter.weight
      v[P_ALPHA + 0] = double(raw[0]);
      v[P_ALPHA + 1] = double(raw[1]);
      for (int i = 0; i < 4; ++i) v[P_PHYS + i] = double(raw[2 + i]);
      extract_cross_taps<T>(raw + 6, 2, 1.0 / (d.dx * d.dx), v);
      const int offs[4] = {-2, -1, 1, 2};
      for (int k = 0; k < 4; ++k) {
        v[P_PHYS + 4 + k] = double(raw[6 + 25 + 12 + offs[k] * 5]) / d.dx;   // dx_op: along rows (axis 0)
        v[P_PHYS + 8 + k] = double(raw[6 + 50 + 12 + offs[k]]) / d.dx;       // dy_op: along columns (axis 1)
      }
    } else {
      // nu_u nu_v C1..C5_u C1..C5_v [C6_v] | laplace_op.filter.weight
      const int nc = (d.flags & PERCNN_FLAG_LO_C6) ? 13 : 12;
      v[P_ALPHA + 0] = double(raw[0]);
      v[P_ALPHA + 1] = double(raw[1]);
      for (int i = 0; i < 10; ++i) v[P_PHYS + i] = double(raw[2 + i]);
      v[P_PHYS + 10] = nc == 13 ? double(raw[12]) : 0.0;
      extract_cross_taps<T>(raw + nc, 2, 1.0 / (d.dx * d.dx), v);
    }
  }
  __syncthreads();
  if (threadIdx.x < 13) {   // mirrored taps for the adjoint kernels
    const int i = threadIdx.x;
    v[P_LAPT + i] = i == 0 ? v[P_LAP_C0] : v[P_LAP_AX + ((i - 1) / 4) * 4 + (3 - (i - 1) % 4)];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P_SIZE; i += blockDim.x) {
    out->d[i] = v[i];
    out->f[i] = float(v[i]);
  }
  if (k5w != nullptr && d.cell == PERCNN_CELL_PI && d.k == 5) {
    const PiPacking pk(d.ndim, d.k, d.hc);
    const int hc = d.hc, ncp = hc / 2;
    const int nw = k5_weight_floats(hc);
    for (int e = threadIdx.x; e < nw; e += blockDim.x) {
      int r = e;
      const int col = r % kK5RowFloats; r /= kK5RowFloats;
      const int dy = r % 5; r /= 5;
      const int f = r % 2; r /= 2;
      const int cp = r % ncp; r /= ncp;
      const int q = r;
      float val = 0.f;
      if (col < 30) {
        const int dx = col / 6, i = (col % 6) / 2, c = 2 * cp + (col & 1);
        val = float(raw[pk.w(q, i) + ((c * 2 + f) * 5 + dy) * 5 + dx]);
      }
      k5w[e] = val;
    }
    float* bias = k5w + nw;
    for (int e = threadIdx.x; e < 2 * 3 * hc; e += blockDim.x) {
      const int c = e % hc, i = (e / hc) % 3, q = e / (3 * hc);
      bias[e] = float(raw[pk.w(q, i) + hc * 2 * 25 + c]);
    }
    float* w4 = bias + 2 * 3 * hc;
    for (int e = threadIdx.x; e < 2 * hc; e += blockDim.x) w4[e] = float(raw[pk.w4(e / hc) + (e % hc)]);
    if (threadIdx.x < 2) w4[2 * hc + threadIdx.x] = float(raw[pk.w4(threadIdx.x) + hc]);
  }
}

// acc[] (fp64 sums over cells and steps, dt already folded in) -> gradients in the raw packing.
template <typename T>
__global__ void k_finish_small(const T* __restrict__ raw, const double* __restrict__ acc, PrepDesc d, int nparams,
                               T* __restrict__ grads) {
  for (int i = threadIdx.x; i < nparams; i += blockDim.x) grads[i] = T(0);
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (d.cell == PERCNN_CELL_PI) {
    const PiPacking pk(d.ndim, d.k, d.hc);
    for (int q = 0; q < 2; ++q) {
      double g = acc[q];
      if (d.coef_mode == PERCNN_COEF_SIGMOID) {
        const double s = 1.0 / (1.0 + exp(-double(raw[q])));
        g *= d.mu_up * s * (1.0 - s);
      }
      grads[q] = T(g);
    }
    if (d.k != 1) return;
    const int hc = d.hc;
    for (int q = 0; q < 2; ++q) {
      const double* dc = acc + 2 + 10 * q;
      grads[pk.w4(q) + hc] = T(dc[0]);
      for (int ch = 0; ch < hc; ++ch) {
        double A[3][3], dA[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < 3; ++i) {
          A[i][0] = double(raw[pk.w(q, i) + 2 * ch + 0]);
          A[i][1] = double(raw[pk.w(q, i) + 2 * ch + 1]);
          A[i][2] = double(raw[pk.w(q, i) + 2 * hc + ch]);
        }
        const double w4 = double(raw[pk.w4(q) + ch]);
        double dw4 = 0;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) {
              const int a = (i == 0) + (j == 0) + (k == 0), b = (i == 1) + (j == 1) + (k == 1);
              const double tg = dc[mono_index(a, b)];
              dw4 += tg * A[0][i] * A[1][j] * A[2][k];
              dA[0][i] += tg * w4 * A[1][j] * A[2][k];
              dA[1][j] += tg * w4 * A[0][i] * A[2][k];
              dA[2][k] += tg * w4 * A[0][i] * A[1][j];
            }
        grads[pk.w4(q) + ch] = T(dw4);
        for (int i = 0; i < 3; ++i) {
          grads[pk.w(q, i) + 2 * ch + 0] = T(dA[i][0]);
          grads[pk.w(q, i) + 2 * ch + 1] = T(dA[i][1]);
          grads[pk.w(q, i) + 2 * hc + ch] = T(dA[i][2]);
        }
      }
    }
  } else if (d.cell == PERCNN_CELL_BURGERS) {
    for (int i = 0; i < 6; ++i) grads[i] = T(acc[i]);
  } else {
    const int nc = (d.flags & PERCNN_FLAG_LO_C6) ? 13 : 12;
    for (int i = 0; i < nc; ++i) grads[i] = T(acc[i]);
  }
}

}  // namespace percnn
