// Initial-state generator ("upscaler") of the training scripts and its adjoint.
//
// Reference (GS2D:26-41, GS3D:41-56, BUR1:38-52 = LO1:38-52 = BUR3:38-52):
//     two-layer:  ConvTranspose(2 -> C, k5, s2, p2, op1) -> Sigmoid -> ConvTranspose(C -> C, k5, s, p2, op s-1) -> Conv1x1(C -> 2)
//     one-layer:  ConvTranspose2d(2 -> 16, k5, s2, p2, op1) -> Tanh -> Conv1x1(16 -> 2)
// A transposed conv with kernel 5, padding 2 and stride S is, in gather form,
//     out[co, o] = b[co] + sum_ci sum_k in[ci, (o + 2 - k) / S] W[ci, co, k]      over the taps with S | (o + 2 - k) and
//                                                                                 0 <= (o + 2 - k) / S < extent_in
// per axis (zero padding: out-of-range inputs contribute nothing).  Nothing non-linear sits between the second
// transposed conv and the 1x1 conv, so they fold into ONE transposed conv C -> 2,
//     E[ci, f, k] = sum_co W2[ci, co, k] W3[f, co],   be[f] = b3[f] + sum_co W3[f, co] b2[co]
// (fp64, once per call), which cuts the second stage from C*C*K + 2C to 2*C*K multiply-adds per cell (4x for C = 8)
// and never materialises the C-channel full-resolution tensor.  The adjoint exploits the same fold: with
//     C2[f, ci, k] = sum_o g[f, o] mid[ci, i(o, k)]        (correlation of dL/dh0 with the stored activations)
// every gradient of the second stage is a tiny fp64 contraction:  dW2 = W3^T C2,  dW3 = b2 (x) sum g + W2 . C2,
// db2 = W3^T sum g,  db3 = sum g.  All sums are per-block fp64 partials folded in fixed order (deterministic).
//
// Layout: every tensor is channel-major [c][z][y][x]; along the slowest axis a buffer may hold only the planes
// [z0, z0 + nz) of the global grid (slab mode: each rank produces its own planes of h0 from the replicated
// low-resolution input; the transposed convs pad with ZEROS at the global border, they are not periodic).
#pragma once
#include "common.cuh"

namespace percnn {
namespace up {

constexpr int kThreads = 256;
constexpr int kMaxC = 16;

// A channel-major tensor holding planes [z0, z0 + nz) of a global [D][H][W] grid.
struct Grid {
  int D, H, W;       // global extents (2-D: D = 1)
  int z0, nz;        // local plane range
  int64_t cstride;   // elements between channels
};
__device__ __forceinline__ int64_t at(const Grid& g, int z, int y, int x) {
  return (int64_t(z - g.z0) * g.H + y) * g.W + x;
}

template <typename T>
__device__ __forceinline__ T act_fwd(int act, T x) {
  if (act == 0) return T(1) / (T(1) + exp(-x));   // GS2D:34 Sigmoid
  return tanh(x);                                  // BUR1:46 Tanh
}
template <typename T>
__device__ __forceinline__ T act_bwd(int act, T m) {   // derivative from the stored activation
  if (act == 0) return m * (T(1) - m);
  return T(1) - m * m;
}

// acc[j][co] += sum_ci sum_taps in[ci, (o + 2 - k) / S] Wsm[ci][k][co]   for the XT cells o = (z, y, x0 + j);
// x0 is a multiple of XT and XT a multiple of S, so the tap parity of cell j is known at compile time.
template <typename T, int NDIM, int S, int CIN, int COUT, int XT>
__device__ __forceinline__ void convt_gather(const T* __restrict__ in, const Grid& gi, const T* __restrict__ Wsm, int z, int y,
                                             int x0, T (&acc)[XT][COUT]) {
  static_assert(XT % S == 0, "tile width must be a multiple of the stride");
  constexpr int RMIN = -(2 / S);                  // S=1: -2, S=2: -1
  constexpr int RMAX = (XT - 1 + 2) / S;
  constexpr int NR = RMAX - RMIN + 1;
  constexpr int KZ = NDIM == 3 ? 5 : 1;
  const int xi0 = x0 / S;
  for (int ci = 0; ci < CIN; ++ci) {
    const T* inc = in + int64_t(ci) * gi.cstride;
    for (int kz = 0; kz < KZ; ++kz) {
      int iz = 0;
      if (NDIM == 3) {
        const int nz = z + 2 - kz;
        if (nz < 0 || (nz % S) != 0) continue;
        iz = nz / S;
        if (iz >= gi.D) continue;
      }
      for (int ky = 0; ky < 5; ++ky) {
        const int ny = y + 2 - ky;
        if (ny < 0 || (ny % S) != 0) continue;
        const int iy = ny / S;
        if (iy >= gi.H) continue;
        const T* row = inc + at(gi, NDIM == 3 ? iz : gi.z0, iy, 0);
        T v[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int ix = xi0 + RMIN + r;
          v[r] = (ix >= 0 && ix < gi.W) ? __ldg(row + ix) : T(0);
        }
        const T* w = Wsm + (int64_t(ci) * (KZ * 25) + (kz * 5 + ky) * 5) * COUT;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
          for (int j = 0; j < XT; ++j) {
            if ((j + 2 - kx) % S != 0) continue;
            const T val = v[(j + 2 - kx) / S - RMIN];
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[j][co] = fma_t(val, w[kx * COUT + co], acc[j][co]);
          }
        }
      }
    }
  }
}

// Digested weights in the workspace (element type T):
//   A1 [2][K][C] | b1 [C] | E [C][K][2] | be [2] | Et [2][K][C] | W3 [2][C] | b3 [2]          K = 5^ndim
struct PrepOff {
  int A1, b1, E, be, Et, W3, b3, total;
};
__host__ __device__ inline PrepOff prep_off(int C, int K) {
  PrepOff o;
  o.A1 = 0;
  o.b1 = o.A1 + 2 * K * C;
  o.E = o.b1 + C;
  o.be = o.E + C * K * 2;
  o.Et = o.be + 2;
  o.W3 = o.Et + 2 * K * C;
  o.b3 = o.W3 + 2 * C;
  o.total = o.b3 + 2;
  return o;
}
// Raw packing = the reference's state_dict order: W1 [2][C][K] b1 [C] (W2 [C][C][K] b2 [C]) W3 [2][C] b3 [2].
struct RawOff {
  int W1, b1, W2, b2, W3, b3, total;
};
__host__ __device__ inline RawOff raw_off(int C, int K, int layers) {
  RawOff o;
  o.W1 = 0;
  o.b1 = o.W1 + 2 * C * K;
  o.W2 = o.b1 + C;
  o.b2 = o.W2 + (layers == 2 ? C * C * K : 0);
  o.W3 = o.b2 + (layers == 2 ? C : 0);
  o.b3 = o.W3 + 2 * C;
  o.total = o.b3 + 2;
  return o;
}

template <typename T>
__global__ void k_up_prep(const T* __restrict__ raw, T* __restrict__ prep, int C, int K, int layers) {
  const RawOff r = raw_off(C, K, layers);
  const PrepOff p = prep_off(C, K);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < 2 * K * C; i += nth) {   // A1[ci][k][c] = W1[ci][c][k]
    const int c = i % C, k = (i / C) % K, ci = i / (C * K);
    prep[p.A1 + i] = raw[r.W1 + (ci * C + c) * K + k];
  }
  for (int i = tid; i < C; i += nth) prep[p.b1 + i] = raw[r.b1 + i];
  for (int i = tid; i < 2 * C; i += nth) prep[p.W3 + i] = raw[r.W3 + i];
  for (int i = tid; i < 2; i += nth) prep[p.b3 + i] = raw[r.b3 + i];
  if (layers == 2) {
    for (int i = tid; i < C * K * 2; i += nth) {   // E[ci][k][f]
      const int f = i % 2, k = (i / 2) % K, ci = i / (2 * K);
      double s = 0;
      for (int co = 0; co < C; ++co) s += double(raw[r.W2 + (ci * C + co) * K + k]) * double(raw[r.W3 + f * C + co]);
      prep[p.E + i] = T(s);
      prep[p.Et + (f * K + k) * C + ci] = T(s);
    }
    for (int f = tid; f < 2; f += nth) {
      double s = double(raw[r.b3 + f]);
      for (int co = 0; co < C; ++co) s += double(raw[r.W3 + f * C + co]) * double(raw[r.b2 + co]);
      prep[p.be + f] = T(s);
    }
  }
}

template <typename T>
__device__ __forceinline__ void load_smem(T* dst, const T* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// ---- forward, first stage: mid = act(ConvT_s2(low));  one-layer nets also emit h0 = W3 mid + b3 -----------------
template <typename T, int NDIM, int C, bool ONE_LAYER>
__global__ void __launch_bounds__(kThreads) k_up_l1(Grid glow, Grid gmid, Grid gout, int act, const T* __restrict__ prep,
                                                    const T* __restrict__ low, T* __restrict__ mid, T* __restrict__ out) {
  constexpr int K = NDIM == 3 ? 125 : 25;
  constexpr int XT = 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  PrepOff po = prep_off(C, K);
  {   // shared memory holds A1 | b1 | W3 | b3 only
    const int n1 = 2 * K * C + C, n2 = 2 * C + 2;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) sm[i] = prep[po.A1 + i];
    for (int i = threadIdx.x; i < n2; i += blockDim.x) sm[n1 + i] = prep[po.W3 + i];
    __syncthreads();
    po.W3 = n1;
    po.b3 = n1 + 2 * C;
  }
  const int xt = gmid.W / XT;                                   // W_mid = 2 W_low is even
  const int64_t ntile = int64_t(gmid.nz) * gmid.H * xt;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < ntile; t += int64_t(gridDim.x) * blockDim.x) {
    const int x0 = int(t % xt) * XT;
    const int64_t r = t / xt;
    const int y = int(r % gmid.H), z = gmid.z0 + int(r / gmid.H);
    T acc[XT][C];
#pragma unroll
    for (int j = 0; j < XT; ++j)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[j][c] = sm[po.b1 + c];
    convt_gather<T, NDIM, 2, 2, C, XT>(low, glow, sm + po.A1, z, y, x0, acc);
    const int64_t o = at(gmid, z, y, x0);
#pragma unroll
    for (int j = 0; j < XT; ++j) {
      T h0 = sm[po.b3 + 0], h1 = sm[po.b3 + 1];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const T m = act_fwd(act, acc[j][c]);
        mid[int64_t(c) * gmid.cstride + o + j] = m;
        if (ONE_LAYER) {
          h0 = fma_t(sm[po.W3 + c], m, h0);
          h1 = fma_t(sm[po.W3 + C + c], m, h1);
        }
      }
      if (ONE_LAYER) {   // the mid grid IS the output grid
        const int64_t oo = at(gout, z, y, x0 + j);
        out[oo] = h0;
        out[gout.cstride + oo] = h1;
      }
    }
  }
}

// ---- forward, second stage (two-layer nets): h0 = be + ConvT_S(mid) with the folded weights E ---------------------
template <typename T, int NDIM, int C, int S>
__global__ void __launch_bounds__(kThreads) k_up_l2(Grid gmid, Grid gout, const T* __restrict__ prep, const T* __restrict__ mid,
                                                    T* __restrict__ out) {
  constexpr int K = NDIM == 3 ? 125 : 25;
  constexpr int XT = 2 * S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const PrepOff po = prep_off(C, K);
  load_smem(sm, prep + po.E, C * K * 2 + 2);   // E | be are contiguous
  const T* E = sm;
  const T* be = sm + C * K * 2;
  const int xt = (gout.W + XT - 1) / XT;
  const int64_t ntile = int64_t(gout.nz) * gout.H * xt;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < ntile; t += int64_t(gridDim.x) * blockDim.x) {
    const int x0 = int(t % xt) * XT;
    const int64_t r = t / xt;
    const int y = int(r % gout.H), z = gout.z0 + int(r / gout.H);
    T acc[XT][2];
#pragma unroll
    for (int j = 0; j < XT; ++j) acc[j][0] = be[0], acc[j][1] = be[1];
    convt_gather<T, NDIM, S, C, 2, XT>(mid, gmid, E, z, y, x0, acc);
    const int64_t o = at(gout, z, y, x0);
#pragma unroll
    for (int j = 0; j < XT; ++j)
      if (x0 + j < gout.W) {
        out[o + j] = acc[j][0];
        out[gout.cstride + o + j] = acc[j][1];
      }
  }
}

// ---- adjoint: gmid = dL/d(pre-activation of stage 1) ------------------------------------------------------------
// one-layer: gmid[c] = (W3[0][c] g[0] + W3[1][c] g[1]) act'(mid[c]), pointwise.
template <typename T, int C>
__global__ void __launch_bounds__(kThreads) k_up_gmid1(Grid gmid, Grid gout, int act, const T* __restrict__ prep, int K,
                                                       const T* __restrict__ mid, const T* __restrict__ g, T* __restrict__ gm) {
  const PrepOff po = prep_off(C, K);
  const int64_t n = int64_t(gmid.nz) * gmid.H * gmid.W;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(i % gmid.W);
    const int64_t r = i / gmid.W;
    const int y = int(r % gmid.H), z = gmid.z0 + int(r / gmid.H);
    const int64_t oo = at(gout, z, y, x);
    const T g0 = __ldg(g + oo), g1 = __ldg(g + gout.cstride + oo);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const T m = __ldg(mid + int64_t(c) * gmid.cstride + i);
      gm[int64_t(c) * gmid.cstride + i] = fma_t(__ldg(prep + po.W3 + c), g0, __ldg(prep + po.W3 + C + c) * g1) * act_bwd(act, m);
    }
  }
}
// two-layer: d mid[ci, i] = sum_f sum_k g[f, S i - 2 + k] E[ci][k][f]  (a strided gather over dL/dh0), times act'.
// `gmid` describes gm (the owned mid planes); `mid` points at the first owned plane of the stored activations, whose
// channel stride `mid_cstride` may be larger (slab mode keeps extra planes for stage 2).
template <typename T, int NDIM, int C, int S>
__global__ void __launch_bounds__(kThreads) k_up_gmid2(Grid gmid, Grid gout, int act, const T* __restrict__ prep, const T* __restrict__ mid,
                                                       const T* __restrict__ g, T* __restrict__ gm, int64_t mid_cstride) {
  constexpr int K = NDIM == 3 ? 125 : 25;
  constexpr int KZ = NDIM == 3 ? 5 : 1;
  constexpr int XT = 2;
  constexpr int NV = S * (XT - 1) + 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const PrepOff po = prep_off(C, K);
  load_smem(sm, prep + po.Et, 2 * K * C);       // Et[f][k][ci]
  const int xt = gmid.W / XT;
  const int64_t ntile = int64_t(gmid.nz) * gmid.H * xt;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < ntile; t += int64_t(gridDim.x) * blockDim.x) {
    const int x0 = int(t % xt) * XT;
    const int64_t r = t / xt;
    const int y = int(r % gmid.H), z = gmid.z0 + int(r / gmid.H);
    T acc[XT][C];
#pragma unroll
    for (int j = 0; j < XT; ++j)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[j][c] = T(0);
    for (int f = 0; f < 2; ++f) {
      const T* gf = g + int64_t(f) * gout.cstride;
      for (int kz = 0; kz < KZ; ++kz) {
        const int oz = NDIM == 3 ? S * z - 2 + kz : gout.z0;
        if (NDIM == 3 && (oz < 0 || oz >= gout.D)) continue;
        for (int ky = 0; ky < 5; ++ky) {
          const int oy = S * y - 2 + ky;
          if (oy < 0 || oy >= gout.H) continue;
          const T* row = gf + at(gout, oz, oy, 0);
          T v[NV];
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int ox = S * x0 - 2 + q;
            v[q] = (ox >= 0 && ox < gout.W) ? __ldg(row + ox) : T(0);
          }
          const T* w = sm + (int64_t(f) * K + (kz * 5 + ky) * 5) * C;
#pragma unroll
          for (int kx = 0; kx < 5; ++kx)
#pragma unroll
            for (int j = 0; j < XT; ++j) {
              const T val = v[S * j + kx];
#pragma unroll
              for (int c = 0; c < C; ++c) acc[j][c] = fma_t(val, w[kx * C + c], acc[j][c]);
            }
        }
      }
    }
    const int64_t o = at(gmid, z, y, x0);
#pragma unroll
    for (int j = 0; j < XT; ++j)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        gm[int64_t(c) * gmid.cstride + o + j] = acc[j][c] * act_bwd(act, __ldg(mid + int64_t(c) * mid_cstride + o + j));
      }
  }
}

// ---- adjoint: correlations -----------------------------------------------------------------------------------------
// R[ca][cb][k] = sum_o A[ca, o] B[cb, (o + 2 - k) / S]  over the valid taps (the weight gradient of a transposed conv
// whose input is B and whose output gradient is A), followed by CA plain sums  R[CA CB K + ca] = sum_o A[ca, o].
// `single_tap`: K = 1, the centre tap only (B on the same grid as A): the 1x1 conv's weight gradient.
// A block owns a contiguous range of A rows (one fp64 partial vector per block, added up by k_up_fold in fixed order:
// deterministic).  Inside it a WARP takes one (chunk of CAT channels of A, cb, kz, ky) combination at a time and its
// lanes stride along x, so the loads of A and of the five kx-shifted B values are coalesced; every lane keeps the
// CAT x 5 sums of the five kx taps in registers and a shuffle tree folds the lanes once per combination.
// (A first version gave every THREAD its own combination: 32 different rows per load instruction, bound by L1
// sector traffic at 1 % of the FP32 pipe.)
constexpr int kCorrMaxVB = 1184;
template <typename T, int CAT>
__global__ void __launch_bounds__(kThreads) k_up_corr(Grid ga, Grid gb, int ndim, int S, int CA, int CB, int single_tap,
                                                      const T* __restrict__ A, const T* __restrict__ B,
                                                      double* __restrict__ partials) {
  const int KZ = single_tap ? 1 : (ndim == 3 ? 5 : 1);
  const int KY = single_tap ? 1 : 5;
  const int K = KZ * KY * KY;
  const int NS = CA * CB * K;
  const int ncombo = (CA / CAT) * CB * KZ * KY;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t nrows = int64_t(ga.nz) * ga.H;
  const int64_t r0 = nrows * blockIdx.x / gridDim.x, r1 = nrows * (blockIdx.x + 1) / gridDim.x;
  double* out = partials + size_t(blockIdx.x) * (NS + CA);
  for (int combo = warp; combo < ncombo; combo += nwarps) {
    const int kyi = combo % KY, kzi = (combo / KY) % KZ;      // tap indices (0 when the axis has a single tap)
    const int ky = single_tap ? 2 : kyi;
    const int kz = KZ == 1 ? 2 : kzi;                         // (no z axis / single tap: the centre, offset 0)
    const int cb = (combo / (KY * KZ)) % CB;
    const int chunk = combo / (KY * KZ * CB);
    const bool sums_too = cb == 0 && kyi == 0 && kzi == 0;    // one combination per chunk also adds up A itself
    const T* a0 = A + int64_t(chunk * CAT) * ga.cstride;
    const T* b0 = B + int64_t(cb) * gb.cstride;
    T acc[CAT][5], sacc[CAT];
#pragma unroll
    for (int c = 0; c < CAT; ++c) {
      sacc[c] = T(0);
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[c][k] = T(0);
    }
    int y = int(r0 % ga.H), z = ga.z0 + int(r0 / ga.H);     // walked along with r (no 64-bit division per row)
    for (int64_t r = r0; r < r1; ++r, ++y) {
      if (y == ga.H) {
        y = 0;
        ++z;
      }
      bool row_ok = true;
      int iz = gb.z0, iy = 0;
      if (ndim == 3) {
        const int nz = z + 2 - kz;
        row_ok = nz >= 0 && (S == 1 || (nz & 1) == 0);
        iz = S == 1 ? nz : nz >> 1;
        row_ok = row_ok && iz < gb.D;
      }
      {
        const int ny = y + 2 - ky;
        row_ok = row_ok && ny >= 0 && (S == 1 || (ny & 1) == 0);
        iy = S == 1 ? ny : ny >> 1;
        row_ok = row_ok && iy < gb.H;
      }
      if (!row_ok && !sums_too) continue;
      const T* arow = a0 + r * ga.W;
      const T* brow = b0 + (row_ok ? at(gb, iz, iy, 0) : 0);
      for (int x = lane; x < ga.W; x += 32) {
        T a[CAT];
#pragma unroll
        for (int c = 0; c < CAT; ++c) {
          a[c] = __ldg(arow + int64_t(c) * ga.cstride + x);
          sacc[c] += a[c];
        }
        if (!row_ok) continue;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          if (single_tap && kx != 2) continue;
          const int nx = x + 2 - kx;
          if (nx < 0 || (S == 2 && (nx & 1))) continue;
          const int ix = S == 1 ? nx : nx >> 1;
          if (ix >= gb.W) continue;
          const T bv = __ldg(brow + ix);
#pragma unroll
          for (int c = 0; c < CAT; ++c) acc[c][kx] = fma_t(a[c], bv, acc[c][kx]);
        }
      }
    }
    // fold the lanes (fixed shuffle tree), lane 0 stores
#pragma unroll
    for (int c = 0; c < CAT; ++c) {
      const int ca = chunk * CAT + c;
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        if (single_tap && kx != 2) continue;
        double v = double(acc[c][kx]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) {
          if (single_tap)
            out[ca * CB + cb] = v;
          else
            out[(ca * CB + cb) * K + (kzi * 5 + kyi) * 5 + kx] = v;
        }
      }
      if (sums_too) {
        double v = double(sacc[c]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) out[NS + ca] = v;
      }
    }
  }
}
// ---- the dominant correlation (3-D, stride 1, A = dL/dh0 with 2 channels, B = the 8 stored activations) on shared-
// memory tiles: R[f][ci][kz][ky][kx] = sum g[f, z, y, x] mid[ci, z + 2 - kz, y + 2 - ky, x + 2 - kx].
// A block owns an xy tile (C3_TY x C3_TX cells of A) and a chunk of z planes and MARCHES along z with the last five mid
// planes of the tile (+2 halo in x and y, zero-filled outside the grid = the transposed conv's zero padding) in a ring
// in shared memory, so every mid value is loaded once per chunk and then feeds 250 multiply-adds from shared memory.
// Thread = (ci, kz, ky) x half of the tile rows: it slides along x with a 5-value register window of its mid row
// and keeps the 5 kx x 2 f sums in registers: 3 LDS (one of them a broadcast) per 10 FFMA.  Row and slot strides are
// padded (69 / 6625 floats) so that the 32 (ky, kz, ci) rows a warp reads fall into distinct banks.
// The generic kernel above (one warp per combination, operands from L1) spent 0.58 ms on the 48^3 grid of the
// script itself and ~20 ms at 256^3, 3 % of the FP32 pipe.
constexpr int C3_TY = 8, C3_TX = 64, C3_CB = 8, C3_THREADS = 512;
constexpr int C3_ROWS = C3_TY + 4, C3_STRIDE = C3_TX + 5;                 // 12 rows of 69 floats per channel
constexpr int C3_SLOT = C3_CB * C3_ROWS * C3_STRIDE + 1;                  // 6625 floats per ring slot
constexpr int C3_SMEM_FLOATS = 5 * C3_SLOT + 2 * C3_TY * C3_TX;
__global__ void __launch_bounds__(C3_THREADS, 1) k_up_corr3(Grid ga, Grid gb, int ntx, int nty, int zc, const float* __restrict__ A,
                                                           const float* __restrict__ B, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  float* gt = ring + 5 * C3_SLOT;                                          // [2][TY][TX]
  __shared__ double s_half[200 * 10];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % (ntx * nty), chunk = blockIdx.x / (ntx * nty);
  const int x0 = (tile % ntx) * C3_TX, y0 = (tile / ntx) * C3_TY;
  const int zb = ga.z0 + chunk * zc, ze = min(zb + zc, ga.z0 + ga.nz);     // output planes [zb, ze)
  const bool worker = tid < 400;
  const int combo = tid % 200, half = tid / 200;                           // (half 2 = the 112 spare threads)
  const int ky = combo % 5, kz = (combo / 5) % 5, ci = combo / 25;
  float acc[2][5];
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[f][k] = 0.f;
  float gsum = 0.f;                                                        // threads 400, 401: sum of g[f] over the tile
  auto load_mid_plane = [&](int zp) {                                      // global plane zp -> ring slot zp mod 5
    float* slot = ring + ((zp % 5 + 5) % 5) * C3_SLOT;
    const bool zok = zp >= 0 && zp < gb.D;
    for (int e = tid; e < C3_CB * C3_ROWS * (C3_TX + 4); e += C3_THREADS) {
      const int mx = e % (C3_TX + 4);
      const int my = (e / (C3_TX + 4)) % C3_ROWS;
      const int c = e / ((C3_TX + 4) * C3_ROWS);
      const int gx = x0 - 2 + mx, gy = y0 - 2 + my;
      float v = 0.f;
      if (zok && gx >= 0 && gx < gb.W && gy >= 0 && gy < gb.H) v = __ldg(B + int64_t(c) * gb.cstride + at(gb, zp, gy, gx));
      slot[(c * C3_ROWS + my) * C3_STRIDE + mx] = v;
    }
  };
  for (int zp = zb - 2; zp < zb + 2; ++zp) load_mid_plane(zp);             // warm-up: planes z-2 .. z+1 of the first output plane
  for (int z = zb; z < ze; ++z) {
    load_mid_plane(z + 2);
    for (int e = tid; e < 2 * C3_TY * C3_TX; e += C3_THREADS) {
      const int rx = e % C3_TX, ry = (e / C3_TX) % C3_TY, f = e / (C3_TX * C3_TY);
      const int gx = x0 + rx, gy = y0 + ry;
      gt[e] = (gx < ga.W && gy < ga.H) ? __ldg(A + int64_t(f) * ga.cstride + at(ga, z, gy, gx)) : 0.f;
    }
    __syncthreads();
    if (worker) {
      const int zs = z + 2 - kz;                                           // mid plane of this thread's tap
      const float* mslot = ring + ((zs % 5 + 5) % 5) * C3_SLOT + ci * C3_ROWS * C3_STRIDE;
#pragma unroll 1
      for (int r = 0; r < C3_TY / 2; ++r) {
        const int ry = half * (C3_TY / 2) + r;
        const float* mrow = mslot + (ry + 4 - ky) * C3_STRIDE;             // mid row y + 2 - ky; element j <-> x0 - 2 + j
        const float* g0 = gt + ry * C3_TX;
        const float* g1 = g0 + C3_TY * C3_TX;
        // window w[j] = mid[x + j - 2 .. ], tap kx reads mid[x + 2 - kx] = mrow[rx + 4 - kx]
        float w0 = mrow[0], w1 = mrow[1], w2 = mrow[2], w3 = mrow[3];
#pragma unroll 4
        for (int rx = 0; rx < C3_TX; ++rx) {
          const float w4 = mrow[rx + 4];
          const float a0 = g0[rx], a1 = g1[rx];
          acc[0][0] = fmaf(a0, w4, acc[0][0]); acc[1][0] = fmaf(a1, w4, acc[1][0]);   // kx = 0 -> mrow[rx + 4]
          acc[0][1] = fmaf(a0, w3, acc[0][1]); acc[1][1] = fmaf(a1, w3, acc[1][1]);
          acc[0][2] = fmaf(a0, w2, acc[0][2]); acc[1][2] = fmaf(a1, w2, acc[1][2]);
          acc[0][3] = fmaf(a0, w1, acc[0][3]); acc[1][3] = fmaf(a1, w1, acc[1][3]);
          acc[0][4] = fmaf(a0, w0, acc[0][4]); acc[1][4] = fmaf(a1, w0, acc[1][4]);   // kx = 4 -> mrow[rx]
          w0 = w1; w1 = w2; w2 = w3; w3 = w4;
        }
      }
    } else if (tid < 402) {
      const float* gf = gt + (tid - 400) * C3_TY * C3_TX;
      float t = 0.f;
      for (int e = 0; e < C3_TY * C3_TX; ++e) t += gf[e];
      gsum += t;
    }
    __syncthreads();                                                       // the slot of plane z - 2 and the g tile are rewritten next
  }
  // fold the two row halves in fp64 and store this block's partial vector
  const int NS = 2 * C3_CB * 125;
  double* out = partials + size_t(blockIdx.x) * (NS + 2);
  if (worker && half == 1) {
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int k = 0; k < 5; ++k) s_half[combo * 10 + f * 5 + k] = double(acc[f][k]);
  }
  __syncthreads();
  if (worker && half == 0) {
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
        out[(f * C3_CB + ci) * 125 + (kz * 5 + ky) * 5 + kx] = double(acc[f][kx]) + s_half[combo * 10 + f * 5 + kx];
  } else if (tid >= 400 && tid < 402) {
    out[NS + (tid - 400)] = double(gsum);
  }
}

__global__ void __launch_bounds__(256) k_up_fold(const double* __restrict__ partials, int nblocks, int n, double* __restrict__ sums) {
  // 32 sums per block; eight thread groups stride over the partial vectors (coalesced along the sums, eight loads in
  // flight per sum instead of one serial chain of L2 round trips), then a fixed-order fold of the eight
  __shared__ double sm[8][32];
  const int il = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + il;
  double s0 = 0, s1 = 0;
  if (i < n) {
    int b = grp;
    for (; b + 8 < nblocks; b += 16) {
      s0 += partials[size_t(b) * n + i];
      s1 += partials[size_t(b + 8) * n + i];
    }
    if (b < nblocks) s0 += partials[size_t(b) * n + i];
  }
  sm[grp][il] = s0 + s1;
  __syncthreads();
  if (grp == 0 && i < n) {
    double t = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sm[q][il];
    sums[i] = t;
  }
}

// ---- adjoint: chain rule back to the raw packing (fp64) -----------------------------------------------------------
// sums1 = [C][2][K] correlation of gmid with low | [C] sums of gmid;
// sums2 = two-layer: [2][C][K] correlation of g with mid | [2] sums of g;  one-layer: [2][C] | [2].
// `accumulate`: add to g_params instead of overwriting (second gradient source, e.g. the IC loss).
template <typename T>
__global__ void k_up_finish(const T* __restrict__ raw, const double* __restrict__ sums1, const double* __restrict__ sums2, int C,
                            int K, int layers, T* __restrict__ gp, int accumulate) {
  const RawOff r = raw_off(C, K, layers);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  auto put = [&](int idx, double v) { gp[idx] = accumulate ? T(double(gp[idx]) + v) : T(v); };
  for (int i = tid; i < 2 * C * K; i += nth) {     // dW1[ci][c][k] = sums1[c][ci][k]
    const int k = i % K, c = (i / K) % C, ci = i / (K * C);
    put(r.W1 + i, sums1[(c * 2 + ci) * K + k]);
  }
  for (int c = tid; c < C; c += nth) put(r.b1 + c, sums1[C * 2 * K + c]);
  if (layers == 2) {
    const double* C2 = sums2;                        // [f][ci][k]
    const double* Sg = sums2 + 2 * C * K;
    for (int i = tid; i < C * C * K; i += nth) {     // dW2[ci][co][k] = sum_f W3[f][co] C2[f][ci][k]
      const int k = i % K, co = (i / K) % C, ci = i / (K * C);
      put(r.W2 + i, double(raw[r.W3 + co]) * C2[(0 * C + ci) * K + k] + double(raw[r.W3 + C + co]) * C2[(1 * C + ci) * K + k]);
    }
    for (int co = tid; co < C; co += nth) put(r.b2 + co, double(raw[r.W3 + co]) * Sg[0] + double(raw[r.W3 + C + co]) * Sg[1]);
    // (dW3 is a C*K-term contraction per entry: k_up_finish_w3, one block per entry)
    for (int f = tid; f < 2; f += nth) put(r.b3 + f, Sg[f]);
  } else {
    for (int i = tid; i < 2 * C; i += nth) put(r.W3 + i, sums2[i]);
    for (int f = tid; f < 2; f += nth) put(r.b3 + f, sums2[2 * C + f]);
  }
}

// dW3[f][co] = b2[co] sum g[f] + sum_{ci,k} W2[ci][co][k] C2[f][ci][k]: one block per entry, 128 threads stride over the
// C*K terms, fixed-order tree in shared memory (deterministic).
template <typename T>
__global__ void __launch_bounds__(128) k_up_finish_w3(const T* __restrict__ raw, const double* __restrict__ sums2, int C, int K,
                                                      T* __restrict__ gp, int accumulate) {
  __shared__ double sm[128];
  const RawOff r = raw_off(C, K, 2);
  const double* C2 = sums2;
  const double* Sg = sums2 + 2 * C * K;
  const int i = blockIdx.x, co = i % C, f = i / C;
  double s = 0;
  for (int t = threadIdx.x; t < C * K; t += 128) {
    const int ci = t / K, k = t - ci * K;
    s += double(raw[r.W2 + (ci * C + co) * K + k]) * C2[(f * C + ci) * K + k];
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double v = sm[0] + double(raw[r.b2 + co]) * Sg[f];
    gp[r.W3 + i] = accumulate ? T(double(gp[r.W3 + i]) + v) : T(v);
  }
}

// ---- IC loss (GS2D:331-338): mse(pred, target) and its gradient ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_mse_partial(const T* __restrict__ a, const T* __restrict__ b, int64_t n,
                                                          double* __restrict__ partials) {
  __shared__ double sm[kThreads / 32];
  double s = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double d = double(__ldg(a + i)) - double(__ldg(b + i));
    s += d * d;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sm[w];
    partials[blockIdx.x] = t;
  }
}
template <typename T>
__global__ void k_mse_finish(const double* __restrict__ partials, int nblocks, double inv_n, T* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += partials[b];
    out[0] = T(s * inv_n);
  }
}
// g[i] (+)= gscale * 2/n * (a[i] - b[i])
template <typename T>
__global__ void __launch_bounds__(kThreads) k_mse_grad(const T* __restrict__ a, const T* __restrict__ b, int64_t n,
                                                       double two_over_n, const T* __restrict__ gscale, T* __restrict__ g,
                                                       int accumulate) {
  const T coef = T(two_over_n * (gscale != nullptr ? double(__ldg(gscale)) : 1.0));
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const T v = coef * (__ldg(a + i) - __ldg(b + i));
    g[i] = accumulate ? g[i] + v : v;
  }
}

}  // namespace up
}  // namespace percnn
