// Generic cross-stencil step kernels: any extent, 2-D or 3-D, fp32 or fp64, periodic or slab-ghosted
// along the slowest axis.  One thread walks a grid-stride list of cells; neighbours come through
// L1/L2 with wrap-around index arithmetic.  These cover the reference's own sizes (100^2, 48^3), the
// fp64 variants and every adjoint; the TMA z-marching kernel (kernels_gs3d_tma.cuh) takes over for
// large aligned 3-D fp32 grids.
#pragma once
#include "point_ops.cuh"

namespace percnn {


constexpr int kGenericThreads = 256;

// Element offsets (within one field) of a cell and of its cross neighbours.
template <int NDIM>
struct CellOffsets {
  int64_t c;
  int64_t n[NDIM][4];
};

template <int NDIM>
__device__ __forceinline__ CellOffsets<NDIM> cell_offsets(const Geom& g, int64_t cell) {
  CellOffsets<NDIM> o;
  const int offs[4] = {-2, -1, 1, 2};
  const int x = int(cell % g.W);
  const int64_t r = cell / g.W;
  if (NDIM == 2) {
    const int y = int(r);  // slowest axis: ghost-aware
    const int64_t row = int64_t(y + g.ghost) * g.W;
    o.c = row + x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = g.ghost ? (y + g.ghost + offs[k]) : wrap_idx(y + offs[k], g.H);
      o.n[0][k] = int64_t(yy) * g.W + x;
      o.n[NDIM - 1][k] = row + wrap_idx(x + offs[k], g.W);
    }
  } else {
    const int y = int(r % g.H);
    const int z = int(r / g.H);
    const int64_t pl = int64_t(z + g.ghost) * g.plane;
    const int64_t row = pl + int64_t(y) * g.W;
    o.c = row + x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int zz = g.ghost ? (z + g.ghost + offs[k]) : wrap_idx(z + offs[k], g.D);
      o.n[0][k] = int64_t(zz) * g.plane + int64_t(y) * g.W + x;
      o.n[1][k] = pl + int64_t(wrap_idx(y + offs[k], g.H)) * g.W + x;
      o.n[NDIM - 1][k] = row + wrap_idx(x + offs[k], g.W);
    }
  }
  return o;
}

template <typename T, int NDIM>
__device__ __forceinline__ Cross<T, NDIM> gather(const T* __restrict__ f, const CellOffsets<NDIM>& o) {
  Cross<T, NDIM> q;
  q.c = __ldg(f + o.c);
#pragma unroll
  for (int a = 0; a < NDIM; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) q.n[a][k] = __ldg(f + o.n[a][k]);
  return q;
}

// Deterministic grid reduction: per-block partials in fp64, the last block to finish adds the column
// sums (fixed block order) into acc[].  No float atomics anywhere, so gradients are reproducible.
template <typename T, int NRED>
__device__ __forceinline__ void reduce_into_acc(const T (&v)[NRED], double* __restrict__ partials,
                                                unsigned* __restrict__ counter, double* __restrict__ acc) {
  __shared__ double sm[kGenericThreads / 32][NRED];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NRED; ++i) {
    double d = double(v[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_down_sync(0xffffffffu, d, off);
    if (lane == 0) sm[warp][i] = d;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kGenericThreads / 32; ++w) s += sm[w][threadIdx.x];
    partials[size_t(blockIdx.x) * NRED + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    fold_partials(partials, gridDim.x, NRED, acc, 1.0);
    if (threadIdx.x == 0) *counter = 0;
  }
}

// ---- Pi-block, k = 1 ------------------------------------------------------------------------
template <typename T, int NDIM, bool BRANCH>
__global__ void __launch_bounds__(kGenericThreads) k_pi_k1_fwd(Geom g, int slot, int hc, const T* __restrict__ src,
                                                               T* __restrict__ dst) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<NDIM> o = cell_offsets<NDIM>(g, cell);
    const Cross<T, NDIM> U = gather<T, NDIM>(src, o);
    const Cross<T, NDIM> V = gather<T, NDIM>(src + g.field, o);
    const T Lu = lap_apply<T, NDIM>(U, P), Lv = lap_apply<T, NDIM>(V, P);
    T ou, ov;
    if (BRANCH)
      pi_k1_fwd_branch<T>(U.c, V.c, Lu, Lv, P, hc, ou, ov);
    else
      pi_k1_fwd_poly<T>(U.c, V.c, Lu, Lv, P, ou, ov);
    dst[o.c] = ou;
    dst[g.field + o.c] = ov;
  }
}

template <typename T, int NDIM>
__global__ void __launch_bounds__(kGenericThreads) k_pi_k1_bwd(Geom g, int slot, const T* __restrict__ h,
                                                               const T* __restrict__ gout, const T* __restrict__ gadd,
                                                               T* __restrict__ gin, double* __restrict__ partials,
                                                               unsigned* __restrict__ counter, double* __restrict__ acc,
                                                               Inject<T> inj) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  const T icoef = inject_coef(inj);
  T red[kRedPiK1];
#pragma unroll
  for (int i = 0; i < kRedPiK1; ++i) red[i] = T(0);
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<NDIM> o = cell_offsets<NDIM>(g, cell);
    const Cross<T, NDIM> GU = gather<T, NDIM>(gout, o);
    const Cross<T, NDIM> GV = gather<T, NDIM>(gout + g.field, o);
    const T u = __ldg(h + o.c), v = __ldg(h + g.field + o.c);
    T gu, gv;
    pi_k1_bwd_poly<T>(u, v, GU.c, GV.c, lap_apply_T<T, NDIM>(GU, P), lap_apply_T<T, NDIM>(GV, P), P, gu, gv, red);
    if (gadd != nullptr) {
      gu += __ldg(gadd + o.c);
      gv += __ldg(gadd + g.field + o.c);
    }
    if (inj.target != nullptr) {
      const int64_t r = cell / g.W;
      inject_cell<T>(inj, icoef, NDIM == 3 ? int(r / g.H) : 0, NDIM == 3 ? int(r % g.H) : int(r), int(cell % g.W), u, v, gu, gv);
    }
    gin[o.c] = gu;
    gin[g.field + o.c] = gv;
  }
  reduce_into_acc<T, kRedPiK1>(red, partials, counter, acc);
}

// ---- fused data loss, forward value (percnn_data_loss_fwd) -------------------------------------------
// Sum over the sampling lattice of ONE state of (h - target)^2, both fields, added to acc[0] (fp64, fixed order).
// Reads only the sampled points: 1/s^ndim of the state.
template <typename T>
__global__ void __launch_bounds__(kGenericThreads) k_data_loss(Geom g, const T* __restrict__ h, Inject<T> inj,
                                                               double* __restrict__ partials, unsigned* __restrict__ counter,
                                                               double* __restrict__ acc) {
  double red[1] = {0.0};
  const int64_t npts = 2 * inj.lfield;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npts; i += int64_t(gridDim.x) * blockDim.x) {
    const int f = int(i / inj.lfield);
    const int64_t r = i - f * inj.lfield;
    const int lx = int(r % inj.lw);
    const int64_t t = r / inj.lw;
    const int ly = int(t % inj.lh), lz = int(t / inj.lh);
    int64_t off;
    if (g.ndim == 3)
      off = int64_t(lz * inj.s + g.ghost) * g.plane + int64_t(ly * inj.s) * g.W + lx * inj.s;
    else
      off = int64_t(ly * inj.s + g.ghost) * g.W + lx * inj.s;
    const T d = __ldg(h + f * g.field + off) - __ldg(inj.target + i);
    red[0] += double(d) * double(d);
  }
  reduce_into_acc<double, 1>(red, partials, counter, acc);
}
// The loss gradient of the LAST state of a rollout (no adjoint step differentiates it): g += coef (h - target) on
// the sampling lattice.  One thread per low-res point.
template <typename T>
__global__ void __launch_bounds__(kGenericThreads) k_inject_only(Geom g, const T* __restrict__ h, T* __restrict__ gbuf,
                                                                 Inject<T> inj) {
  const T coef = inject_coef(inj);
  const int64_t npts = 2 * inj.lfield;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npts; i += int64_t(gridDim.x) * blockDim.x) {
    const int f = int(i / inj.lfield);
    const int64_t r = i - f * inj.lfield;
    const int lx = int(r % inj.lw);
    const int64_t t = r / inj.lw;
    const int ly = int(t % inj.lh), lz = int(t / inj.lh);
    int64_t off;
    if (g.ndim == 3)
      off = int64_t(lz * inj.s + g.ghost) * g.plane + int64_t(ly * inj.s) * g.W + lx * inj.s;
    else
      off = int64_t(ly * inj.s + g.ghost) * g.W + lx * inj.s;
    off += f * g.field;
    gbuf[off] = fma_t(coef, __ldg(h + off) - __ldg(inj.target + i), gbuf[off]);
  }
}
template <typename T>
__global__ void k_data_loss_finish(const double* __restrict__ acc, double inv_n, T* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = T(acc[0] * inv_n);
}

// ---- Stage-3 physics cells (2-D) --------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenericThreads) k_burgers_fwd(Geom g, int slot, const T* __restrict__ src,
                                                                 T* __restrict__ dst) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.H) * g.W;
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<2> o = cell_offsets<2>(g, cell);
    const Cross<T, 2> U = gather<T, 2>(src, o);
    const Cross<T, 2> V = gather<T, 2>(src + g.field, o);
    T ou, ov;
    burgers_fwd<T>(U, V, P, ou, ov);
    dst[o.c] = ou;
    dst[g.field + o.c] = ov;
  }
}

template <typename T>
__global__ void __launch_bounds__(kGenericThreads) k_burgers_bwd(Geom g, int slot, const T* __restrict__ h,
                                                                 const T* __restrict__ gout, const T* __restrict__ gadd,
                                                                 T* __restrict__ gin, double* __restrict__ partials,
                                                                 unsigned* __restrict__ counter, double* __restrict__ acc,
                                                                 Inject<T> inj) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.H) * g.W;
  const T icoef = inject_coef(inj);
  T red[kRedBurgers];
#pragma unroll
  for (int i = 0; i < kRedBurgers; ++i) red[i] = T(0);
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<2> o = cell_offsets<2>(g, cell);
    const Cross<T, 2> U = gather<T, 2>(h, o);
    const Cross<T, 2> V = gather<T, 2>(h + g.field, o);
    const Cross<T, 2> GU = gather<T, 2>(gout, o);
    const Cross<T, 2> GV = gather<T, 2>(gout + g.field, o);
    T gu, gv;
    burgers_bwd<T>(U, V, GU, GV, P, gu, gv, red);
    if (gadd != nullptr) {
      gu += __ldg(gadd + o.c);
      gv += __ldg(gadd + g.field + o.c);
    }
    if (inj.target != nullptr) inject_cell<T>(inj, icoef, 0, int(cell / g.W), int(cell % g.W), U.c, V.c, gu, gv);
    gin[o.c] = gu;
    gin[g.field + o.c] = gv;
  }
  reduce_into_acc<T, kRedBurgers>(red, partials, counter, acc);
}

template <typename T>
__global__ void __launch_bounds__(kGenericThreads) k_lo_fwd(Geom g, int slot, const T* __restrict__ src,
                                                            T* __restrict__ dst) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.H) * g.W;
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<2> o = cell_offsets<2>(g, cell);
    const Cross<T, 2> U = gather<T, 2>(src, o);
    const Cross<T, 2> V = gather<T, 2>(src + g.field, o);
    T ou, ov;
    lo_fwd<T>(U.c, V.c, lap_apply<T, 2>(U, P), lap_apply<T, 2>(V, P), P, ou, ov);
    dst[o.c] = ou;
    dst[g.field + o.c] = ov;
  }
}

template <typename T>
__global__ void __launch_bounds__(kGenericThreads) k_lo_bwd(Geom g, int slot, const T* __restrict__ h,
                                                            const T* __restrict__ gout, const T* __restrict__ gadd,
                                                            T* __restrict__ gin, double* __restrict__ partials,
                                                            unsigned* __restrict__ counter, double* __restrict__ acc,
                                                            Inject<T> inj) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.H) * g.W;
  const T icoef = inject_coef(inj);
  T red[kRedLO];
#pragma unroll
  for (int i = 0; i < kRedLO; ++i) red[i] = T(0);
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<2> o = cell_offsets<2>(g, cell);
    const Cross<T, 2> GU = gather<T, 2>(gout, o);
    const Cross<T, 2> GV = gather<T, 2>(gout + g.field, o);
    const T u = __ldg(h + o.c), v = __ldg(h + g.field + o.c);
    T gu, gv;
    lo_bwd<T>(u, v, GU.c, GV.c, lap_apply_T<T, 2>(GU, P), lap_apply_T<T, 2>(GV, P), P, gu, gv, red);
    if (gadd != nullptr) {
      gu += __ldg(gadd + o.c);
      gv += __ldg(gadd + g.field + o.c);
    }
    if (inj.target != nullptr) inject_cell<T>(inj, icoef, 0, int(cell / g.W), int(cell % g.W), u, v, gu, gv);
    gin[o.c] = gu;
    gin[g.field + o.c] = gv;
  }
  reduce_into_acc<T, kRedLO>(red, partials, counter, acc);
}

// ---- Stage-3 physics cells, classical RK4 step (forward_rk4, BUR3:159-206 / LO3: same) ------------------------
// One launch per stage.  Stage i evaluates k_i = f_rhs(h + a * k_{i-1}) (the intermediate state is formed on the fly
// for the whole cross neighbourhood, never stored), keeps k_i for the next stage and accumulates k1 + 2 k2 + 2 k3 in
// `acc`; the last stage writes acc <- h + dt * (acc + k4) / 6 in the reference's operation order.  `acc` is the
// caller's output buffer, so an RK4 step needs two state-sized scratch buffers (k ping-pong) instead of the ~12
// full-size temporaries of the stock-op version.
//   stage 0: kprev == nullptr, acc = k1;  stages 1, 2: acc += 2 k;  stage 3 (last): output.   CELL: 1 = Burgers, 2 = lambda-omega
template <typename T, int CELL>
__global__ void __launch_bounds__(kGenericThreads) k_rk4_stage(Geom g, int slot, const T* __restrict__ h, const T* __restrict__ kprev,
                                                               T a, T* __restrict__ kout, T* __restrict__ acc, int stage) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.H) * g.W;
  for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
       cell += int64_t(gridDim.x) * blockDim.x) {
    const CellOffsets<2> o = cell_offsets<2>(g, cell);
    Cross<T, 2> U = gather<T, 2>(h, o);
    Cross<T, 2> V = gather<T, 2>(h + g.field, o);
    const T u0 = U.c, v0 = V.c;
    if (kprev != nullptr) {   // y = h + k * dt / 2  (BUR3:186-187; a = dt / 2 or dt)
      const Cross<T, 2> KU = gather<T, 2>(kprev, o);
      const Cross<T, 2> KV = gather<T, 2>(kprev + g.field, o);
      U.c = u0 + KU.c * a;
      V.c = v0 + KV.c * a;
#pragma unroll
      for (int ax = 0; ax < 2; ++ax)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          U.n[ax][k] = U.n[ax][k] + KU.n[ax][k] * a;
          V.n[ax][k] = V.n[ax][k] + KV.n[ax][k] * a;
        }
    }
    T fu, fv;
    if (CELL == 1) burgers_rhs<T>(U, V, P, fu, fv);
    else lo_rhs<T>(U.c, V.c, lap_apply<T, 2>(U, P), lap_apply<T, 2>(V, P), P, fu, fv);
    if (stage == 0) {
      kout[o.c] = fu;
      kout[g.field + o.c] = fv;
      acc[o.c] = fu;
      acc[g.field + o.c] = fv;
    } else if (stage < 3) {
      kout[o.c] = fu;
      kout[g.field + o.c] = fv;
      acc[o.c] = acc[o.c] + T(2) * fu;
      acc[g.field + o.c] = acc[g.field + o.c] + T(2) * fv;
    } else {   // u0 + dt * (k1 + 2 k2 + 2 k3 + k4) / 6  (BUR3:201-202)
      const T dt = P[P_DT];
      acc[o.c] = u0 + dt * (acc[o.c] + fu) / T(6);
      acc[g.field + o.c] = v0 + dt * (acc[g.field + o.c] + fv) / T(6);
    }
  }
}

}  // namespace percnn

// =====================================================================================================
// Persistent multi-step kernel for small grids.  The 2-D configs (<= 512^2) and the reference's own sizes
// (100^2, 48^3) hold their whole state in L2 and a per-step kernel takes ~2 us, so a rollout of per-step
// launches is bound by launch latency (~4.3 us/step measured).  Here ONE cooperative launch runs all
// steps; blocks meet at a grid barrier (monotonic counter, release/acquire) between steps.  State loads
// bypass L1 (ld.global.cg): the buffers are rewritten by other SMs during the kernel.
// =====================================================================================================
namespace percnn {

template <typename T>
struct MultiStepArgs {
  const T* h0;      // initial state
  T* tape;          // if non-null: step s reads tape + s*stride, writes tape + (s+1)*stride (tape[0] must hold h0)
  T* ping;          // otherwise: ping-pong scratch ...
  T* pong;
  T* final_state;   // ... with the last step written here
  int nsteps;
  int64_t stride;   // elements between tape slots
};

template <typename T, int NDIM>
__device__ __forceinline__ Cross<T, NDIM> gather_cg(const T* __restrict__ f, const CellOffsets<NDIM>& o) {
  Cross<T, NDIM> q;
  q.c = __ldcg(f + o.c);
#pragma unroll
  for (int a = 0; a < NDIM; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) q.n[a][k] = __ldcg(f + o.n[a][k]);
  return q;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// CELL: 0 = Pi-block k=1 (folded cubic), 1 = Burgers physics, 2 = lambda-omega physics
// 512-thread blocks, at most one per SM: the barrier costs one serialized L2 atomic per block, so fewer, fatter
// blocks are faster than the 256-thread grids of the per-step kernels (measured 2.7 -> see profiles).
constexpr int kMultiThreads = 512;
template <typename T, int NDIM, int CELL>
__global__ void __launch_bounds__(kMultiThreads) k_multi_step(Geom g, int slot, MultiStepArgs<T> m, unsigned* counter) {
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  for (int s = 0; s < m.nsteps; ++s) {
    const T* src;
    T* dst;
    if (m.tape != nullptr) {
      src = m.tape + int64_t(s) * m.stride;
      dst = m.tape + int64_t(s + 1) * m.stride;
    } else {
      src = s == 0 ? m.h0 : ((s & 1) ? m.ping : m.pong);
      dst = s == m.nsteps - 1 ? m.final_state : ((s & 1) ? m.pong : m.ping);
    }
    for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
         cell += int64_t(gridDim.x) * blockDim.x) {
      const CellOffsets<NDIM> o = cell_offsets<NDIM>(g, cell);
      const Cross<T, NDIM> U = gather_cg<T, NDIM>(src, o);
      const Cross<T, NDIM> V = gather_cg<T, NDIM>(src + g.field, o);
      T ou, ov;
      if (CELL == 0) {
        pi_k1_fwd_poly<T>(U.c, V.c, lap_apply<T, NDIM>(U, P), lap_apply<T, NDIM>(V, P), P, ou, ov);
      } else if (CELL == 1) {
        if constexpr (NDIM == 2) burgers_fwd<T>(U, V, P, ou, ov);
      } else {
        lo_fwd<T>(U.c, V.c, lap_apply<T, NDIM>(U, P), lap_apply<T, NDIM>(V, P), P, ou, ov);
      }
      dst[o.c] = ou;
      dst[g.field + o.c] = ov;
    }
    if (s + 1 < m.nsteps) grid_barrier(counter, unsigned(s + 1) * gridDim.x);
  }
}

}  // namespace percnn

// =====================================================================================================
// Persistent multi-step kernel for SMALL SLABS (multi-GPU, cfg4-class grids: 128^3 over 2..8 GPUs).  A slab of a
// few MB holds a few microseconds of work per step, less than one kernel boundary plus the NVLink flag latency of
// the per-step slab kernel (kernels_gs3d_slab.cuh) -- round 1 measured 128^3 getting SLOWER with more GPUs.  Here
// ONE cooperative launch per rank runs the whole rollout: each step computes the slab (gather kernel, state in
// L2), stores the boundary planes also into the neighbours' ghost planes (peer mapping, st.global), and the
// per-step grid barrier doubles as the exchange point: block 0 waits for all blocks, raises both neighbours'
// flags (st.release.sys), waits for its own two flags (ld.acquire.sys) and only then opens the barrier.  Flags and
// epochs follow the protocol of the per-step kernels, so the two are interchangeable between rollout calls.
// =====================================================================================================
namespace percnn {

struct SlabMultiArgs {
  float* buf[2];              // my ping-pong state buffers [2][D+4][H][W]
  float* peer_lo[2];          // lower / upper neighbour's mappings of theirs
  float* peer_hi[2];
  const uint32_t* my_flags;   // [0] lower ghosts valid up to epoch, [1] upper ghosts
  uint32_t* post_lo_flag;     // lower neighbour's flags[1]
  uint32_t* post_hi_flag;     // upper neighbour's flags[0]
  uint32_t* err;              // error word (spin deadline)
  uint32_t epoch0;
  uint32_t spin_limit;
  int nsteps;
  int cur;                    // step s reads buf[cur ^ (s & 1)]
};

static __global__ void __launch_bounds__(kMultiThreads) k_multi_step_slab(Geom g, int slot, SlabMultiArgs a, unsigned* counter) {
  const float* P = PrepView<float>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  unsigned* go = counter + 1;   // second word of the barrier block: last step every block may leave
  for (int s = 0; s < a.nsteps; ++s) {
    const int si = a.cur ^ (s & 1), di = si ^ 1;
    const float* src = a.buf[si];
    float* dst = a.buf[di];
    float* plo = a.peer_lo[di];
    float* phi = a.peer_hi[di];
    bool wrote_peer = false;
    for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
         cell += int64_t(gridDim.x) * blockDim.x) {
      const CellOffsets<3> o = cell_offsets<3>(g, cell);
      const Cross<float, 3> U = gather_cg<float, 3>(src, o);
      const Cross<float, 3> V = gather_cg<float, 3>(src + g.field, o);
      float ou, ov;
      pi_k1_fwd_poly<float>(U.c, V.c, lap_apply<float, 3>(U, P), lap_apply<float, 3>(V, P), P, ou, ov);
      dst[o.c] = ou;
      dst[g.field + o.c] = ov;
      const int z = int(cell / g.plane);
      if (z < 2) {                      // -> the lower neighbour's upper ghost planes D+2, D+3
        const int64_t m = o.c + int64_t(g.D) * g.plane;
        plo[m] = ou;
        plo[g.field + m] = ov;
        wrote_peer = true;
      }
      if (z >= g.D - 2) {               // -> the upper neighbour's lower ghost planes 0, 1
        const int64_t m = o.c - int64_t(g.D) * g.plane;
        phi[m] = ou;
        phi[g.field + m] = ov;
        wrote_peer = true;
      }
    }
    if (wrote_peer) __threadfence_system();
    // ---- grid barrier + halo hand-shake ----
    const unsigned target = unsigned(s + 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      if (blockIdx.x == 0) {
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target * gridDim.x);
        // one system-scope fence orders every block's (already fenced) peer stores before both flags; the flags
        // themselves and the polling loads are relaxed, an acquire fence follows the poll: two NVLink round trips
        // less than release-store + acquire-load per neighbour
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        const uint32_t e = a.epoch0 + uint32_t(s) + 1u;
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.post_lo_flag), "r"(e) : "memory");
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.post_hi_flag), "r"(e) : "memory");
        for (uint32_t spins = 0;; ++spins) {
          uint32_t w0, w1;
          asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(w0) : "l"(a.my_flags + 0) : "memory");
          asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(w1) : "l"(a.my_flags + 1) : "memory");
          if (int32_t(w0 - e) >= 0 && int32_t(w1 - e) >= 0) break;
          if (spins >= a.spin_limit) {   // a lost neighbour must not hang the GPU, nor may the rollout go on
            atomicExch(a.err, 1u);
            __threadfence_system();
            __trap();
          }
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(target) : "memory");
      } else {
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(go) : "memory");
        } while (v < target);
      }
    }
    __syncthreads();
  }
}

}  // namespace percnn

// =====================================================================================================
// The same, COMMUNICATION-AVOIDING: neighbours exchange 2K ghost planes every K steps instead of 2 planes every step.
// The per-step hand-shake of k_multi_step_slab (system-scope fences behind the peer stores, flag round trip over
// NVLink, grid barrier) costs ~10 us whatever the slab holds -- more than the 8.8 us ONE GPU needs for a whole 128^3
// step, which is why cfg4 got slower on more GPUs.  Here a block of K steps runs between two hand-shakes:
//   * the state lives in "wide" buffers [2][D + 4K][H][W] (2K ghost planes per side);
//   * sub-step j of a block computes the planes [-e, D + e), e = 2 (K - 1 - j), from the previous sub-step's output:
//     the planes that can still be valid shrink by 2 per sub-step and end at the interior (redundant work: at D = 16,
//     K = 4 it is +37 % planes for a quarter of the hand-shakes); only a plain grid barrier separates sub-steps;
//   * the LAST sub-step stores the interior and mirrors its first / last 2K planes into the neighbours' ghost planes
//     of the buffer the NEXT block starts from.  Blocks alternate between two buffer PAIRS, so a fast neighbour never
//     writes ghost planes that a slow rank still reads (it can be at most one block ahead: it needs this rank's
//     planes of the previous block, published after this rank's last read of that block).
// The first step reads the caller's standard buffer (2 ghost planes: a block of one step) and the last block writes
// the standard buffer and the neighbours' standard ghost planes, with the same flags and epochs as the per-step
// kernels, so the three slab paths stay interchangeable between rollout calls.  Per-cell arithmetic is the gather
// kernel's: results are bit-identical to the single-GPU rollout.
// =====================================================================================================
namespace percnn {

struct SlabBlockedArgs {
  float* buf[2];              // standard ping-pong buffers [2][D+4][H][W] and the neighbours' mappings of theirs
  float* peer_lo[2];
  float* peer_hi[2];
  float* w[4];                // wide buffers [2][D+4K][H][W]: pairs (0,1) and (2,3)
  float* wlo[4];
  float* whi[4];
  const uint32_t* my_flags;
  uint32_t* post_lo_flag;
  uint32_t* post_hi_flag;
  uint32_t* err;
  uint32_t epoch0;
  uint32_t spin_limit;
  int nsteps;
  int cur;
  int K;
};

static __global__ void __launch_bounds__(kMultiThreads) k_multi_step_slab_tb(Geom g, int slot, SlabBlockedArgs a, unsigned* counter) {
  const float* P = PrepView<float>::get(c_prep[slot]);
  unsigned* go = counter + 1;
  const int G = 2 * a.K;
  const int64_t wfield = int64_t(g.D + 2 * G) * g.plane;
  unsigned nbar = 0;            // barriers passed so far (both kinds count on the same words)
  int done = 0, blk = 0;
  while (done < a.nsteps) {
    const bool first = done == 0;
    const int Kb = first ? 1 : min(a.K, a.nsteps - done);
    const bool last = done + Kb == a.nsteps;
    const int p = blk & 1;      // wide pair this block works in (blocks >= 1)
    for (int j = 0; j < Kb; ++j) {
      const int e = 2 * (Kb - 1 - j);
      const bool fin = j == Kb - 1;
      // ---- source: planes [-e-2, D+e+2) must be valid ----
      Geom gs = g;
      const float* src;
      if (first) {
        src = a.buf[a.cur];
      } else {
        src = a.w[2 * p + (j & 1)];
        gs.D = g.D + 2 * e;
        gs.ghost = G - e;
        gs.field = wfield;
      }
      // ---- destination ----
      float* dst;
      float* plo = nullptr;
      float* phi = nullptr;
      int64_t dfield;
      int dghost, send = 0;
      if (!fin) {
        dst = a.w[2 * p + ((j & 1) ^ 1)];
        dfield = wfield;
        dghost = G;
      } else if (last) {
        const int di = a.cur ^ (a.nsteps & 1);
        dst = a.buf[di];
        plo = a.peer_lo[di];
        phi = a.peer_hi[di];
        dfield = g.field;
        dghost = g.ghost;
        send = 2;
      } else {
        const int nw = 2 * ((blk + 1) & 1);     // the next block's first buffer (the other pair)
        dst = a.w[nw];
        plo = a.wlo[nw];
        phi = a.whi[nw];
        dfield = wfield;
        dghost = G;
        send = G;
      }
      // One thread = four x-adjacent cells: 11 x 128-bit L2 loads per field (centre, the two neighbouring quads, four
      // y rows, four z planes) instead of 52 scalar ones, and one round of loads per sub-step for most threads -- the
      // sub-step is bound by L2 latency, not by arithmetic.  The per-cell arithmetic is the gather kernel's own
      // (Cross -> lap_apply -> pi_k1_fwd_poly), so the results stay bit-identical to it.
      const int W4 = g.W >> 2;
      const int64_t nquad = int64_t(gs.D) * g.H * W4;
      bool wrote_peer = false;
      for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < nquad; q += int64_t(gridDim.x) * blockDim.x) {
        const int xq = int(q % W4) << 2;
        const int64_t r = q / W4;
        const int y = int(r % g.H), zl = int(r / g.H);
        const int64_t pl = int64_t(zl + gs.ghost) * g.plane;
        const int64_t row = pl + int64_t(y) * g.W;
        const int xl = xq == 0 ? g.W - 4 : xq - 4, xr = xq + 4 == g.W ? 0 : xq + 4;
        int64_t yo[4];
        {
          const int offs[4] = {-2, -1, 1, 2};
#pragma unroll
          for (int k = 0; k < 4; ++k) yo[k] = pl + int64_t(wrap_idx(y + offs[k], g.H)) * g.W + xq;
        }
        float lapv[2][4], ctr[2][4];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float* sf = src + f * gs.field;
          auto ld4 = [&](int64_t off) { return __ldcg(reinterpret_cast<const float4*>(sf + off)); };
          const float4 c4 = ld4(row + xq), l4 = ld4(row + xl), r4 = ld4(row + xr);
          const float4 y0 = ld4(yo[0]), y1 = ld4(yo[1]), y2 = ld4(yo[2]), y3 = ld4(yo[3]);
          const float4 z0 = ld4(row + xq - 2 * g.plane), z1 = ld4(row + xq - g.plane), z2 = ld4(row + xq + g.plane),
                       z3 = ld4(row + xq + 2 * g.plane);
          const float line[12] = {l4.x, l4.y, l4.z, l4.w, c4.x, c4.y, c4.z, c4.w, r4.x, r4.y, r4.z, r4.w};
          const float ya[4][4] = {{y0.x, y0.y, y0.z, y0.w}, {y1.x, y1.y, y1.z, y1.w}, {y2.x, y2.y, y2.z, y2.w}, {y3.x, y3.y, y3.z, y3.w}};
          const float za[4][4] = {{z0.x, z0.y, z0.z, z0.w}, {z1.x, z1.y, z1.z, z1.w}, {z2.x, z2.y, z2.z, z2.w}, {z3.x, z3.y, z3.z, z3.w}};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            Cross<float, 3> Q;
            Q.c = line[4 + j];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              Q.n[0][k] = za[k][j];
              Q.n[1][k] = ya[k][j];
            }
            Q.n[2][0] = line[2 + j];
            Q.n[2][1] = line[3 + j];
            Q.n[2][2] = line[5 + j];
            Q.n[2][3] = line[6 + j];
            lapv[f][j] = lap_apply<float, 3>(Q, P);
            ctr[f][j] = Q.c;
          }
        }
        float ou[4], ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pi_k1_fwd_poly<float>(ctr[0][j], ctr[1][j], lapv[0][j], lapv[1][j], P, ou[j], ov[j]);
        const float4 o_u = make_float4(ou[0], ou[1], ou[2], ou[3]), o_v = make_float4(ov[0], ov[1], ov[2], ov[3]);
        const int z = zl - e;                               // interior index
        const int64_t od = int64_t(z + dghost) * g.plane + int64_t(y) * g.W + xq;
        *reinterpret_cast<float4*>(dst + od) = o_u;
        *reinterpret_cast<float4*>(dst + dfield + od) = o_v;
        if (send) {
          if (z < send) {                    // -> the lower neighbour's upper ghost planes D .. D + send - 1
            const int64_t m = od + int64_t(g.D) * g.plane;
            *reinterpret_cast<float4*>(plo + m) = o_u;
            *reinterpret_cast<float4*>(plo + dfield + m) = o_v;
            wrote_peer = true;
          }
          if (z >= g.D - send) {             // -> the upper neighbour's lower ghost planes -send .. -1
            const int64_t m = od - int64_t(g.D) * g.plane;
            *reinterpret_cast<float4*>(phi + m) = o_u;
            *reinterpret_cast<float4*>(phi + dfield + m) = o_v;
            wrote_peer = true;
          }
        }
      }
      ++nbar;
      if (!fin) {
        grid_barrier(counter, nbar * gridDim.x);
        continue;
      }
      // ---- grid barrier + halo hand-shake (as in k_multi_step_slab) ----
      if (wrote_peer) __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        if (blockIdx.x == 0) {
          unsigned v;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          } while (v < nbar * gridDim.x);
          asm volatile("fence.acq_rel.sys;" ::: "memory");
          const uint32_t ep = a.epoch0 + uint32_t(done + Kb);
          asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.post_lo_flag), "r"(ep) : "memory");
          asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.post_hi_flag), "r"(ep) : "memory");
          for (uint32_t spins = 0;; ++spins) {
            uint32_t w0, w1;
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(w0) : "l"(a.my_flags + 0) : "memory");
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(w1) : "l"(a.my_flags + 1) : "memory");
            if (int32_t(w0 - ep) >= 0 && int32_t(w1 - ep) >= 0) break;
            if (spins >= a.spin_limit) {
              atomicExch(a.err, 1u);
              __threadfence_system();
              __trap();
            }
          }
          asm volatile("fence.acq_rel.sys;" ::: "memory");
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(nbar) : "memory");
        } else {
          unsigned v;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(go) : "memory");
          } while (v < nbar);
        }
      }
      __syncthreads();
    }
    done += Kb;
    ++blk;
  }
}

}  // namespace percnn

// =====================================================================================================
// Persistent multi-step ADJOINT for small grids: the backward twin of k_multi_step.  A training step on the 2-D
// configs (and the reference's own 100^2 / 48^3 grids) is bound by one launch + one 22-value grid reduction per
// time step (measured: 11 us per step at 128^2 fp64, of which the stencil is ~2 us).  Here ONE cooperative launch
// walks all steps backwards; the gradient ping-pongs between two L2-resident buffers (ld.global.cg: they are
// rewritten by other SMs during the kernel), the stored states / injected gradients / loss targets are read-only,
// and the parameter-gradient sums stay in per-thread fp64 accumulators across ALL steps -- the block reduction
// and the last-block fold run once per rollout instead of once per step.
// =====================================================================================================
namespace percnn {

constexpr int kMaxMultiBwdSteps = 4096;   // selection masks travel as kernel parameters (2 x 516 B)

template <typename T>
struct MultiBwdArgs {
  const T* tape;        // h_0 .. h_nsteps
  const T* g_tape;      // dense gradients of the masked states, packed in increasing step order (nullable)
  const T* g_init;      // G_nsteps (already holds dL/dh_nsteps or zeros)
  T* ping;              // scratch gradient buffers; g_init may alias ping
  T* pong;
  T* g_h0;              // receives dL/dh_0
  int nsteps;
  int g_slots;          // number of masked states among 0 .. nsteps-1
  int64_t stride;       // elements between tape slots
  Inject<T> inj;        // fused data loss: target = FIRST packed frame; frames are 2 * lfield apart
  int inj_slots;        // number of selected states among 0 .. nsteps-1
  uint32_t gmask[kMaxMultiBwdSteps / 32 + 1];     // bit s: dense gradient present for state s
  uint32_t selmask[kMaxMultiBwdSteps / 32 + 1];   // bit s: state s enters the fused data loss
};

template <int CELL>
struct MultiBwdRed {
  static constexpr int value = CELL == 0 ? kRedPiK1 : (CELL == 1 ? kRedBurgers : kRedLO);
};

template <typename T, int NDIM, int CELL>
__global__ void __launch_bounds__(kMultiThreads) k_multi_step_bwd(Geom g, int slot, const __grid_constant__ MultiBwdArgs<T> m,
                                                                  unsigned* counter, double* __restrict__ partials,
                                                                  unsigned* __restrict__ red_counter, double* __restrict__ acc) {
  constexpr int NRED = MultiBwdRed<CELL>::value;
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  double racc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; ++i) racc[i] = 0.0;
  const T icoef = inject_coef(m.inj);
  int gslot = m.g_slots, islot = m.inj_slots;
  const T* src = m.g_init;
  int flip = (m.g_init == m.ping) ? 1 : 0;
  for (int s = m.nsteps - 1; s >= 0; --s) {
    const T* gadd = nullptr;
    if (m.g_tape != nullptr && ((m.gmask[s >> 5] >> (s & 31)) & 1u)) gadd = m.g_tape + int64_t(--gslot) * m.stride;
    Inject<T> inj = m.inj;
    inj.target = nullptr;
    if (m.inj.target != nullptr && ((m.selmask[s >> 5] >> (s & 31)) & 1u)) inj.target = m.inj.target + int64_t(--islot) * 2 * m.inj.lfield;
    T* dst = (s == 0) ? m.g_h0 : (flip ? m.pong : m.ping);
    const T* h = m.tape + int64_t(s) * m.stride;
    // fp32: per-step partial sums in fp32, folded into the fp64 accumulators after every step (as the per-step
    // kernels do); fp64: accumulate in place (a second set of 22 doubles spilled)
    constexpr bool kDirect = sizeof(T) == 8;
    T red_step[kDirect ? 1 : NRED];
    T* red;
    if constexpr (kDirect) {
      red = reinterpret_cast<T*>(racc);
    } else {
#pragma unroll
      for (int i = 0; i < NRED; ++i) red_step[i] = T(0);
      red = red_step;
    }
    for (int64_t cell = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; cell < ncell;
         cell += int64_t(gridDim.x) * blockDim.x) {
      const CellOffsets<NDIM> o = cell_offsets<NDIM>(g, cell);
      const Cross<T, NDIM> GU = gather_cg<T, NDIM>(src, o);
      const Cross<T, NDIM> GV = gather_cg<T, NDIM>(src + g.field, o);
      T gu, gv, u, v;
      if constexpr (CELL == 1) {
        const Cross<T, 2> U = gather<T, 2>(h, o);
        const Cross<T, 2> V = gather<T, 2>(h + g.field, o);
        u = U.c;
        v = V.c;
        burgers_bwd<T>(U, V, GU, GV, P, gu, gv, red);
      } else {
        u = __ldg(h + o.c);
        v = __ldg(h + g.field + o.c);
        if constexpr (CELL == 0)
          pi_k1_bwd_poly<T>(u, v, GU.c, GV.c, lap_apply_T<T, NDIM>(GU, P), lap_apply_T<T, NDIM>(GV, P), P, gu, gv, red);
        else
          lo_bwd<T>(u, v, GU.c, GV.c, lap_apply_T<T, NDIM>(GU, P), lap_apply_T<T, NDIM>(GV, P), P, gu, gv, red);
      }
      if (gadd != nullptr) {
        gu += __ldg(gadd + o.c);
        gv += __ldg(gadd + g.field + o.c);
      }
      if (inj.target != nullptr) {
        const int64_t r = cell / g.W;
        inject_cell<T>(inj, icoef, NDIM == 3 ? int(r / g.H) : 0, NDIM == 3 ? int(r % g.H) : int(r), int(cell % g.W), u, v, gu, gv);
      }
      dst[o.c] = gu;
      dst[g.field + o.c] = gv;
    }
    if constexpr (!kDirect) {
#pragma unroll
      for (int i = 0; i < NRED; ++i) racc[i] += double(red_step[i]);
    }
    src = dst;
    flip ^= 1;
    if (s > 0) grid_barrier(counter, unsigned(m.nsteps - s) * gridDim.x);
  }
  // one block reduction + last-block fold for the whole rollout (kMultiThreads threads per block)
  __shared__ double sm[kMultiThreads / 32][NRED];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NRED; ++i) {
    double d = racc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_down_sync(0xffffffffu, d, off);
    if (lane == 0) sm[warp][i] = d;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kMultiThreads / 32; ++w) t += sm[w][threadIdx.x];
    partials[size_t(blockIdx.x) * NRED + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(red_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    fold_partials(partials, gridDim.x, NRED, acc, 1.0);
    if (threadIdx.x == 0) *red_counter = 0;
  }
}

}  // namespace percnn
