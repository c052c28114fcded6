// Physics-residual loss of the training scripts, fused.
//
// Reference (FWD:288-357 `loss_generator.get_phy_Loss` + `loss_gen`; GS2D:270-353; GS3D:286-345):
//     output <- periodic pad (2 cells before, 3 after, every spatial axis)
//     lap    = valid 5-point-per-axis conv of output[0:-2] with the 4th-order table / dx^2     -> extent + 1 points
//     q_t    = Conv1d([-1, 1, 0]) / dt over time of output[:, q, 2:-2, ...]                      -> (q[t+1] - q[t]) / dt
//     f_q    = D_q * lap_q + R_q(u, v) - q_t          R_q = the PDE's reaction term, a cubic in (u, v)
//     loss   = mse(f_u, 0) + mse(f_v, 0)
// which the reference executes as 2 full-trajectory pad copies, 2 convs, 2 permute+reshape copies, 2 conv1d and
// ~15 pointwise passes.  Because of the (2, 3) padding the residual is evaluated on extent + 1 points per axis,
// the last one being the periodic image of the first: on the un-padded grid every point simply carries the weight
//     w(x) = prod_axes (1 + [x_axis == 0])
// and N = (nframes - 2) * prod (extent + 1).  Here:
//   k_phys_resid : one pass over frames 0 .. nframes-3: f_u, f_v per cell from the cross neighbourhood (periodic
//                  addressing, no padding copies), sum of w f^2 (fp64 partials, deterministic), and optionally the
//                  residual gradient  R_q = 2 w f_q / N  for the backward pass.
//   k_phys_grad  : dL/d frame_t = D_q Lap^T(R_q[t]) + sum_p dR_p/dq R_p[t] + (R_q[t] - R_q[t-1]) / dt   (dense, it
//                  touches every frame), scaled by the upstream gradient read from device memory.
#pragma once
#include "kernels_generic.cuh"

namespace percnn {

template <typename T>
struct PhysLossDev {
  T diff[2];        // D_u, D_v
  T poly[2][10];    // reaction cubics, c00 c10 c01 c20 c11 c02 c30 c21 c12 c03
  T dpoly[4][6];    // (dRu/du, dRu/dv, dRv/du, dRv/dv) over monomials 1 u v u^2 uv v^2
  T inv_dt;
  T tap[5];         // 1-D 4th-order second-derivative taps at offsets -2 -1 0 +1 +2, already / dx^2
  double two_over_n;
  int nframes;
  int64_t stride;   // elements between consecutive frames
};

// 4th-order cross Laplacian with the loss module's own (isotropic) table; symmetric, so it is its own transpose.
template <typename T, int NDIM>
__device__ __forceinline__ T phys_lap(const Cross<T, NDIM>& q, const T* __restrict__ tap) {
  T acc = T(NDIM) * tap[2] * q.c;
#pragma unroll
  for (int a = 0; a < NDIM; ++a) {
    acc = fma_t(tap[0], q.n[a][0] + q.n[a][3], acc);
    acc = fma_t(tap[1], q.n[a][1] + q.n[a][2], acc);
  }
  return acc;
}

template <int NDIM>
__device__ __forceinline__ int phys_weight(const Geom& g, int64_t cell) {
  const int x = int(cell % g.W);
  const int64_t r = cell / g.W;
  int w = (x == 0) ? 2 : 1;
  if (NDIM == 2) {
    w *= (r == 0) ? 2 : 1;
  } else {
    w *= (r % g.H == 0) ? 2 : 1;
    w *= (r / g.H == 0) ? 2 : 1;
  }
  return w;
}

template <typename T, int NDIM>
__global__ void __launch_bounds__(kGenericThreads) k_phys_resid(Geom g, PhysLossDev<T> pl, const T* __restrict__ frames,
                                                                T* __restrict__ resid, double* __restrict__ partials,
                                                                unsigned* __restrict__ counter, double* __restrict__ acc) {
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  const int64_t total = ncell * (pl.nframes - 2);
  double red[1] = {0.0};
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(i / ncell);
    const int64_t cell = i - int64_t(t) * ncell;
    const T* cur = frames + int64_t(t) * pl.stride;
    const T* nxt = cur + pl.stride;
    const CellOffsets<NDIM> o = cell_offsets<NDIM>(g, cell);
    const Cross<T, NDIM> U = gather<T, NDIM>(cur, o);
    const Cross<T, NDIM> V = gather<T, NDIM>(cur + g.field, o);
    const T fu = fma_t(pl.diff[0], phys_lap<T, NDIM>(U, pl.tap), cubic_eval(pl.poly[0], U.c, V.c)) -
                 (__ldg(nxt + o.c) - U.c) * pl.inv_dt;
    const T fv = fma_t(pl.diff[1], phys_lap<T, NDIM>(V, pl.tap), cubic_eval(pl.poly[1], U.c, V.c)) -
                 (__ldg(nxt + g.field + o.c) - V.c) * pl.inv_dt;
    const int w = phys_weight<NDIM>(g, cell);
    red[0] += double(w) * (double(fu) * double(fu) + double(fv) * double(fv));
    if (resid != nullptr) {
      const T s = T(pl.two_over_n * double(w));
      T* r = resid + int64_t(t) * pl.stride;
      r[o.c] = s * fu;
      r[g.field + o.c] = s * fv;
    }
  }
  reduce_into_acc<double, 1>(red, partials, counter, acc);
}

template <typename T, int NDIM>
__global__ void __launch_bounds__(kGenericThreads) k_phys_grad(Geom g, PhysLossDev<T> pl, const T* __restrict__ frames,
                                                               const T* __restrict__ resid, const T* __restrict__ gscale,
                                                               T* __restrict__ gout) {
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  const int64_t total = ncell * pl.nframes;
  const T scale = gscale != nullptr ? __ldg(gscale) : T(1);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(i / ncell);
    const int64_t cell = i - int64_t(t) * ncell;
    const CellOffsets<NDIM> o = cell_offsets<NDIM>(g, cell);
    T gu = T(0), gv = T(0);
    if (t <= pl.nframes - 3) {
      const T* r = resid + int64_t(t) * pl.stride;
      const T* h = frames + int64_t(t) * pl.stride;
      const Cross<T, NDIM> RU = gather<T, NDIM>(r, o);
      const Cross<T, NDIM> RV = gather<T, NDIM>(r + g.field, o);
      const T u = __ldg(h + o.c), v = __ldg(h + g.field + o.c);
      gu = fma_t(pl.diff[0], phys_lap<T, NDIM>(RU, pl.tap), RU.c * pl.inv_dt);
      gv = fma_t(pl.diff[1], phys_lap<T, NDIM>(RV, pl.tap), RV.c * pl.inv_dt);
      gu = fma_t(RU.c, quad_eval(pl.dpoly[0], u, v), fma_t(RV.c, quad_eval(pl.dpoly[2], u, v), gu));
      gv = fma_t(RU.c, quad_eval(pl.dpoly[1], u, v), fma_t(RV.c, quad_eval(pl.dpoly[3], u, v), gv));
    }
    if (t >= 1 && t <= pl.nframes - 2) {
      const T* rp = resid + int64_t(t - 1) * pl.stride;
      gu -= __ldg(rp + o.c) * pl.inv_dt;
      gv -= __ldg(rp + g.field + o.c) * pl.inv_dt;
    }
    T* d = gout + int64_t(t) * pl.stride;
    d[o.c] = scale * gu;
    d[g.field + o.c] = scale * gv;
  }
}

}  // namespace percnn
