// Translation unit: Stage-2 library of candidate terms ("Stage-2 library construction").
//
// Reference: DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-2/derivatives.py:129-199 (`Loss_generator.get_phy_residual`,
// lambda-omega twin: `get_library`) fed by `get_residual_mse`'s periodic (2, 3) padding (derivatives.py:207-208), and
// PDE_FIND_u.py:185-193,246-259 (`gen_library`, the eval'd column products).  The reference runs 6 valid convs, two
// permute+reshape+Conv1d round trips and ~10 pointwise passes over the padded trajectory and then evaluates 70 column
// expressions on the host; here one kernel writes all twelve non-trivial terms from the cross neighbourhood (periodic
// addressing instead of padding copies) and one kernel gathers the sampled rows of the 70-column matrix in fp64.
#include "kernels_generic.cuh"
#include "plan.h"

using namespace percnn;

namespace {

constexpr int kTerms = 12;   // f_u f_v u v u_t v_t u_x u_y v_x v_y lap_u lap_v   ('ones' is implicit)

template <typename T>
struct LibDev {
  T dtap[4];     // first-derivative taps at offsets -2 -1 +1 +2, / dx   (derivatives.py:10-14, 115-119)
  T ltap[5];     // second-derivative taps at offsets -2..+2, / dx^2      (derivatives.py:24-28, 101-105)
  T inv_dt;
  int kind;      // 0 Burgers residual (derivatives.py:189-192), 1 lambda-omega (stage-2/derivatives.py:188-192)
  int H, W, nframes;
};

// One thread per output point (t, i, j), i in [0, H], j in [0, W]: point (i, j) of the padded-valid grid is cell
// (i mod H, j mod W) of the periodic grid.
template <typename T>
__global__ void __launch_bounds__(256) k_lib_terms(LibDev<T> L, const T* __restrict__ frames, T* __restrict__ terms) {
  const int H1 = L.H + 1, W1 = L.W + 1;
  const int64_t per = int64_t(H1) * W1, total = per * (L.nframes - 2);
  const int64_t field = int64_t(L.H) * L.W;
  for (int64_t n = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; n < total; n += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(n / per);
    const int r = int(n - int64_t(t) * per);
    const int i = (r / W1) % L.H, j = (r % W1) % L.W;
    const T* u = frames + int64_t(t) * 2 * field;
    const T* v = u + field;
    auto at = [&](const T* q, int di, int dj) {
      return __ldg(q + int64_t(wrap_idx(i + di, L.H)) * L.W + wrap_idx(j + dj, L.W));
    };
    const T uc = at(u, 0, 0), vc = at(v, 0, 0);
    // the conv filters are cross-correlations: tap k multiplies the value at offset k - 2 (derivatives.py:10-28)
    T ux = T(0), uy = T(0), vx = T(0), vy = T(0);
    T lu = T(2) * L.ltap[2] * uc, lv = T(2) * L.ltap[2] * vc;
    const int off[4] = {-2, -1, 1, 2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const T ui = at(u, off[k], 0), uj = at(u, 0, off[k]), vi = at(v, off[k], 0), vj = at(v, 0, off[k]);
      ux = fma_t(L.dtap[k], ui, ux);      // dx operator: taps along tensor dim 2 (derivatives.py:10-14)
      uy = fma_t(L.dtap[k], uj, uy);      // dy operator: taps along tensor dim 3 (derivatives.py:16-20)
      vx = fma_t(L.dtap[k], vi, vx);
      vy = fma_t(L.dtap[k], vj, vy);
      const T lt = L.ltap[k < 2 ? k : k + 1];
      lu = fma_t(lt, ui + uj, lu);
      lv = fma_t(lt, vi + vj, lv);
    }
    const T ut = (__ldg(u + 2 * field + int64_t(i) * L.W + j) - uc) * L.inv_dt;
    const T vt = (__ldg(v + 2 * field + int64_t(i) * L.W + j) - vc) * L.inv_dt;
    T fu, fv;
    if (L.kind == 0) {
      const T nu = T(1.0 / 200.0);
      fu = ut - nu * lu + uc * ux + vc * uy;
      fv = vt - nu * lv + uc * vx + vc * vy;
    } else {
      const T a = uc * uc + vc * vc;
      fu = ut - (T(0.1) * lu + (T(1) - a) * uc + a * vc);
      fv = vt - (T(0.1) * lv + (T(1) - a) * vc - a * uc);
    }
    const T vals[kTerms] = {fu, fv, uc, vc, ut, vt, ux, uy, vx, vy, lu, lv};
#pragma unroll
    for (int q = 0; q < kTerms; ++q) terms[int64_t(q) * total + n] = vals[q];
  }
}

// theta[r][a * 7 + b] = A_a(u, v) * B_b at point idx[r], in fp64 from the stored terms (PDE_FIND_u.py:185-193:
// listA = ones u v u**2 u*v v**2 u**3 u**2*v u*v**2 v**3, listB = ones u_x u_y v_x v_y lap_u lap_v);
// rhs[r] = (u_t, v_t).  One thread per row.
template <typename T>
__global__ void __launch_bounds__(256) k_lib_theta(const T* __restrict__ terms, int64_t total, const int64_t* __restrict__ idx,
                                                   int64_t n, double* __restrict__ theta, double* __restrict__ rhs) {
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x) {
    const int64_t p = idx[r];
    auto term = [&](int q) { return double(__ldg(terms + int64_t(q) * total + p)); };
    const double u = term(2), v = term(3);
    const double A[10] = {1.0, u, v, u * u, u * v, v * v, u * u * u, u * u * v, u * (v * v), v * v * v};
    const double B[7] = {1.0, term(6), term(7), term(8), term(9), term(10), term(11)};
    double* row = theta + r * 70;
#pragma unroll
    for (int a = 0; a < 10; ++a)
#pragma unroll
      for (int b = 0; b < 7; ++b) row[a * 7 + b] = A[a] * B[b];
    rhs[2 * r] = term(4);
    rhs[2 * r + 1] = term(5);
  }
}

int lib_check(const percnn_library_t* d) {
  if (!d) return fail(PERCNN_ERR_INVALID, "null library descriptor");
  if (d->dtype != PERCNN_F32 && d->dtype != PERCNN_F64) return fail(PERCNN_ERR_INVALID, "dtype must be f32 or f64");
  if (d->kind != 0 && d->kind != 1) return fail(PERCNN_ERR_INVALID, "kind must be 0 (Burgers) or 1 (lambda-omega)");
  if (d->H < 2 || d->W < 2 || d->H > (1 << 20) || d->W > (1 << 20)) return fail(PERCNN_ERR_INVALID, "bad extents (need 2 <= H, W)");
  if (d->nframes < 3) return fail(PERCNN_ERR_INVALID, "the library needs at least 3 frames (derivatives.py:146-147: output[0:-2])");
  if (!(d->dt > 0) || !(d->dx > 0)) return fail(PERCNN_ERR_INVALID, "dt and dx must be positive");
  if (!percnn_device_ok(d->device)) return fail(PERCNN_ERR_NO_DEVICE, "no sm_100 CUDA device with this ordinal");
  return PERCNN_OK;
}
template <typename T>
LibDev<T> lib_dev(const percnn_library_t* d) {
  LibDev<T> L;
  const double d1[4] = {1.0 / 12.0, -8.0 / 12.0, 8.0 / 12.0, -1.0 / 12.0};
  const double l1[5] = {-1.0 / 12.0, 4.0 / 3.0, -5.0 / 2.0, 4.0 / 3.0, -1.0 / 12.0};
  // the reference holds the tables in fp32 and divides the conv output by the resolution (derivatives.py:44-49,70-75)
  for (int k = 0; k < 4; ++k) L.dtap[k] = T(double(T(d1[k])) / d->dx);
  for (int k = 0; k < 5; ++k) L.ltap[k] = T(double(T(l1[k])) / (d->dx * d->dx));
  L.inv_dt = T(1.0 / d->dt);
  L.kind = d->kind;
  L.H = int(d->H);
  L.W = int(d->W);
  L.nframes = d->nframes;
  return L;
}
int lib_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return int(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int64_t percnn_library_points(const percnn_library_t* d) {
  if (lib_check(d)) return -1;
  return int64_t(d->nframes - 2) * (d->H + 1) * (d->W + 1);
}

int percnn_library_terms(const percnn_library_t* d, const void* frames, void* terms, void* stream) {
  if (int rc = lib_check(d)) return rc;
  if (!frames || !terms) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(d->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = percnn_library_points(d);
  if (d->dtype == PERCNN_F32)
    k_lib_terms<float><<<lib_grid(total), 256, 0, st>>>(lib_dev<float>(d), static_cast<const float*>(frames), static_cast<float*>(terms));
  else
    k_lib_terms<double><<<lib_grid(total), 256, 0, st>>>(lib_dev<double>(d), static_cast<const double*>(frames), static_cast<double*>(terms));
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

int percnn_library_theta(const percnn_library_t* d, const void* terms, const int64_t* idx, int64_t n, double* theta, double* rhs,
                         void* stream) {
  if (int rc = lib_check(d)) return rc;
  if (!terms || !idx || !theta || !rhs || n < 1) return fail(PERCNN_ERR_INVALID, "bad argument");
  DeviceGuard guard(d->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = percnn_library_points(d);
  if (d->dtype == PERCNN_F32)
    k_lib_theta<float><<<lib_grid(n), 256, 0, st>>>(static_cast<const float*>(terms), total, idx, n, theta, rhs);
  else
    k_lib_theta<double><<<lib_grid(n), 256, 0, st>>>(static_cast<const double*>(terms), total, idx, n, theta, rhs);
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

}  // extern "C"
