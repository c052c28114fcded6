// Translation unit: the initial-state generator ("upscaler"), its adjoint and the IC loss.
// Stand-alone entry points (no plan): percnn_upscaler_sizes / _fwd / _bwd, percnn_mse_fwd / _bwd.
#include <cstdlib>
#include <type_traits>

#include "kernels_upscaler.cuh"
#include "plan.h"

using namespace percnn;
using namespace percnn::up;

namespace {

struct UpGeom {
  int ndim, C, K, layers, S2, act;
  Grid low, mid, own, out;      // mid: the planes stage 2 reads; own: the mid planes whose sums this call owns
  int64_t mid_elems, own_elems; // per-buffer element counts (all channels)
  RawOff raw;
  PrepOff prep;
  int n1, n2;                   // lengths of the two sum vectors
  size_t off_prep, off_sums1, off_sums2, off_partials, off_gmid, ws_bytes;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

int up_geom(const percnn_upscaler_t* d, UpGeom* u, bool need_device = true) {
  if (!d) return fail(PERCNN_ERR_INVALID, "null upscaler descriptor");
  if (d->ndim != 2 && d->ndim != 3) return fail(PERCNN_ERR_INVALID, "ndim must be 2 or 3");
  if (d->dtype != PERCNN_F32 && d->dtype != PERCNN_F64) return fail(PERCNN_ERR_INVALID, "dtype must be f32 or f64");
  if (d->channels != 8 && d->channels != 16) return fail(PERCNN_ERR_UNSUPPORTED, "upscaler channels must be 8 (GS2D:31) or 16 (BUR1:44)");
  if (d->act != 0 && d->act != 1) return fail(PERCNN_ERR_INVALID, "act must be 0 (sigmoid) or 1 (tanh)");
  if (d->layers != 1 && d->layers != 2) return fail(PERCNN_ERR_INVALID, "layers must be 1 or 2");
  if (d->layers == 2 && d->stride2 != 1 && d->stride2 != 2) return fail(PERCNN_ERR_INVALID, "stride2 must be 1 or 2");
  for (int i = 0; i < 3; ++i)
    if (d->low_extent[i] < 1 || d->low_extent[i] > (1 << 14)) return fail(PERCNN_ERR_INVALID, "bad low-resolution extents");
  if (d->ndim == 2 && d->low_extent[0] != 1) return fail(PERCNN_ERR_INVALID, "2-D needs low_extent[0] == 1");
  if (need_device && !percnn_device_ok(d->device)) return fail(PERCNN_ERR_NO_DEVICE, "no sm_100 CUDA device with this ordinal");
  const bool three = d->ndim == 3;
  u->ndim = d->ndim;
  u->C = d->channels;
  u->K = three ? 125 : 25;
  u->layers = d->layers;
  u->S2 = d->layers == 2 ? d->stride2 : 1;
  u->act = d->act;
  const int S2 = u->S2;
  const int Dl = int(d->low_extent[0]), Hl = int(d->low_extent[1]), Wl = int(d->low_extent[2]);
  const int Dm = three ? 2 * Dl : 1, Hm = 2 * Hl, Wm = 2 * Wl;
  const int Do = three ? S2 * Dm : 1, Ho = S2 * Hm, Wo = S2 * Wm;
  u->low = Grid{Dl, Hl, Wl, 0, Dl, int64_t(Dl) * Hl * Wl};
  int z0 = 0, nz = Do;
  if (d->out_nz != 0) {
    if (!three) return fail(PERCNN_ERR_INVALID, "slab ranges (out_nz != 0) are for 3-D upscalers");
    if (d->out_z0 < 0 || d->out_nz < 1 || d->out_z0 + d->out_nz > Do || (d->out_z0 % S2) || (d->out_nz % S2))
      return fail(PERCNN_ERR_INVALID, "bad output plane range (must lie inside the grid and be a multiple of stride2)");
    z0 = int(d->out_z0);
    nz = int(d->out_nz);
  }
  const int64_t dense = int64_t(nz) * Ho * Wo;
  if (d->out_field_stride != 0 && d->out_field_stride < dense) return fail(PERCNN_ERR_INVALID, "out_field_stride smaller than one field");
  u->out = Grid{Do, Ho, Wo, z0, nz, d->out_field_stride != 0 ? d->out_field_stride : dense};
  // mid planes stage 2 reads for output planes [z0, z0 + nz): iz = (o + 2 - k) / S2, clamped to the grid
  int mlo = 0, mhi = Dm;
  int olo = 0, ohi = Dm;
  if (three) {
    olo = z0 / S2;
    ohi = (z0 + nz) / S2;
    mlo = u->layers == 2 ? (z0 - 2 < 0 ? 0 : (z0 - 2 + S2 - 1) / S2) : olo;
    mhi = u->layers == 2 ? (z0 + nz - 1 + 2) / S2 + 1 : ohi;
    if (mhi > Dm) mhi = Dm;
  }
  const int64_t mplane = int64_t(Hm) * Wm;
  u->mid = Grid{Dm, Hm, Wm, mlo, mhi - mlo, int64_t(mhi - mlo) * mplane};
  u->own = Grid{Dm, Hm, Wm, olo, ohi - olo, int64_t(ohi - olo) * mplane};
  u->mid_elems = u->mid.cstride * u->C;
  u->own_elems = u->own.cstride * u->C;
  u->raw = raw_off(u->C, u->K, u->layers);
  u->prep = prep_off(u->C, u->K);
  u->n1 = u->C * 2 * u->K + u->C;
  u->n2 = u->layers == 2 ? 2 * u->C * u->K + 2 : 2 * u->C + 2;
  const size_t elt = d->dtype == PERCNN_F32 ? 4 : 8;
  size_t o = 0;
  u->off_prep = o;
  o += align256(size_t(u->prep.total) * elt);
  u->off_sums1 = o;
  o += align256(size_t(u->n1) * 8);
  u->off_sums2 = o;
  o += align256(size_t(u->n2) * 8);
  u->off_partials = o;
  o += align256(size_t(kCorrMaxVB) * size_t(u->n1 > u->n2 ? u->n1 : u->n2) * 8);
  u->off_gmid = o;
  o += align256(size_t(u->own_elems) * elt);
  u->ws_bytes = o;
  return PERCNN_OK;
}

int blocks_for(int64_t items) {
  int64_t b = (items + kThreads - 1) / kThreads;
  if (b > 148 * 8) b = 148 * 8;
  return int(b < 1 ? 1 : b);
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) != cudaSuccess)
    return fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(upscaler) failed");
  return PERCNN_OK;
}

template <typename T, int NDIM, int C>
int fwd_t(const UpGeom& u, const T* raw, const T* low, T* mid, T* out, char* ws, cudaStream_t st) {
  T* prep = reinterpret_cast<T*>(ws + u.off_prep);
  k_up_prep<T><<<8, 256, 0, st>>>(raw, prep, C, u.K, u.layers);
  PERCNN_CUDA(cudaGetLastError());
  const size_t sm1 = size_t(2 * u.K * C + C + 2 * C + 2) * sizeof(T);
  const int64_t tiles1 = int64_t(u.mid.nz) * u.mid.H * (u.mid.W / 2);
  if (u.layers == 1) {
    auto kern = k_up_l1<T, NDIM, C, true>;
    if (int rc = set_smem(kern, sm1)) return rc;
    kern<<<blocks_for(tiles1), kThreads, sm1, st>>>(u.low, u.mid, u.out, u.act, prep, low, mid, out);
    PERCNN_CUDA(cudaGetLastError());
    return PERCNN_OK;
  }
  auto k1 = k_up_l1<T, NDIM, C, false>;
  if (int rc = set_smem(k1, sm1)) return rc;
  k1<<<blocks_for(tiles1), kThreads, sm1, st>>>(u.low, u.mid, u.out, u.act, prep, low, mid, out);
  PERCNN_CUDA(cudaGetLastError());
  const size_t sm2 = size_t(C * u.K * 2 + 2) * sizeof(T);
  if (u.S2 == 1) {
    auto k2 = k_up_l2<T, NDIM, C, 1>;
    if (int rc = set_smem(k2, sm2)) return rc;
    k2<<<blocks_for(int64_t(u.out.nz) * u.out.H * ((u.out.W + 1) / 2)), kThreads, sm2, st>>>(u.mid, u.out, prep, mid, out);
  } else {
    auto k2 = k_up_l2<T, NDIM, C, 2>;
    if (int rc = set_smem(k2, sm2)) return rc;
    k2<<<blocks_for(int64_t(u.out.nz) * u.out.H * ((u.out.W + 3) / 4)), kThreads, sm2, st>>>(u.mid, u.out, prep, mid, out);
  }
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

template <typename T>
int corr(const Grid& ga, const Grid& gb, int ndim, int S, int CA, int CB, int single, const T* A, const T* B, int n, double* partials,
         double* sums, cudaStream_t st) {
  const int64_t nrows = int64_t(ga.nz) * ga.H;
  int64_t nvb = nrows < kCorrMaxVB ? nrows : kCorrMaxVB;   // the kernel is latency-bound per warp: as many blocks as rows
  if (nvb < 1) nvb = 1;
  if (CA >= 4)
    k_up_corr<T, 4><<<int(nvb), kThreads, 0, st>>>(ga, gb, ndim, S, CA, CB, single, A, B, partials);
  else
    k_up_corr<T, 2><<<int(nvb), kThreads, 0, st>>>(ga, gb, ndim, S, CA, CB, single, A, B, partials);
  PERCNN_CUDA(cudaGetLastError());
  k_up_fold<<<(n + 31) / 32, 256, 0, st>>>(partials, int(nvb), n, sums);
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

// The dominant correlation (3-D, stride 1, dL/dh0 x the 8 stored activations, fp32) on shared-memory tiles.
int corr3(const Grid& ga, const Grid& gb, const float* A, const float* B, int n, double* partials, double* sums, cudaStream_t st) {
  const int ntx = (ga.W + C3_TX - 1) / C3_TX, nty = (ga.H + C3_TY - 1) / C3_TY;
  int zc = 1;                                         // planes per chunk: as few as keeps the block count <= kCorrMaxVB
  while (int64_t(ntx) * nty * ((ga.nz + zc - 1) / zc) > kCorrMaxVB) ++zc;
  const int nblocks = ntx * nty * ((ga.nz + zc - 1) / zc);
  const size_t smem = size_t(C3_SMEM_FLOATS) * sizeof(float);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    if (cudaFuncSetAttribute(k_up_corr3, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(k_up_corr3) failed");
    attr_set[dev] = true;
  }
  k_up_corr3<<<nblocks, C3_THREADS, smem, st>>>(ga, gb, ntx, nty, zc, A, B, partials);
  PERCNN_CUDA(cudaGetLastError());
  k_up_fold<<<(n + 31) / 32, 256, 0, st>>>(partials, nblocks, n, sums);
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

template <typename T, int NDIM, int C>
int bwd_t(const UpGeom& u, const T* raw, const T* low, const T* mid, const T* g, T* gp, int accumulate, char* ws, cudaStream_t st) {
  T* prep = reinterpret_cast<T*>(ws + u.off_prep);
  double* sums1 = reinterpret_cast<double*>(ws + u.off_sums1);
  double* sums2 = reinterpret_cast<double*>(ws + u.off_sums2);
  double* partials = reinterpret_cast<double*>(ws + u.off_partials);
  T* gm = reinterpret_cast<T*>(ws + u.off_gmid);
  k_up_prep<T><<<8, 256, 0, st>>>(raw, prep, C, u.K, u.layers);
  PERCNN_CUDA(cudaGetLastError());
  // `mid` holds planes [mid.z0, ...); the owned planes start (own.z0 - mid.z0) planes in
  const T* mid_own = mid + int64_t(u.own.z0 - u.mid.z0) * u.mid.H * u.mid.W;
  Grid mid_as_own = u.own;            // same planes as `own`, but addressed inside the (larger) mid buffer
  mid_as_own.cstride = u.mid.cstride;
  if (u.layers == 1) {
    k_up_gmid1<T, C><<<blocks_for(int64_t(u.own.nz) * u.own.H * u.own.W), kThreads, 0, st>>>(mid_as_own, u.out, u.act, prep, u.K, mid_own, g, gm);
    PERCNN_CUDA(cudaGetLastError());
    // gmid1 indexes mid and gm with the same linear index i, so both must share a channel stride: enforced by
    // own == mid for one-layer nets (no slab mode there).
    if (int rc = corr<T>(u.out, mid_as_own, u.ndim, 1, 2, C, 1, g, mid_own, u.n2, partials, sums2, st)) return rc;
  } else {
    const size_t smg = size_t(2 * u.K * C) * sizeof(T);
    const int64_t tiles = int64_t(u.own.nz) * u.own.H * (u.own.W / 2);
    // gmid2 reads act'(mid) at the owned cells through the mid buffer's own geometry and writes gm in `own` geometry
    if (u.S2 == 1) {
      auto kg = k_up_gmid2<T, NDIM, C, 1>;
      if (int rc = set_smem(kg, smg)) return rc;
      kg<<<blocks_for(tiles), kThreads, smg, st>>>(u.own, u.out, u.act, prep, mid_own, g, gm, u.mid.cstride);
    } else {
      auto kg = k_up_gmid2<T, NDIM, C, 2>;
      if (int rc = set_smem(kg, smg)) return rc;
      kg<<<blocks_for(tiles), kThreads, smg, st>>>(u.own, u.out, u.act, prep, mid_own, g, gm, u.mid.cstride);
    }
    PERCNN_CUDA(cudaGetLastError());
    if constexpr (std::is_same<T, float>::value && NDIM == 3 && C == C3_CB) {
      if (u.S2 == 1 && !getenv("PERCNN_UP_GENERIC_CORR")) {
        if (int rc = corr3(u.out, u.mid, g, mid, u.n2, partials, sums2, st)) return rc;
      } else if (int rc = corr<T>(u.out, u.mid, u.ndim, u.S2, 2, C, 0, g, mid, u.n2, partials, sums2, st)) {
        return rc;
      }
    } else if (int rc = corr<T>(u.out, u.mid, u.ndim, u.S2, 2, C, 0, g, mid, u.n2, partials, sums2, st)) {
      return rc;
    }
  }
  if (int rc = corr<T>(u.own, u.low, u.ndim, 2, C, 2, 0, gm, low, u.n1, partials, sums1, st)) return rc;
  k_up_finish<T><<<32, 256, 0, st>>>(raw, sums1, sums2, C, u.K, u.layers, gp, accumulate);
  PERCNN_CUDA(cudaGetLastError());
  if (u.layers == 2) {
    k_up_finish_w3<T><<<2 * C, 128, 0, st>>>(raw, sums2, C, u.K, gp, accumulate);
    PERCNN_CUDA(cudaGetLastError());
  }
  return PERCNN_OK;
}

template <typename T>
int fwd_dispatch(const UpGeom& u, const void* raw, const void* low, void* mid, void* out, char* ws, cudaStream_t st) {
  const T* r = static_cast<const T*>(raw);
  const T* l = static_cast<const T*>(low);
  T* m = static_cast<T*>(mid);
  T* o = static_cast<T*>(out);
  if (u.ndim == 2) return u.C == 8 ? fwd_t<T, 2, 8>(u, r, l, m, o, ws, st) : fwd_t<T, 2, 16>(u, r, l, m, o, ws, st);
  return u.C == 8 ? fwd_t<T, 3, 8>(u, r, l, m, o, ws, st) : fwd_t<T, 3, 16>(u, r, l, m, o, ws, st);
}
template <typename T>
int bwd_dispatch(const UpGeom& u, const void* raw, const void* low, const void* mid, const void* g, void* gp, int acc, char* ws,
                 cudaStream_t st) {
  const T* r = static_cast<const T*>(raw);
  const T* l = static_cast<const T*>(low);
  const T* m = static_cast<const T*>(mid);
  const T* gg = static_cast<const T*>(g);
  T* p = static_cast<T*>(gp);
  if (u.ndim == 2) return u.C == 8 ? bwd_t<T, 2, 8>(u, r, l, m, gg, p, acc, ws, st) : bwd_t<T, 2, 16>(u, r, l, m, gg, p, acc, ws, st);
  return u.C == 8 ? bwd_t<T, 3, 8>(u, r, l, m, gg, p, acc, ws, st) : bwd_t<T, 3, 16>(u, r, l, m, gg, p, acc, ws, st);
}

}  // namespace

extern "C" {

int percnn_upscaler_sizes(const percnn_upscaler_t* d, int64_t* nparams, int64_t* mid_elems, int64_t* out_extent,
                          size_t* ws_bytes) {
  UpGeom u;
  if (int rc = up_geom(d, &u, /*need_device=*/false)) return rc;   // pure host arithmetic: usable for planning without a GPU
  if (nparams) *nparams = u.raw.total;
  if (mid_elems) *mid_elems = u.mid_elems;
  if (out_extent) {
    out_extent[0] = u.out.D;
    out_extent[1] = u.out.H;
    out_extent[2] = u.out.W;
  }
  if (ws_bytes) *ws_bytes = u.ws_bytes;
  return PERCNN_OK;
}

int percnn_upscaler_fwd(const percnn_upscaler_t* d, const void* params, const void* low, void* mid, void* h0, void* ws,
                        void* stream) {
  UpGeom u;
  if (int rc = up_geom(d, &u)) return rc;
  if (!params || !low || !mid || !h0 || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(d->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->dtype == PERCNN_F32 ? fwd_dispatch<float>(u, params, low, mid, h0, static_cast<char*>(ws), st)
                                : fwd_dispatch<double>(u, params, low, mid, h0, static_cast<char*>(ws), st);
}

int percnn_upscaler_bwd(const percnn_upscaler_t* d, const void* params, const void* low, const void* mid, const void* g_h0,
                        void* g_params, int accumulate, void* ws, void* stream) {
  UpGeom u;
  if (int rc = up_geom(d, &u)) return rc;
  if (!params || !low || !mid || !g_h0 || !g_params || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(d->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->dtype == PERCNN_F32
             ? bwd_dispatch<float>(u, params, low, mid, g_h0, g_params, accumulate, static_cast<char*>(ws), st)
             : bwd_dispatch<double>(u, params, low, mid, g_h0, g_params, accumulate, static_cast<char*>(ws), st);
}

size_t percnn_mse_workspace_bytes(void) { return 1024 * sizeof(double); }

int percnn_mse_fwd(int dtype, int device, const void* a, const void* b, int64_t n, void* loss_out, void* ws, void* stream) {
  if (dtype != PERCNN_F32 && dtype != PERCNN_F64) return fail(PERCNN_ERR_INVALID, "dtype must be f32 or f64");
  if (!a || !b || !loss_out || !ws || n < 1) return fail(PERCNN_ERR_INVALID, "bad argument");
  if (!percnn_device_ok(device)) return fail(PERCNN_ERR_NO_DEVICE, "no sm_100 CUDA device with this ordinal");
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t nb = (n + kThreads - 1) / kThreads;
  if (nb > 1024) nb = 1024;
  double* partials = static_cast<double*>(ws);
  if (dtype == PERCNN_F32) {
    k_mse_partial<float><<<int(nb), kThreads, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(b), n, partials);
    k_mse_finish<float><<<1, 32, 0, st>>>(partials, int(nb), 1.0 / double(n), static_cast<float*>(loss_out));
  } else {
    k_mse_partial<double><<<int(nb), kThreads, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b), n, partials);
    k_mse_finish<double><<<1, 32, 0, st>>>(partials, int(nb), 1.0 / double(n), static_cast<double*>(loss_out));
  }
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

int percnn_mse_bwd(int dtype, int device, const void* a, const void* b, int64_t n, const void* gscale, void* g, int accumulate,
                   void* stream) {
  if (dtype != PERCNN_F32 && dtype != PERCNN_F64) return fail(PERCNN_ERR_INVALID, "dtype must be f32 or f64");
  if (!a || !b || !g || n < 1) return fail(PERCNN_ERR_INVALID, "bad argument");
  if (!percnn_device_ok(device)) return fail(PERCNN_ERR_NO_DEVICE, "no sm_100 CUDA device with this ordinal");
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = blocks_for(n);
  if (dtype == PERCNN_F32)
    k_mse_grad<float><<<nb, kThreads, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(b), n, 2.0 / double(n),
                                               static_cast<const float*>(gscale), static_cast<float*>(g), accumulate);
  else
    k_mse_grad<double><<<nb, kThreads, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b), n, 2.0 / double(n),
                                                static_cast<const double*>(gscale), static_cast<double*>(g), accumulate);
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

}  // extern "C"
