// C-ABI of libpercnn_b200.so (see include/percnn_b200.h).  Single translation unit: the kernels share
// one __constant__ parameter block, so everything is compiled together for sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "kernels_generic.cuh"
#include "kernels_phys_loss.cuh"
#include "plan.h"

using namespace percnn;

namespace {
thread_local std::string g_err;

// The six __constant__ parameter slots exist once per device (constant memory belongs to the context).
constexpr int kMaxDevices = 64;
std::mutex g_slot_mutex;
bool g_slot_used[kMaxDevices][kPrepSlots] = {};

}  // namespace

namespace percnn {
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
}  // namespace percnn

namespace {

// workspace: [accumulators | counter | per-block partials | two state-sized scratch buffers]
size_t ws_header_bytes(const percnn_plan* p) {   // bytes zeroed by param_grads_begin
  return is_k5(p) ? (size_t(p->nparams) * sizeof(double) + 255) / 256 * 256 : kWsPartials;
}
size_t ws_states_off(const percnn_plan* p) {
  if (!is_k5(p)) return kWsStates;
  const size_t part = size_t(k5_blocks(p)) * size_t(p->nparams) * sizeof(float);
  return ws_header_bytes(p) + (part + 255) / 256 * 256;
}


// per_sm: resident blocks per SM to aim for.  The adjoint kernels end in a 6..22-value block reduction, so they
// get fewer, longer-running blocks (more cells per thread to amortise it).
int generic_grid(const percnn_plan* p, int per_sm = 8) {
  const int64_t ncell = int64_t(p->g.D) * p->g.H * p->g.W;
  int64_t blocks = (ncell + kGenericThreads - 1) / kGenericThreads;
  const int64_t cap = int64_t(p->sm_count) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  return int(blocks < 1 ? 1 : blocks);
}

template <typename T>
int step_fwd_t(percnn_plan* p, const T* src, T* dst, cudaStream_t st) {
  const Geom& g = p->g;
  const int grid = generic_grid(p);
  const bool branch = (p->desc.flags & PERCNN_FLAG_EVAL_BRANCH) != 0;
  if (p->desc.cell == PERCNN_CELL_PI && p->desc.ksize == 1) {
    if (g.ndim == 3) {
      if (branch)
        k_pi_k1_fwd<T, 3, true><<<grid, kGenericThreads, 0, st>>>(g, p->slot, p->desc.hidden, src, dst);
      else
        k_pi_k1_fwd<T, 3, false><<<grid, kGenericThreads, 0, st>>>(g, p->slot, p->desc.hidden, src, dst);
    } else {
      if (branch)
        k_pi_k1_fwd<T, 2, true><<<grid, kGenericThreads, 0, st>>>(g, p->slot, p->desc.hidden, src, dst);
      else
        k_pi_k1_fwd<T, 2, false><<<grid, kGenericThreads, 0, st>>>(g, p->slot, p->desc.hidden, src, dst);
    }
  } else if (p->desc.cell == PERCNN_CELL_BURGERS) {
    k_burgers_fwd<T><<<grid, kGenericThreads, 0, st>>>(g, p->slot, src, dst);
  } else if (p->desc.cell == PERCNN_CELL_LO) {
    k_lo_fwd<T><<<grid, kGenericThreads, 0, st>>>(g, p->slot, src, dst);
  } else {
    return fail(PERCNN_ERR_UNSUPPORTED, "no generic forward kernel for this cell");
  }
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

template <typename T>
const void* multi_step_kernel(const percnn_plan* p) {
  const int cell = p->desc.cell;
  if (cell == PERCNN_CELL_PI)
    return p->g.ndim == 3 ? (const void*)k_multi_step<T, 3, 0> : (const void*)k_multi_step<T, 2, 0>;
  if (cell == PERCNN_CELL_BURGERS) return (const void*)k_multi_step<T, 2, 1>;
  return (const void*)k_multi_step<T, 2, 2>;
}
// The persistent one-block-per-SM kernels pay a grid barrier per step; that only wins while the ping-pong working
// set stays in L2 and a per-step launch would be latency-bound.  Larger non-TMA plans (fp64 3-D, W % 128 != 0)
// run one generic kernel per step.
constexpr size_t kMultiStepMaxStateBytes = size_t(40) << 20;
constexpr size_t kSmallSlabBytes = size_t(12) << 20;   // per-rank state (incl. ghosts) below which a slab rollout runs persistently
bool multi_step_eligible(const percnn_plan* p) {
  if (p->use_tma || is_k5(p) || p->desc.slab_ghost) return false;
  if (state_bytes(p) > kMultiStepMaxStateBytes) return false;
  if (p->desc.cell == PERCNN_CELL_PI && (p->desc.flags & PERCNN_FLAG_EVAL_BRANCH)) return false;
  return p->multi_grid > 0 && p->d_sync != nullptr;
}
template <typename T>
int launch_multi_step(percnn_plan* p, const void* h0, void* tape, void* ping, void* pong, void* final_state, int nsteps,
                      cudaStream_t st) {
  MultiStepArgs<T> m;
  m.h0 = static_cast<const T*>(h0);
  m.tape = static_cast<T*>(tape);
  m.ping = static_cast<T*>(ping);
  m.pong = static_cast<T*>(pong);
  m.final_state = static_cast<T*>(final_state);
  m.nsteps = nsteps;
  m.stride = p->state_elems;
  PERCNN_CUDA(cudaMemsetAsync(p->d_sync, 0, sizeof(unsigned), st));
  Geom g = p->g;
  int slot = p->slot;
  unsigned* counter = p->d_sync;
  void* args[] = {&g, &slot, &m, &counter};
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  int grid = int((ncell + kMultiThreads - 1) / kMultiThreads);
  if (grid > p->multi_grid) grid = p->multi_grid;
  PERCNN_CUDA(cudaLaunchCooperativeKernel(multi_step_kernel<T>(p), dim3(grid), dim3(kMultiThreads), args, 0, st));
  p->launches++;
  return PERCNN_OK;
}

template <typename T>
const void* multi_bwd_kernel(const percnn_plan* p) {
  const int cell = p->desc.cell;
  if (cell == PERCNN_CELL_PI)
    return p->g.ndim == 3 ? (const void*)k_multi_step_bwd<T, 3, 0> : (const void*)k_multi_step_bwd<T, 2, 0>;
  if (cell == PERCNN_CELL_BURGERS) return (const void*)k_multi_step_bwd<T, 2, 1>;
  return (const void*)k_multi_step_bwd<T, 2, 2>;
}
// Whole backward rollout in one cooperative launch (see k_multi_step_bwd).  G_init must already hold dL/dh_nsteps.
template <typename T>
int launch_multi_bwd(percnn_plan* p, const void* tape, const void* g_tape, const uint8_t* gmask, const void* g_init,
                     void* ping, void* pong, void* g_h0, const percnn_data_loss_t* dl, int64_t dl_n, int nsteps, char* ws,
                     cudaStream_t st) {
  static_assert(sizeof(MultiBwdArgs<T>) < 3500, "kernel parameter space");
  MultiBwdArgs<T> m;
  memset(&m, 0, sizeof(m));
  m.tape = static_cast<const T*>(tape);
  m.g_tape = static_cast<const T*>(g_tape);
  m.g_init = static_cast<const T*>(g_init);
  m.ping = static_cast<T*>(ping);
  m.pong = static_cast<T*>(pong);
  m.g_h0 = static_cast<T*>(g_h0);
  m.nsteps = nsteps;
  m.stride = p->state_elems;
  m.inj.s = 1;
  if (g_tape)
    for (int s = 0; s < nsteps; ++s)
      if (gmask[s]) {
        m.gmask[s >> 5] |= 1u << (s & 31);
        m.g_slots++;
      }
  if (dl) {
    InjectHost ih;
    ih.target = dl->target;
    ih.gscale = dl->gscale;
    ih.stride = dl->stride;
    ih.n_total = dl_n;
    m.inj = make_inject<T>(p, &ih);
    for (int s = 0; s < nsteps; ++s)
      if (dl->sel[s]) {
        m.selmask[s >> 5] |= 1u << (s & 31);
        m.inj_slots++;
      }
    if (m.inj_slots == 0) m.inj.target = nullptr;
  }
  PERCNN_CUDA(cudaMemsetAsync(p->d_sync, 0, sizeof(unsigned), st));
  Geom g = p->g;
  int slot = p->slot;
  unsigned* counter = p->d_sync;
  double* partials = reinterpret_cast<double*>(ws + kWsPartials);
  unsigned* red_counter = reinterpret_cast<unsigned*>(ws + kWsCounter);
  double* acc = reinterpret_cast<double*>(ws + kWsAcc);
  void* args[] = {&g, &slot, &m, &counter, &partials, &red_counter, &acc};
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  int grid = int((ncell + kMultiThreads - 1) / kMultiThreads);
  if (grid > p->multi_bwd_grid) grid = p->multi_bwd_grid;
  PERCNN_CUDA(cudaLaunchCooperativeKernel(multi_bwd_kernel<T>(p), dim3(grid), dim3(kMultiThreads), args, 0, st));
  p->launches++;
  return PERCNN_OK;
}

int step_fwd_any(percnn_plan* p, const void* src, void* dst, cudaStream_t st) {
  if (p->use_tma) return tma_fwd_launch(p, static_cast<const float*>(src), static_cast<float*>(dst), 0, p->g.D, st, nullptr);
  if (is_k5(p)) return k5_step_fwd(p, static_cast<const float*>(src), static_cast<float*>(dst), st);
  return p->elt == 4 ? step_fwd_t<float>(p, static_cast<const float*>(src), static_cast<float*>(dst), st)
                     : step_fwd_t<double>(p, static_cast<const double*>(src), static_cast<double*>(dst), st);
}

template <typename T>
int step_bwd_t(percnn_plan* p, const T* h, const T* gout, const T* gadd, T* gin, char* ws, cudaStream_t st,
               const InjectHost* ih) {
  const Geom& g = p->g;
  const int grid = generic_grid(p, 2);
  const Inject<T> inj = make_inject<T>(p, ih);
  double* acc = reinterpret_cast<double*>(ws + kWsAcc);
  unsigned* counter = reinterpret_cast<unsigned*>(ws + kWsCounter);
  double* partials = reinterpret_cast<double*>(ws + kWsPartials);
  if (p->desc.cell == PERCNN_CELL_PI && p->desc.ksize == 1) {
    if (g.ndim == 3)
      k_pi_k1_bwd<T, 3><<<grid, kGenericThreads, 0, st>>>(g, p->slot, h, gout, gadd, gin, partials, counter, acc, inj);
    else
      k_pi_k1_bwd<T, 2><<<grid, kGenericThreads, 0, st>>>(g, p->slot, h, gout, gadd, gin, partials, counter, acc, inj);
  } else if (p->desc.cell == PERCNN_CELL_BURGERS) {
    k_burgers_bwd<T><<<grid, kGenericThreads, 0, st>>>(g, p->slot, h, gout, gadd, gin, partials, counter, acc, inj);
  } else if (p->desc.cell == PERCNN_CELL_LO) {
    k_lo_bwd<T><<<grid, kGenericThreads, 0, st>>>(g, p->slot, h, gout, gadd, gin, partials, counter, acc, inj);
  } else {
    return fail(PERCNN_ERR_UNSUPPORTED, "adjoint of the 5x5 Pi-block cell is not implemented yet");
  }
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

int step_bwd_any(percnn_plan* p, const void* h, const void* gout, const void* gadd, void* gin, void* ws, cudaStream_t st,
                 const SlabLink* link = nullptr, const InjectHost* ih = nullptr) {
  if (p->use_tma)
    return tma_bwd_launch(p, static_cast<const float*>(h), static_cast<const float*>(gout), static_cast<const float*>(gadd),
                          static_cast<float*>(gin), static_cast<char*>(ws), st, link, ih);
  if (link) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo adjoint needs a TMA plan");
  if (is_k5(p)) {
    char* w = static_cast<char*>(ws);
    int rc = k5_step_bwd(p, static_cast<const float*>(h), static_cast<const float*>(gout), static_cast<const float*>(gadd),
                         static_cast<float*>(gin), reinterpret_cast<double*>(w),
                         reinterpret_cast<float*>(w + ws_header_bytes(p)), st);
    if (rc) return rc;
    if (ih && ih->target) {
      // the 5x5 adjoint sits at its register limit; the (rare, 1/s^2-sized) loss injection runs as its own pass
      k_inject_only<float><<<generic_grid(p), kGenericThreads, 0, st>>>(p->g, static_cast<const float*>(h),
                                                                         static_cast<float*>(gin), make_inject<float>(p, ih));
      PERCNN_CUDA(cudaGetLastError());
      p->launches++;
    }
    return PERCNN_OK;
  }
  return p->elt == 4 ? step_bwd_t<float>(p, static_cast<const float*>(h), static_cast<const float*>(gout),
                                         static_cast<const float*>(gadd), static_cast<float*>(gin),
                                         static_cast<char*>(ws), st, ih)
                     : step_bwd_t<double>(p, static_cast<const double*>(h), static_cast<const double*>(gout),
                                          static_cast<const double*>(gadd), static_cast<double*>(gin),
                                          static_cast<char*>(ws), st, ih);
}

// percnn_slab_link_t -> SlabLink (checks included)
int resolve_link(const percnn_plan* p, const percnn_slab_link_t* link, SlabLink* l) {
  if (!p->use_tma || !p->desc.slab_ghost) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo steps need a slab-mode TMA plan");
  if (p->g.D < 4) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo steps need at least 4 planes per rank");
  if (!link->peer_lo_out || !link->peer_hi_out || !link->my_flags || !link->peer_lo_flags || !link->peer_hi_flags || !link->scratch)
    return fail(PERCNN_ERR_INVALID, "incomplete slab link");
  l->peer_lo_dst = static_cast<float*>(link->peer_lo_out);
  l->peer_hi_dst = static_cast<float*>(link->peer_hi_out);
  l->my_flags = link->my_flags;
  l->post_lo_flag = link->peer_lo_flags + 1;   // I provide the lower neighbour's UPPER ghosts
  l->post_hi_flag = link->peer_hi_flags + 0;   // and the upper neighbour's LOWER ghosts
  l->scratch = link->scratch;
  l->epoch_wait = link->epoch;
  l->epoch_post = link->epoch + 1;
  l->flush_prev = (link->flags & PERCNN_SLAB_FLUSH_PREV) != 0;
  l->defer_late = (link->flags & PERCNN_SLAB_DEFER_LATE) != 0;
  l->peer_lo_src = static_cast<float*>(link->peer_lo_in);
  l->peer_hi_src = static_cast<float*>(link->peer_hi_in);
  if (l->flush_prev && (!l->peer_lo_src || !l->peer_hi_src))
    return fail(PERCNN_ERR_INVALID, "PERCNN_SLAB_FLUSH_PREV needs peer_lo_in / peer_hi_in");
  return PERCNN_OK;
}

}  // namespace

extern "C" {

int percnn_abi_version(void) { return PERCNN_ABI_VERSION; }

const char* percnn_last_error(void) { return g_err.c_str(); }

int percnn_device_ok(int device) {
  // cudaGetDeviceProperties costs ~1 ms; the stand-alone loss entry points validate their device on every call,
  // so the answer is cached per ordinal (a device's compute capability cannot change under a live process).
  static std::mutex mu;
  static signed char cache[64];   // 0 = unknown, 1 = sm_100, -1 = anything else
  if (device < 0) return 0;
  if (device < 64) {
    std::lock_guard<std::mutex> lk(mu);
    if (cache[device] != 0) return cache[device] > 0 ? 1 : 0;
  }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return 0;
  const int ok = major == 10 ? 1 : 0;
  if (device < 64) {
    std::lock_guard<std::mutex> lk(mu);
    cache[device] = ok ? 1 : -1;
  }
  return ok;
}

int percnn_plan_create(const percnn_desc_t* d, percnn_plan_t** out) {
  if (!d || !out) return fail(PERCNN_ERR_INVALID, "null descriptor or output pointer");
  *out = nullptr;
  if (d->abi_version != PERCNN_ABI_VERSION) return fail(PERCNN_ERR_INVALID, "abi_version mismatch");
  if (d->ndim != 2 && d->ndim != 3) return fail(PERCNN_ERR_INVALID, "ndim must be 2 or 3");
  if (d->dtype != PERCNN_F32 && d->dtype != PERCNN_F64) return fail(PERCNN_ERR_INVALID, "dtype must be f32 or f64");
  if (d->extent[1] < 1 || d->extent[2] < 1 || d->extent[0] < 1) return fail(PERCNN_ERR_INVALID, "extents must be >= 1");
  if (d->ndim == 2 && d->extent[0] != 1) return fail(PERCNN_ERR_INVALID, "2-D plans need extent[0] == 1");
  if (d->extent[0] > (1 << 20) || d->extent[1] > (1 << 20) || d->extent[2] > (1 << 20))
    return fail(PERCNN_ERR_INVALID, "extent too large");
  if (d->cell == PERCNN_CELL_PI) {
    if (d->ksize != 1 && d->ksize != 5) return fail(PERCNN_ERR_INVALID, "Pi conv kernel size must be 1 or 5");
    if (d->hidden < 1 || d->hidden > kMaxHidden) return fail(PERCNN_ERR_INVALID, "hidden channels must be in 1..16");
    if (d->ksize == 5) {
      if (d->ndim != 2 || d->dtype != PERCNN_F32)
        return fail(PERCNN_ERR_UNSUPPORTED, "the 5x5 Pi-block cell exists in 2-D fp32 only (BUR1/LO1)");
      if (d->hidden % 2) return fail(PERCNN_ERR_UNSUPPORTED, "5x5 Pi-block needs an even channel count");
    }
  } else if (d->cell == PERCNN_CELL_BURGERS || d->cell == PERCNN_CELL_LO) {
    if (d->ndim != 2) return fail(PERCNN_ERR_INVALID, "Stage-3 physics cells are 2-D");
    if (!(d->dx > 0)) return fail(PERCNN_ERR_INVALID, "dx must be positive");
  } else {
    return fail(PERCNN_ERR_INVALID, "unknown cell kind");
  }
  if (d->slab_ghost != 0 && d->slab_ghost != 1) return fail(PERCNN_ERR_INVALID, "slab_ghost must be 0 or 1");
  if (!percnn_device_ok(d->device)) return fail(PERCNN_ERR_NO_DEVICE, "no sm_100 CUDA device with this ordinal");

  percnn_plan* p = new (std::nothrow) percnn_plan();
  if (!p) return fail(PERCNN_ERR_INVALID, "out of host memory");
  p->desc = *d;
  p->elt = d->dtype == PERCNN_F32 ? 4 : 8;
  Geom& g = p->g;
  g.ndim = d->ndim;
  g.D = int(d->extent[0]);
  g.H = int(d->extent[1]);
  g.W = int(d->extent[2]);
  g.ghost = d->slab_ghost ? 2 : 0;
  if (d->ndim == 3) {
    g.plane = int64_t(g.H) * g.W;
    g.field = int64_t(g.D + 2 * g.ghost) * g.plane;
  } else {
    g.plane = g.W;
    g.field = int64_t(g.H + 2 * g.ghost) * g.W;
  }
  p->state_elems = 2 * g.field;
  p->pd = PrepDesc{d->cell, d->ndim, d->ksize, d->hidden, d->coef_mode, d->flags, d->mu_up, d->dt, d->dx};
  if (d->cell == PERCNN_CELL_PI) {
    p->nparams = PiPacking(d->ndim, d->ksize, d->hidden).total();
    p->nred = d->ksize == 1 ? kRedPiK1 : 0;
  } else if (d->cell == PERCNN_CELL_BURGERS) {
    p->nparams = 6 + 75;
    p->nred = kRedBurgers;
  } else {
    p->nparams = ((d->flags & PERCNN_FLAG_LO_C6) ? 13 : 12) + 25;
    p->nred = kRedLO;
  }
  int rc = PERCNN_OK;
  DeviceGuard guard(d->device);   // the caller's current device is restored on return
  do {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != d->device) { rc = fail(PERCNN_ERR_CUDA, "cudaSetDevice failed"); break; }
    if (d->device >= kMaxDevices) { rc = fail(PERCNN_ERR_INVALID, "device ordinal too large"); break; }
    if (cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, d->device) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaDeviceGetAttribute failed"); break; }
    {
      std::lock_guard<std::mutex> lk(g_slot_mutex);
      for (int s = 0; s < kPrepSlots; ++s)
        if (!g_slot_used[d->device][s]) { g_slot_used[d->device][s] = true; p->slot = s; break; }
    }
    if (p->slot < 0) { rc = fail(PERCNN_ERR_INVALID, "too many live plans on this device (6 parameter slots)"); break; }
    if (cudaMalloc(&p->d_prep, sizeof(PrepBlock)) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaMalloc(prep) failed"); break; }
    if (is_k5(p)) {
      rc = k5_setup(p);
      if (rc) break;
    }
    if (!is_k5(p) && !d->slab_ghost && !getenv("PERCNN_NO_MULTISTEP")) {
      int coop = 0, per_sm = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device);
      const void* fn = p->elt == 4 ? multi_step_kernel<float>(p) : multi_step_kernel<double>(p);
      if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kMultiThreads, 0) == cudaSuccess && per_sm > 0) {
        if (cudaMalloc(&p->d_sync, 256) == cudaSuccess) p->multi_grid = p->sm_count;   // one block per SM
      }
      const void* bfn = p->elt == 4 ? multi_bwd_kernel<float>(p) : multi_bwd_kernel<double>(p);
      per_sm = 0;
      if (p->multi_grid > 0 && !getenv("PERCNN_NO_MULTISTEP_BWD") &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bfn, kMultiThreads, 0) == cudaSuccess && per_sm > 0)
        p->multi_bwd_grid = p->sm_count;
    }
    if (d->slab_ghost && d->cell == PERCNN_CELL_PI && d->ksize == 1 && d->ndim == 3 && d->dtype == PERCNN_F32 &&
        !(d->flags & PERCNN_FLAG_EVAL_BRANCH) && !getenv("PERCNN_NO_MULTISTEP")) {
      int coop = 0, per_sm = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device);
      if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_multi_step_slab, kMultiThreads, 0) == cudaSuccess &&
          per_sm > 0 && cudaMalloc(&p->d_sync, 256) == cudaSuccess)
        p->multi_slab_grid = p->sm_count;
    }
    p->use_tma = d->cell == PERCNN_CELL_PI && d->ksize == 1 && d->ndim == 3 && d->dtype == PERCNN_F32 &&
                 !(d->flags & (PERCNN_FLAG_NO_TMA | PERCNN_FLAG_EVAL_BRANCH)) && g.W % 128 == 0 &&
                 g.H >= 4 && g.D >= 4;
    if (p->use_tma) {
      rc = tma_fwd_setup(p);
      if (!rc) rc = tma_bwd_setup(p);
      if (rc) break;
    }
    if (tile2d_eligible(p)) {
      rc = tile2d_setup(p);
      if (rc) break;
    }
  } while (0);
  if (rc != PERCNN_OK) {
    std::string keep = g_err;
    percnn_plan_destroy(p);
    g_err = keep;
    return rc;
  }
  *out = p;
  return PERCNN_OK;
}

int percnn_plan_destroy(percnn_plan_t* p) {
  if (!p) return PERCNN_OK;
  DeviceGuard guard(p->desc.device);
  if (p->d_prep) cudaFree(p->d_prep);
  if (p->d_k5w) cudaFree(p->d_k5w);
  if (p->d_sync) cudaFree(p->d_sync);
  if (p->h_params_dev) cudaFree(p->h_params_dev);
  if (p->h_states) cudaFree(p->h_states);
  if (p->h_stream) cudaStreamDestroy(p->h_stream);
  if (p->slot >= 0 && p->desc.device >= 0 && p->desc.device < kMaxDevices) {
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    g_slot_used[p->desc.device][p->slot] = false;
  }
  delete p;
  return PERCNN_OK;
}

int64_t percnn_param_count(const percnn_plan_t* p) { return p ? p->nparams : -1; }
int64_t percnn_state_elems(const percnn_plan_t* p) { return p ? p->state_elems : -1; }
int percnn_plan_uses_tma(const percnn_plan_t* p) { return p && p->use_tma ? 1 : 0; }
int64_t percnn_plan_launch_count(const percnn_plan_t* p) { return p ? p->launches : -1; }
int percnn_plan_uses_tile2d(const percnn_plan_t* p) { return p && p->use_tile2d ? p->t2_K : 0; }
int percnn_plan_slab_persistent(const percnn_plan_t* p) {
  return p && p->multi_slab_grid > 0 && state_bytes(p) <= kSmallSlabBytes && !getenv("PERCNN_SLAB_NO_PERSISTENT") ? 1 : 0;
}

size_t percnn_workspace_bytes(const percnn_plan_t* p, int nsteps) {
  (void)nsteps;
  if (!p) return 0;
  return ws_states_off(p) + 2 * ((state_bytes(p) + 255) / 256 * 256);
}

int percnn_params_load(percnn_plan_t* p, const void* params, void* stream) {
  if (!p || !params) return fail(PERCNN_ERR_INVALID, "null plan or params");
  DeviceGuard guard(p->desc.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->elt == 4)
    k_prep<float><<<1, 256, 0, st>>>(static_cast<const float*>(params), p->pd, p->d_prep, p->d_k5w);
  else
    k_prep<double><<<1, 256, 0, st>>>(static_cast<const double*>(params), p->pd, p->d_prep, p->d_k5w);
  PERCNN_CUDA(cudaGetLastError());
  // every translation unit keeps its own copy of the constant block; fill the ones this plan's kernels read
  PERCNN_CUDA(cudaMemcpyToSymbolAsync(c_prep, p->d_prep, sizeof(PrepBlock), size_t(p->slot) * sizeof(PrepBlock),
                                      cudaMemcpyDeviceToDevice, st));
  if (p->use_tma) {
    PERCNN_CUDA(tma_fwd_load_prep(p->d_prep, p->slot, st));
    PERCNN_CUDA(tma_bwd_load_prep(p->d_prep, p->slot, st));
  }
  if (is_k5(p)) PERCNN_CUDA(k5_load_prep(p->d_prep, p->slot, st));
  if (p->use_tile2d) PERCNN_CUDA(tile2d_load_prep(p->d_prep, p->slot, st));
  p->launches++;
  return PERCNN_OK;
}

int percnn_step_fwd(percnn_plan_t* p, const void* h_in, void* h_out, void* stream) {
  if (!p || !h_in || !h_out) return fail(PERCNN_ERR_INVALID, "null argument");
  if (h_in == h_out) return fail(PERCNN_ERR_INVALID, "step_fwd cannot run in place");
  DeviceGuard guard(p->desc.device);
  return step_fwd_any(p, h_in, h_out, static_cast<cudaStream_t>(stream));
}

// One classical RK4 step of a Stage-3 physics cell (RCNNCell.forward_rk4, BUR3:159-206): four launches of the fused
// right-hand side.  `ws` provides the two scratch states.
int percnn_step_rk4(percnn_plan_t* p, const void* h_in, void* h_out, void* ws, void* stream) {
  if (!p || !h_in || !h_out || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (h_in == h_out) return fail(PERCNN_ERR_INVALID, "step_rk4 cannot run in place");
  if (p->desc.cell != PERCNN_CELL_BURGERS && p->desc.cell != PERCNN_CELL_LO)
    return fail(PERCNN_ERR_UNSUPPORTED, "forward_rk4 exists for the Stage-3 physics cells only (BUR3:159, LO3:153)");
  if (p->desc.slab_ghost) return fail(PERCNN_ERR_UNSUPPORTED, "no slab-mode RK4");
  DeviceGuard guard(p->desc.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sb = (state_bytes(p) + 255) / 256 * 256;
  char* k[2] = {static_cast<char*>(ws) + ws_states_off(p), static_cast<char*>(ws) + ws_states_off(p) + sb};
  const int grid = generic_grid(p);
  const double dt = p->desc.dt;
  const double a_of[4] = {0.0, dt / 2.0, dt / 2.0, dt};
  for (int stage = 0; stage < 4; ++stage) {
    const void* kprev = stage == 0 ? nullptr : k[(stage - 1) & 1];
    void* kout = k[stage & 1];
    const bool burgers = p->desc.cell == PERCNN_CELL_BURGERS;
    if (p->elt == 8) {
      if (burgers) k_rk4_stage<double, 1><<<grid, kGenericThreads, 0, st>>>(p->g, p->slot, static_cast<const double*>(h_in), static_cast<const double*>(kprev), a_of[stage], static_cast<double*>(kout), static_cast<double*>(h_out), stage);
      else k_rk4_stage<double, 2><<<grid, kGenericThreads, 0, st>>>(p->g, p->slot, static_cast<const double*>(h_in), static_cast<const double*>(kprev), a_of[stage], static_cast<double*>(kout), static_cast<double*>(h_out), stage);
    } else {
      if (burgers) k_rk4_stage<float, 1><<<grid, kGenericThreads, 0, st>>>(p->g, p->slot, static_cast<const float*>(h_in), static_cast<const float*>(kprev), float(a_of[stage]), static_cast<float*>(kout), static_cast<float*>(h_out), stage);
      else k_rk4_stage<float, 2><<<grid, kGenericThreads, 0, st>>>(p->g, p->slot, static_cast<const float*>(h_in), static_cast<const float*>(kprev), float(a_of[stage]), static_cast<float*>(kout), static_cast<float*>(h_out), stage);
    }
    PERCNN_CUDA(cudaGetLastError());
    p->launches++;
  }
  return PERCNN_OK;
}

// Forward step restricted to interior planes [z_lo, z_hi) of a 3-D TMA plan (slab mode overlap: the
// planes that need ghosts are launched after the halo exchange, the rest before).  Not part of the
// reference surface; used by percnn_b200.halo.
int percnn_step_fwd_range(percnn_plan_t* p, const void* h_in, void* h_out, int z_lo, int z_hi, void* stream) {
  if (!p || !h_in || !h_out) return fail(PERCNN_ERR_INVALID, "null argument");
  if (!p->use_tma) return fail(PERCNN_ERR_UNSUPPORTED, "step_fwd_range needs a TMA plan");
  if (z_lo < 0 || z_hi > p->g.D || z_lo >= z_hi) return fail(PERCNN_ERR_INVALID, "bad plane range");
  DeviceGuard guard(p->desc.device);
  return tma_fwd_launch(p, static_cast<const float*>(h_in), static_cast<float*>(h_out), z_lo, z_hi,
                        static_cast<cudaStream_t>(stream), nullptr);
}

// One fused slab step: a single z-march (direction = parity of the epoch) whose boundary planes are also stored
// into the neighbours' ghost planes through the peer mapping; flags raised from inside the kernel.
int percnn_step_fwd_fused_halo(percnn_plan_t* p, const void* h_in, void* h_out, const percnn_slab_link_t* link, void* stream) {
  if (!p || !h_in || !h_out || !link) return fail(PERCNN_ERR_INVALID, "null argument");
  SlabLink l;
  int rc = resolve_link(p, link, &l);
  if (rc) return rc;
  DeviceGuard guard(p->desc.device);
  return tma_fwd_launch(p, static_cast<const float*>(h_in), static_cast<float*>(h_out), 0, p->g.D,
                        static_cast<cudaStream_t>(stream), &l);
}

// Whole slab-mode forward rollout: `nsteps` fused steps issued from one host call.  Step s reads
// buf[cur ^ (s & 1)] and writes the other buffer (locally and, for the boundary planes, at both peers).
int percnn_slab_rollout_fwd(percnn_plan_t* p, const percnn_slab_ring_t* ring, int cur, int nsteps, uint32_t epoch,
                            void* stream) {
  if (!p || !ring) return fail(PERCNN_ERR_INVALID, "null argument");
  if (nsteps < 0 || (cur != 0 && cur != 1)) return fail(PERCNN_ERR_INVALID, "bad nsteps / cur");
  if (!ring->buf[0] || !ring->buf[1]) return fail(PERCNN_ERR_INVALID, "incomplete slab ring");
  DeviceGuard guard(p->desc.device);
  // Small slabs (a few microseconds of work per step): the whole rollout as ONE persistent cooperative kernel whose
  // grid barrier is also the halo hand-shake (k_multi_step_slab); the per-step kernels pay a kernel boundary plus
  // the NVLink flag latency per step, which such slabs cannot hide.
  if (nsteps >= 2 && percnn_plan_slab_persistent(p)) {
    if (!ring->peer_lo_buf[0] || !ring->peer_lo_buf[1] || !ring->peer_hi_buf[0] || !ring->peer_hi_buf[1] || !ring->my_flags ||
        !ring->peer_lo_flags || !ring->peer_hi_flags || !ring->scratch)
      return fail(PERCNN_ERR_INVALID, "incomplete slab ring");
    SlabMultiArgs a;
    for (int i = 0; i < 2; ++i) {
      a.buf[i] = static_cast<float*>(ring->buf[i]);
      a.peer_lo[i] = static_cast<float*>(ring->peer_lo_buf[i]);
      a.peer_hi[i] = static_cast<float*>(ring->peer_hi_buf[i]);
    }
    a.my_flags = ring->my_flags;
    a.post_lo_flag = ring->peer_lo_flags + 1;
    a.post_hi_flag = ring->peer_hi_flags + 0;
    a.err = ring->scratch + 1;
    a.epoch0 = epoch;
    a.spin_limit = p->flag_spin_limit;
    a.nsteps = nsteps;
    a.cur = cur;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PERCNN_CUDA(cudaMemsetAsync(p->d_sync, 0, 8, st));
    Geom g = p->g;
    int slot = p->slot;
    unsigned* counter = p->d_sync;
    void* args[] = {&g, &slot, &a, &counter};
    const int64_t ncell = int64_t(g.D) * g.H * g.W;
    int grid = int((ncell + kMultiThreads - 1) / kMultiThreads);
    if (grid > p->multi_slab_grid) grid = p->multi_slab_grid;
    PERCNN_CUDA(cudaLaunchCooperativeKernel((const void*)k_multi_step_slab, dim3(grid), dim3(kMultiThreads), args, 0, st));
    p->launches++;
    return PERCNN_OK;
  }
  for (int s = 0; s < nsteps; ++s) {
    const int src = cur ^ (s & 1), dst = src ^ 1;
    percnn_slab_link_t link;
    link.peer_lo_out = ring->peer_lo_buf[dst];
    link.peer_hi_out = ring->peer_hi_buf[dst];
    link.my_flags = ring->my_flags;
    link.peer_lo_flags = ring->peer_lo_flags;
    link.peer_hi_flags = ring->peer_hi_flags;
    link.scratch = ring->scratch;
    link.epoch = epoch + uint32_t(s);
    link.flags = (s > 0 ? PERCNN_SLAB_FLUSH_PREV : 0) | (s + 1 < nsteps ? PERCNN_SLAB_DEFER_LATE : 0);
    link.peer_lo_in = ring->peer_lo_buf[src];
    link.peer_hi_in = ring->peer_hi_buf[src];
    SlabLink l;
    int rc = resolve_link(p, &link, &l);
    if (rc) return rc;
    rc = tma_fwd_launch(p, static_cast<const float*>(ring->buf[src]), static_cast<float*>(ring->buf[dst]), 0, p->g.D,
                        static_cast<cudaStream_t>(stream), &l);
    if (rc) return rc;
  }
  return PERCNN_OK;
}

// The persistent small-slab rollout with K time steps per halo exchange (k_multi_step_slab_tb): `wide` holds four
// buffers [2][D + 4K][H][W] of this rank and the neighbours' mappings of theirs.
int percnn_slab_rollout_fwd_blocked(percnn_plan_t* p, const percnn_slab_ring_t* ring, const percnn_slab_wide_t* wide, int cur,
                                    int nsteps, uint32_t epoch, void* stream) {
  if (!p || !ring || !wide) return fail(PERCNN_ERR_INVALID, "null argument");
  if (nsteps < 1 || (cur != 0 && cur != 1)) return fail(PERCNN_ERR_INVALID, "bad nsteps / cur");
  if (!percnn_plan_slab_persistent(p)) return fail(PERCNN_ERR_UNSUPPORTED, "plan has no persistent slab kernel (slab too large, or not a slab plan)");
  if (wide->k < 1 || 2 * wide->k > p->g.D) return fail(PERCNN_ERR_INVALID, "need 1 <= k and 2 k <= planes per rank");
  if (!ring->buf[0] || !ring->buf[1] || !ring->peer_lo_buf[0] || !ring->peer_lo_buf[1] || !ring->peer_hi_buf[0] ||
      !ring->peer_hi_buf[1] || !ring->my_flags || !ring->peer_lo_flags || !ring->peer_hi_flags || !ring->scratch)
    return fail(PERCNN_ERR_INVALID, "incomplete slab ring");
  SlabBlockedArgs a;
  for (int i = 0; i < 2; ++i) {
    a.buf[i] = static_cast<float*>(ring->buf[i]);
    a.peer_lo[i] = static_cast<float*>(ring->peer_lo_buf[i]);
    a.peer_hi[i] = static_cast<float*>(ring->peer_hi_buf[i]);
  }
  for (int i = 0; i < 4; ++i) {
    if (!wide->buf[i] || !wide->peer_lo_buf[i] || !wide->peer_hi_buf[i]) return fail(PERCNN_ERR_INVALID, "incomplete wide buffers");
    a.w[i] = static_cast<float*>(wide->buf[i]);
    a.wlo[i] = static_cast<float*>(wide->peer_lo_buf[i]);
    a.whi[i] = static_cast<float*>(wide->peer_hi_buf[i]);
  }
  a.my_flags = ring->my_flags;
  a.post_lo_flag = ring->peer_lo_flags + 1;
  a.post_hi_flag = ring->peer_hi_flags + 0;
  a.err = ring->scratch + 1;
  a.epoch0 = epoch;
  a.spin_limit = p->flag_spin_limit;
  a.nsteps = nsteps;
  a.cur = cur;
  a.K = wide->k;
  DeviceGuard guard(p->desc.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PERCNN_CUDA(cudaMemsetAsync(p->d_sync, 0, 8, st));
  Geom g = p->g;
  int slot = p->slot;
  unsigned* counter = p->d_sync;
  void* args[] = {&g, &slot, &a, &counter};
  const int64_t ncell = int64_t(g.D) * g.H * g.W;
  int grid = int((ncell + kMultiThreads - 1) / kMultiThreads);
  if (grid > p->multi_slab_grid) grid = p->multi_slab_grid;
  PERCNN_CUDA(cudaLaunchCooperativeKernel((const void*)k_multi_step_slab_tb, dim3(grid), dim3(kMultiThreads), args, 0, st));
  p->launches++;
  return PERCNN_OK;
}

// Taped slab rollout: step t reads tape slot t and writes slot t+1 (and the boundary planes of the peers' slot t+1).
int percnn_slab_rollout_tape(percnn_plan_t* p, void* tape, void* peer_lo_tape, void* peer_hi_tape,
                             const percnn_slab_ring_t* ring, int nsteps, uint32_t epoch, void* stream) {
  if (!p || !tape || !peer_lo_tape || !peer_hi_tape || !ring) return fail(PERCNN_ERR_INVALID, "null argument");
  if (nsteps < 0) return fail(PERCNN_ERR_INVALID, "nsteps must be >= 0");
  DeviceGuard guard(p->desc.device);
  const size_t sb = state_bytes(p);
  for (int t = 0; t < nsteps; ++t) {
    percnn_slab_link_t link;
    link.peer_lo_out = static_cast<char*>(peer_lo_tape) + size_t(t + 1) * sb;
    link.peer_hi_out = static_cast<char*>(peer_hi_tape) + size_t(t + 1) * sb;
    link.my_flags = ring->my_flags;
    link.peer_lo_flags = ring->peer_lo_flags;
    link.peer_hi_flags = ring->peer_hi_flags;
    link.scratch = ring->scratch;
    link.epoch = epoch + uint32_t(t);
    link.flags = (t > 0 ? PERCNN_SLAB_FLUSH_PREV : 0) | (t + 1 < nsteps ? PERCNN_SLAB_DEFER_LATE : 0);
    link.peer_lo_in = static_cast<char*>(peer_lo_tape) + size_t(t) * sb;
    link.peer_hi_in = static_cast<char*>(peer_hi_tape) + size_t(t) * sb;
    SlabLink l;
    int rc = resolve_link(p, &link, &l);
    if (rc) return rc;
    rc = tma_fwd_launch(p, reinterpret_cast<const float*>(static_cast<char*>(tape) + size_t(t) * sb),
                        reinterpret_cast<float*>(static_cast<char*>(tape) + size_t(t + 1) * sb), 0, p->g.D,
                        static_cast<cudaStream_t>(stream), &l);
    if (rc) return rc;
  }
  return PERCNN_OK;
}

// Adjoint counterpart of percnn_step_fwd_fused_halo: the gradient's boundary planes are mirrored into the
// neighbours' ghost planes of their g_in buffers.
int percnn_step_bwd_fused_halo(percnn_plan_t* p, const void* h_in, const void* g_out, const void* g_add, void* g_in,
                               void* ws, const percnn_slab_link_t* link, void* stream) {
  if (!link) return fail(PERCNN_ERR_INVALID, "null argument");
  return percnn_step_bwd_loss(p, h_in, g_out, g_add, nullptr, 1, 0, nullptr, g_in, ws, link, stream);
}

// Adjoint step with the fused data-loss gradient of state h_in injected (and, with `link`, the fused halo exchange).
int percnn_step_bwd_loss(percnn_plan_t* p, const void* h_in, const void* g_out, const void* g_add,
                         const void* target_frame, int stride, int64_t n_total, const void* gscale, void* g_in,
                         void* ws, const percnn_slab_link_t* link, void* stream) {
  if (!p || !h_in || !g_out || !g_in || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (g_out == g_in) return fail(PERCNN_ERR_INVALID, "step_bwd cannot run in place");
  InjectHost ih;
  if (target_frame) {
    if (stride < 1) return fail(PERCNN_ERR_INVALID, "data-loss stride must be >= 1");
    if (n_total < 1) return fail(PERCNN_ERR_INVALID, "data-loss n_total must be >= 1 for a single step");
    ih.target = target_frame;
    ih.gscale = gscale;
    ih.stride = stride;
    ih.n_total = n_total;
  }
  DeviceGuard guard(p->desc.device);
  if (!link) return step_bwd_any(p, h_in, g_out, g_add, g_in, ws, static_cast<cudaStream_t>(stream), nullptr, &ih);
  SlabLink l;
  int rc = resolve_link(p, link, &l);
  if (rc) return rc;
  return step_bwd_any(p, h_in, g_out, g_add, g_in, ws, static_cast<cudaStream_t>(stream), &l, &ih);
}

// Whole slab-mode backward rollout.  ring->buf[0] must hold G_nsteps (ghosts exchanged); step t = nsteps-1 .. 0 reads
// the gradient from buf[b], the stored state from tape slot t and writes buf[b ^ 1]; dL/dh_0 ends in
// buf[nsteps & 1].  `g_tape` (nullable): dense dL/d(tape[t]) in the ghosted layout, slot t added at step t.
// `loss` (nullable): fused data loss, target frames of the selected states packed in increasing step order.
// The caller zeroes the accumulator before (percnn_param_grads_begin) and all-reduces / finishes after.
int percnn_slab_rollout_bwd(percnn_plan_t* p, const void* tape, const void* g_tape, const percnn_data_loss_t* loss,
                            const percnn_slab_ring_t* ring, int nsteps, uint32_t epoch, void* ws, void* stream) {
  if (!p || !tape || !ring || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (nsteps < 1) return fail(PERCNN_ERR_INVALID, "nsteps must be >= 1");
  if (!ring->buf[0] || !ring->buf[1]) return fail(PERCNN_ERR_INVALID, "incomplete slab ring");
  if (loss && (!loss->target || !loss->sel || loss->stride < 1 || loss->n_total < 1))
    return fail(PERCNN_ERR_INVALID, "slab data loss needs target, sel, stride >= 1 and the GLOBAL n_total");
  DeviceGuard guard(p->desc.device);
  const size_t sb = state_bytes(p);
  const size_t frame_bytes = loss ? size_t(2 * lowres_field_elems(p, loss->stride)) * p->elt : 0;
  int slot = 0;
  if (loss)
    for (int t = 0; t < nsteps; ++t) slot += loss->sel[t] ? 1 : 0;
  int b = 0;
  for (int t = nsteps - 1; t >= 0; --t) {
    const int nxt = b ^ 1;
    percnn_slab_link_t link;
    link.peer_lo_out = ring->peer_lo_buf[nxt];
    link.peer_hi_out = ring->peer_hi_buf[nxt];
    link.my_flags = ring->my_flags;
    link.peer_lo_flags = ring->peer_lo_flags;
    link.peer_hi_flags = ring->peer_hi_flags;
    link.scratch = ring->scratch;
    link.epoch = epoch + uint32_t(nsteps - 1 - t);
    link.flags = (t < nsteps - 1 ? PERCNN_SLAB_FLUSH_PREV : 0) | (t > 0 ? PERCNN_SLAB_DEFER_LATE : 0);
    link.peer_lo_in = ring->peer_lo_buf[b];
    link.peer_hi_in = ring->peer_hi_buf[b];
    SlabLink l;
    int rc = resolve_link(p, &link, &l);
    if (rc) return rc;
    InjectHost ih;
    if (loss && loss->sel[t]) {
      --slot;
      ih.target = static_cast<const char*>(loss->target) + size_t(slot) * frame_bytes;
      ih.gscale = loss->gscale;
      ih.stride = loss->stride;
      ih.n_total = loss->n_total;
    }
    const char* add = g_tape ? static_cast<const char*>(g_tape) + size_t(t) * sb : nullptr;
    rc = step_bwd_any(p, static_cast<const char*>(tape) + size_t(t) * sb, ring->buf[b], add, ring->buf[nxt], ws,
                      static_cast<cudaStream_t>(stream), &l, &ih);
    if (rc) return rc;
    b = nxt;
  }
  return PERCNN_OK;
}

namespace {
int check_data_loss(const percnn_plan* p, const percnn_data_loss_t* dl, int nsteps, int* nsel_out, int64_t* n_out) {
  if (!dl->target || !dl->sel) return fail(PERCNN_ERR_INVALID, "data loss needs a target and a selection mask");
  if (dl->stride < 1) return fail(PERCNN_ERR_INVALID, "data-loss stride must be >= 1");
  if (dl->reserved != 0) return fail(PERCNN_ERR_INVALID, "percnn_data_loss_t.reserved must be 0");
  if (dl->n_total < 0) return fail(PERCNN_ERR_INVALID, "data-loss n_total must be >= 0");
  int nsel = 0;
  for (int s = 0; s <= nsteps; ++s) nsel += dl->sel[s] ? 1 : 0;
  if (nsel == 0) return fail(PERCNN_ERR_INVALID, "data loss selects no state");
  *nsel_out = nsel;
  *n_out = dl->n_total > 0 ? dl->n_total : int64_t(nsel) * 2 * lowres_field_elems(p, dl->stride);
  return PERCNN_OK;
}
}  // namespace

// Forward value of the fused data loss: one small launch per selected state (reads the sampled points only).
int percnn_data_loss_fwd(percnn_plan_t* p, const void* tape, int nsteps, const percnn_data_loss_t* dl, void* loss_out,
                         void* ws, void* stream) {
  if (!p || !tape || !dl || !loss_out || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (nsteps < 0) return fail(PERCNN_ERR_INVALID, "nsteps must be >= 0");
  int nsel = 0;
  int64_t n = 0;
  int rc = check_data_loss(p, dl, nsteps, &nsel, &n);
  if (rc) return rc;
  DeviceGuard guard(p->desc.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(ws);
  double* acc = reinterpret_cast<double*>(w + kWsAcc);
  unsigned* counter = reinterpret_cast<unsigned*>(w + kWsCounter);
  double* partials = reinterpret_cast<double*>(w + kWsPartials);
  PERCNN_CUDA(cudaMemsetAsync(w, 0, kWsPartials, st));
  const int64_t lf = lowres_field_elems(p, dl->stride);
  int64_t blocks = (2 * lf + kGenericThreads - 1) / kGenericThreads;
  if (blocks > 256) blocks = 256;   // partials must fit the smallest workspace header (5x5 plans, hidden = 2)
  const size_t sb = state_bytes(p);
  int slot = 0;
  for (int s = 0; s <= nsteps; ++s) {
    if (!dl->sel[s]) continue;
    InjectHost ih;
    ih.target = static_cast<const char*>(dl->target) + size_t(slot) * size_t(2 * lf) * p->elt;
    ih.stride = dl->stride;
    ih.n_total = n;
    const char* h = static_cast<const char*>(tape) + size_t(s) * sb;
    if (p->elt == 4)
      k_data_loss<float><<<int(blocks), kGenericThreads, 0, st>>>(p->g, reinterpret_cast<const float*>(h),
                                                                  make_inject<float>(p, &ih), partials, counter, acc);
    else
      k_data_loss<double><<<int(blocks), kGenericThreads, 0, st>>>(p->g, reinterpret_cast<const double*>(h),
                                                                   make_inject<double>(p, &ih), partials, counter, acc);
    PERCNN_CUDA(cudaGetLastError());
    p->launches++;
    ++slot;
  }
  if (p->elt == 4)
    k_data_loss_finish<float><<<1, 32, 0, st>>>(acc, 1.0 / double(n), static_cast<float*>(loss_out));
  else
    k_data_loss_finish<double><<<1, 32, 0, st>>>(acc, 1.0 / double(n), static_cast<double*>(loss_out));
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

int percnn_param_grads_begin(percnn_plan_t* p, void* ws, void* stream) {
  if (!p || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(p->desc.device);
  PERCNN_CUDA(cudaMemsetAsync(ws, 0, ws_header_bytes(p), static_cast<cudaStream_t>(stream)));
  return PERCNN_OK;
}

int percnn_step_bwd(percnn_plan_t* p, const void* h_in, const void* g_out, const void* g_add, void* g_in, void* ws,
                    void* stream) {
  if (!p || !h_in || !g_out || !g_in || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (g_out == g_in) return fail(PERCNN_ERR_INVALID, "step_bwd cannot run in place");
  DeviceGuard guard(p->desc.device);
  return step_bwd_any(p, h_in, g_out, g_add, g_in, ws, static_cast<cudaStream_t>(stream));
}

int percnn_param_grads_finish(percnn_plan_t* p, const void* params, void* param_grads, void* ws, void* stream) {
  if (!p || !params || !param_grads || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(p->desc.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double* acc = reinterpret_cast<const double*>(static_cast<char*>(ws) + kWsAcc);
  if (is_k5(p)) return k5_grads_finish(p, static_cast<const float*>(params), acc, static_cast<float*>(param_grads), st);
  if (p->elt == 4)
    k_finish_small<float><<<1, 256, 0, st>>>(static_cast<const float*>(params), acc, p->pd, int(p->nparams),
                                             static_cast<float*>(param_grads));
  else
    k_finish_small<double><<<1, 256, 0, st>>>(static_cast<const double*>(params), acc, p->pd, int(p->nparams),
                                              static_cast<double*>(param_grads));
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

int percnn_rollout_fwd(percnn_plan_t* p, const void* h0, void* traj, const uint8_t* emit, int nsteps, void* h_final,
                       void* tape, void* ws, void* stream) {
  if (!p || !h0) return fail(PERCNN_ERR_INVALID, "null plan or h0");
  if (nsteps < 0) return fail(PERCNN_ERR_INVALID, "nsteps must be >= 0");
  if (traj && !emit) return fail(PERCNN_ERR_INVALID, "traj given without an emit mask");
  if (!tape && !ws) return fail(PERCNN_ERR_INVALID, "rollout_fwd needs a workspace unless a tape is given");
  if (p->desc.slab_ghost) return fail(PERCNN_ERR_UNSUPPORTED, "slab-mode rollouts go through percnn_slab_rollout_fwd (halo exchange between steps)");
  DeviceGuard guard(p->desc.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sb = state_bytes(p);
  char* tp = static_cast<char*>(tape);
  char* pp[2] = {nullptr, nullptr};
  if (ws) {
    pp[0] = static_cast<char*>(ws) + ws_states_off(p);
    pp[1] = pp[0] + (sb + 255) / 256 * 256;
  }
  if (tp && tp != h0) PERCNN_CUDA(cudaMemcpyAsync(tp, h0, sb, cudaMemcpyDeviceToDevice, st));
  // 2-D cells: the whole rollout on shared-memory tiles, K time steps per pass, one launch (tape, emitted frames and/or
  // the final state)
  if (nsteps >= 2 && nsteps <= 4096 && p->use_tile2d && !(tp && traj) && (tp || pp[0]) && (tp || traj || h_final)) {
    int rc = tile2d_rollout(p, tp ? tp : h0, tp ? nullptr : h_final, tp, traj, emit, pp[0], pp[1], nsteps, st);
    if (rc) return rc;
    if (tp && h_final) PERCNN_CUDA(cudaMemcpyAsync(h_final, tp + size_t(nsteps) * sb, sb, cudaMemcpyDeviceToDevice, st));
    return PERCNN_OK;
  }
  // small grids: one persistent cooperative launch for the whole rollout (tape mode, or final-state-only mode)
  if (nsteps >= 2 && multi_step_eligible(p) && !traj && (tp || (h_final && pp[0]))) {
    int rc = p->elt == 4 ? launch_multi_step<float>(p, h0, tp, pp[0], pp[1], h_final, nsteps, st)
                         : launch_multi_step<double>(p, h0, tp, pp[0], pp[1], h_final, nsteps, st);
    if (rc) return rc;
    // tape mode never touches h_final inside the kernel: hand the last slot over, as the per-step path does
    if (tp && h_final) PERCNN_CUDA(cudaMemcpyAsync(h_final, tp + size_t(nsteps) * sb, sb, cudaMemcpyDeviceToDevice, st));
    return PERCNN_OK;
  }
  const char* cur = tp ? tp : static_cast<const char*>(h0);
  int slot = 0, flip = 0;
  for (int s = 0; s < nsteps; ++s) {
    const bool emitted = traj && emit[s];
    char* dst;
    if (tp)
      dst = tp + size_t(s + 1) * sb;
    else if (emitted)
      dst = static_cast<char*>(traj) + size_t(slot) * sb;
    else if (s == nsteps - 1 && h_final)
      dst = static_cast<char*>(h_final);
    else {
      dst = pp[flip];
      flip ^= 1;
    }
    int rc = step_fwd_any(p, cur, dst, st);
    if (rc) return rc;
    if (emitted) {
      if (tp) PERCNN_CUDA(cudaMemcpyAsync(static_cast<char*>(traj) + size_t(slot) * sb, dst, sb, cudaMemcpyDeviceToDevice, st));
      ++slot;
    }
    cur = dst;
  }
  if (h_final && cur != h_final) PERCNN_CUDA(cudaMemcpyAsync(h_final, cur, sb, cudaMemcpyDeviceToDevice, st));
  return PERCNN_OK;
}

int percnn_rollout_bwd(percnn_plan_t* p, const void* params, const void* tape, const void* g_tape, const uint8_t* gmask,
                       int nsteps, void* g_h0, void* param_grads, void* ws, void* stream) {
  return percnn_rollout_bwd_loss(p, params, tape, g_tape, gmask, nullptr, nsteps, g_h0, param_grads, ws, stream);
}

int percnn_rollout_bwd_loss(percnn_plan_t* p, const void* params, const void* tape, const void* g_tape,
                            const uint8_t* gmask, const percnn_data_loss_t* dl, int nsteps, void* g_h0, void* param_grads,
                            void* ws, void* stream) {
  if (!p || !params || !tape || !g_h0 || !param_grads || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (nsteps < 1) return fail(PERCNN_ERR_INVALID, "nsteps must be >= 1");
  if (g_tape && !gmask) return fail(PERCNN_ERR_INVALID, "g_tape given without a mask");
  int dl_nsel = 0;
  int64_t dl_n = 0;
  if (dl) {
    int rcl = check_data_loss(p, dl, nsteps, &dl_nsel, &dl_n);
    if (rcl) return rcl;
  }
  DeviceGuard guard(p->desc.device);
  const size_t dl_frame_bytes = dl ? size_t(2 * lowres_field_elems(p, dl->stride)) * p->elt : 0;
  int dl_slot = dl_nsel;   // walks the packed target frames backwards, like `slot` does for g_tape
  auto inject_for = [&](int s, InjectHost* ih) {
    *ih = InjectHost();
    if (dl && dl->sel[s]) {
      --dl_slot;
      ih->target = static_cast<const char*>(dl->target) + size_t(dl_slot) * dl_frame_bytes;
      ih->gscale = dl->gscale;
      ih->stride = dl->stride;
      ih->n_total = dl_n;
    }
  };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sb = state_bytes(p);
  char* w = static_cast<char*>(ws);
  char* pp[2] = {w + ws_states_off(p), w + ws_states_off(p) + (sb + 255) / 256 * 256};
  int rc = percnn_param_grads_begin(p, ws, stream);
  if (rc) return rc;
  // compact slot index of each masked state
  int nmasked = 0;
  if (g_tape)
    for (int s = 0; s <= nsteps; ++s) nmasked += gmask[s] ? 1 : 0;
  int slot = nmasked;
  const char* gt = static_cast<const char*>(g_tape);
  // G_{nsteps}
  const char* G;
  if (g_tape && gmask[nsteps]) {
    --slot;
    G = gt + size_t(slot) * sb;
  } else {
    PERCNN_CUDA(cudaMemsetAsync(pp[0], 0, sb, st));
    G = pp[0];
  }
  InjectHost ih;
  inject_for(nsteps, &ih);
  if (ih.target) {
    // the last state has no adjoint step of its own: scatter its loss gradient into G_nsteps
    if (G != pp[0]) {
      PERCNN_CUDA(cudaMemcpyAsync(pp[0], G, sb, cudaMemcpyDeviceToDevice, st));
      G = pp[0];
    }
    const int grid = generic_grid(p);
    const char* hT = static_cast<const char*>(tape) + size_t(nsteps) * sb;
    if (p->elt == 4)
      k_inject_only<float><<<grid, kGenericThreads, 0, st>>>(p->g, reinterpret_cast<const float*>(hT),
                                                             reinterpret_cast<float*>(pp[0]), make_inject<float>(p, &ih));
    else
      k_inject_only<double><<<grid, kGenericThreads, 0, st>>>(p->g, reinterpret_cast<const double*>(hT),
                                                              reinterpret_cast<double*>(pp[0]), make_inject<double>(p, &ih));
    PERCNN_CUDA(cudaGetLastError());
    p->launches++;
  }
  // small grids: the whole backward rollout as ONE persistent cooperative launch
  if (nsteps >= 2 && nsteps <= kMaxMultiBwdSteps && multi_step_eligible(p) && p->multi_bwd_grid > 0) {
    rc = p->elt == 4 ? launch_multi_bwd<float>(p, tape, g_tape, gmask, G, pp[0], pp[1], g_h0, dl, dl_n, nsteps, w, st)
                     : launch_multi_bwd<double>(p, tape, g_tape, gmask, G, pp[0], pp[1], g_h0, dl, dl_n, nsteps, w, st);
    if (rc) return rc;
    return percnn_param_grads_finish(p, params, param_grads, ws, stream);
  }
  int flip = (G == pp[0]) ? 1 : 0;
  for (int s = nsteps - 1; s >= 0; --s) {
    const char* add = nullptr;
    if (g_tape && gmask[s]) {
      --slot;
      add = gt + size_t(slot) * sb;
    }
    inject_for(s, &ih);
    char* gin = (s == 0) ? static_cast<char*>(g_h0) : pp[flip];
    if (s != 0) flip ^= 1;
    rc = step_bwd_any(p, static_cast<const char*>(tape) + size_t(s) * sb, G, add, gin, ws, st, nullptr, &ih);
    if (rc) return rc;
    G = gin;
  }
  return percnn_param_grads_finish(p, params, param_grads, ws, stream);
}

// ---- fused physics-residual loss (kernels_phys_loss.cuh) ------------------------------------------------
extern "C++" {
namespace {
int phys_check(const percnn_phys_loss_t* pl, Geom* g) {
  if (!pl) return fail(PERCNN_ERR_INVALID, "null physics-loss descriptor");
  if (pl->ndim != 2 && pl->ndim != 3) return fail(PERCNN_ERR_INVALID, "ndim must be 2 or 3");
  if (pl->dtype != PERCNN_F32 && pl->dtype != PERCNN_F64) return fail(PERCNN_ERR_INVALID, "dtype must be f32 or f64");
  if (pl->extent[0] < 1 || pl->extent[1] < 1 || pl->extent[2] < 1 || pl->extent[0] > (1 << 20) ||
      pl->extent[1] > (1 << 20) || pl->extent[2] > (1 << 20))
    return fail(PERCNN_ERR_INVALID, "bad extents");
  if (pl->ndim == 2 && pl->extent[0] != 1) return fail(PERCNN_ERR_INVALID, "2-D needs extent[0] == 1");
  if (pl->nframes < 3) return fail(PERCNN_ERR_INVALID, "the physics loss needs at least 3 frames (FWD:318-319: output[0:-2])");
  if (!(pl->dt > 0) || !(pl->dx > 0)) return fail(PERCNN_ERR_INVALID, "dt and dx must be positive");
  if (!percnn_device_ok(pl->device)) return fail(PERCNN_ERR_NO_DEVICE, "no sm_100 CUDA device with this ordinal");
  g->ndim = pl->ndim;
  g->D = int(pl->extent[0]);
  g->H = int(pl->extent[1]);
  g->W = int(pl->extent[2]);
  g->ghost = 0;
  g->plane = pl->ndim == 3 ? int64_t(g->H) * g->W : g->W;
  g->field = int64_t(g->D) * g->H * g->W;
  return PERCNN_OK;
}
template <typename T>
PhysLossDev<T> phys_dev(const percnn_phys_loss_t* pl, const Geom& g) {
  PhysLossDev<T> d;
  for (int q = 0; q < 2; ++q) {
    d.diff[q] = T(pl->diff[q]);
    for (int i = 0; i < 10; ++i) d.poly[q][i] = T(pl->poly[q][i]);
    const double* c = pl->poly[q];   // c00 c10 c01 c20 c11 c02 c30 c21 c12 c03
    const double du[6] = {c[1], 2 * c[3], c[4], 3 * c[6], 2 * c[7], c[8]};   // over 1 u v u^2 uv v^2
    const double dv[6] = {c[2], c[4], 2 * c[5], c[7], 2 * c[8], 3 * c[9]};
    for (int i = 0; i < 6; ++i) {
      d.dpoly[2 * q + 0][i] = T(du[i]);
      d.dpoly[2 * q + 1][i] = T(dv[i]);
    }
  }
  d.inv_dt = T(1.0 / pl->dt);
  const double t1[5] = {-1.0 / 12.0, 4.0 / 3.0, -5.0 / 2.0, 4.0 / 3.0, -1.0 / 12.0};   // FWD:18-22 along one axis
  for (int i = 0; i < 5; ++i) d.tap[i] = T(t1[i] / (pl->dx * pl->dx));
  double n = double(pl->nframes - 2);
  n *= double(g.W + 1) * double(g.H + 1) * (g.ndim == 3 ? double(g.D + 1) : 1.0);
  d.two_over_n = 2.0 / n;
  d.nframes = pl->nframes;
  d.stride = 2 * g.field;
  return d;
}
int phys_grid(const Geom& g, int frames) {
  int64_t blocks = (int64_t(g.D) * g.H * g.W * frames + kGenericThreads - 1) / kGenericThreads;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  return int(blocks < 1 ? 1 : blocks);
}
}  // namespace
}  // extern "C++"

size_t percnn_phys_loss_workspace_bytes(void) { return kWsPartials + size_t(kMaxBlocks) * sizeof(double); }

int percnn_phys_loss_fwd(const percnn_phys_loss_t* pl, const void* frames, void* resid, void* loss_out, void* ws,
                         void* stream) {
  Geom g;
  int rc = phys_check(pl, &g);
  if (rc) return rc;
  if (!frames || !loss_out || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(pl->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(ws);
  double* acc = reinterpret_cast<double*>(w + kWsAcc);
  unsigned* counter = reinterpret_cast<unsigned*>(w + kWsCounter);
  double* partials = reinterpret_cast<double*>(w + kWsPartials);
  PERCNN_CUDA(cudaMemsetAsync(w, 0, kWsPartials, st));
  const int grid = phys_grid(g, pl->nframes - 2);
  if (pl->dtype == PERCNN_F32) {
    const PhysLossDev<float> d = phys_dev<float>(pl, g);
    if (g.ndim == 3)
      k_phys_resid<float, 3><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const float*>(frames), static_cast<float*>(resid), partials, counter, acc);
    else
      k_phys_resid<float, 2><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const float*>(frames), static_cast<float*>(resid), partials, counter, acc);
    PERCNN_CUDA(cudaGetLastError());
    k_data_loss_finish<float><<<1, 32, 0, st>>>(acc, 0.5 * d.two_over_n, static_cast<float*>(loss_out));
  } else {
    const PhysLossDev<double> d = phys_dev<double>(pl, g);
    if (g.ndim == 3)
      k_phys_resid<double, 3><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const double*>(frames), static_cast<double*>(resid), partials, counter, acc);
    else
      k_phys_resid<double, 2><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const double*>(frames), static_cast<double*>(resid), partials, counter, acc);
    PERCNN_CUDA(cudaGetLastError());
    k_data_loss_finish<double><<<1, 32, 0, st>>>(acc, 0.5 * d.two_over_n, static_cast<double*>(loss_out));
  }
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

int percnn_phys_loss_bwd(const percnn_phys_loss_t* pl, const void* frames, const void* resid, const void* gscale,
                         void* g_frames, void* stream) {
  Geom g;
  int rc = phys_check(pl, &g);
  if (rc) return rc;
  if (!frames || !resid || !g_frames) return fail(PERCNN_ERR_INVALID, "null argument");
  DeviceGuard guard(pl->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = phys_grid(g, pl->nframes);
  if (pl->dtype == PERCNN_F32) {
    const PhysLossDev<float> d = phys_dev<float>(pl, g);
    if (g.ndim == 3)
      k_phys_grad<float, 3><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const float*>(frames), static_cast<const float*>(resid), static_cast<const float*>(gscale), static_cast<float*>(g_frames));
    else
      k_phys_grad<float, 2><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const float*>(frames), static_cast<const float*>(resid), static_cast<const float*>(gscale), static_cast<float*>(g_frames));
  } else {
    const PhysLossDev<double> d = phys_dev<double>(pl, g);
    if (g.ndim == 3)
      k_phys_grad<double, 3><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const double*>(frames), static_cast<const double*>(resid), static_cast<const double*>(gscale), static_cast<double*>(g_frames));
    else
      k_phys_grad<double, 2><<<grid, kGenericThreads, 0, st>>>(g, d, static_cast<const double*>(frames), static_cast<const double*>(resid), static_cast<const double*>(gscale), static_cast<double*>(g_frames));
  }
  PERCNN_CUDA(cudaGetLastError());
  return PERCNN_OK;
}

int percnn_rollout_fwd_host(percnn_plan_t* p, const void* params_host, const void* h0_host, void* traj_host,
                            const uint8_t* emit, int nsteps, void* h_final_host) {
  if (!p || !params_host || !h0_host) return fail(PERCNN_ERR_INVALID, "null argument");
  if (traj_host && !emit) return fail(PERCNN_ERR_INVALID, "traj given without an emit mask");
  DeviceGuard guard(p->desc.device);
  const size_t sb = (state_bytes(p) + 255) / 256 * 256;
  int nemit = 0;
  if (traj_host)
    for (int s = 0; s < nsteps; ++s) nemit += emit[s] ? 1 : 0;
  const size_t need = ws_states_off(p) + sb * size_t(4 + nemit);
  if (!p->h_stream) PERCNN_CUDA(cudaStreamCreateWithFlags(&p->h_stream, cudaStreamNonBlocking));
  if (!p->h_params_dev) PERCNN_CUDA(cudaMalloc(&p->h_params_dev, size_t(p->nparams) * p->elt));
  if (p->h_states_bytes < need) {
    if (p->h_states) cudaFree(p->h_states);
    p->h_states = nullptr;
    p->h_states_bytes = 0;
    PERCNN_CUDA(cudaMalloc(&p->h_states, need));
    p->h_states_bytes = need;
  }
  cudaStream_t st = p->h_stream;
  char* base = static_cast<char*>(p->h_states);
  char* ws = base;                          // header + 2 ping-pong states
  char* d_h0 = base + ws_states_off(p) + 2 * sb;
  char* d_final = d_h0 + sb;
  char* d_traj = d_final + sb;
  const size_t raw_sb = state_bytes(p);
  PERCNN_CUDA(cudaMemcpyAsync(p->h_params_dev, params_host, size_t(p->nparams) * p->elt, cudaMemcpyHostToDevice, st));
  PERCNN_CUDA(cudaMemcpyAsync(d_h0, h0_host, raw_sb, cudaMemcpyHostToDevice, st));
  int rc = percnn_params_load(p, p->h_params_dev, st);
  if (rc) return rc;
  // the device trajectory uses padded slots; rollout_fwd packs slots state_bytes apart, so run it on a
  // tightly packed view when padding is zero (always the case for sizes that are multiples of 64 cells)
  if (nemit > 0 && raw_sb != sb) {
    // fall back to emitting one frame at a time through h_final
    return fail(PERCNN_ERR_UNSUPPORTED, "host rollout with emitted frames needs state bytes to be a multiple of 256");
  }
  rc = percnn_rollout_fwd(p, d_h0, nemit ? d_traj : nullptr, emit, nsteps, h_final_host ? d_final : nullptr, nullptr, ws, st);
  if (rc) return rc;
  if (nemit) PERCNN_CUDA(cudaMemcpyAsync(traj_host, d_traj, raw_sb * size_t(nemit), cudaMemcpyDeviceToHost, st));
  if (h_final_host) PERCNN_CUDA(cudaMemcpyAsync(h_final_host, d_final, raw_sb, cudaMemcpyDeviceToHost, st));
  PERCNN_CUDA(cudaStreamSynchronize(st));
  return PERCNN_OK;
}

}  // extern "C"
