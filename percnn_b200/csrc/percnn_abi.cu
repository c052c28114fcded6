// C-ABI of libpercnn_b200.so (see include/percnn_b200.h).  Single translation unit: the kernels share
// one __constant__ parameter block, so everything is compiled together for sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "kernels_generic.cuh"
#include "kernels_gs3d_tma.cuh"
#include "kernels_gs3d_tma_bwd.cuh"
#include "kernels_phys_loss.cuh"
#include "kernels_pi_k5.cuh"
#include "kernels_prep.cuh"

using namespace percnn;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define PERCNN_CUDA(call)                                                                                   \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess)                                                                                 \
      return fail(PERCNN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                    \
  } while (0)

std::mutex g_slot_mutex;
bool g_slot_used[kPrepSlots] = {false};

constexpr int kMaxBlocks = 2048;
constexpr size_t kWsAcc = 0;            // kRedMaxSmall doubles
constexpr size_t kWsCounter = 256;      // one unsigned
constexpr size_t kWsPartials = 512;     // kMaxBlocks * kRedMaxSmall doubles
constexpr size_t kWsStates = kWsPartials + size_t(kMaxBlocks) * kRedMaxSmall * sizeof(double);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmaMapPair {
  const void* base = nullptr;
  CUtensorMap main_map, halo_map;
};

}  // namespace

struct percnn_plan {
  percnn_desc_t desc;
  Geom g;
  PrepDesc pd;
  int slot = -1;
  int elt = 4;
  int nred = 0;
  int sm_count = 148;
  int64_t nparams = 0;
  int64_t state_elems = 0;
  int64_t launches = 0;
  bool use_tma = false;
  int ty = 16, tz = 0;
  unsigned* d_sync = nullptr;   // grid-barrier counter of the persistent multi-step kernel
  int multi_grid = 0;           // co-resident grid size of that kernel (0 = not available)
  int multi_bwd_grid = 0;       // same for the persistent adjoint kernel
  bool debug_split = false;   // PERCNN_TMA_SPLIT=1
  bool bwd_split_mono = false;   // PERCNN_BWD_SPLIT_MONO=1: monomial sums in their own streaming kernel
  bool bwd_mw = false;           // PERCNN_BWD_MW=1: warp-specialised adjoint (monomial warps + TMA ring for the stored
                                 // state).  Measured 847 us vs 800 us per 512^3 step: the two monomial warps are the
                                 // critical path (688 us without their work), see DESIGN.md 3.3 -- kept as an experiment.
  bool pdl = true;   // programmatic dependent launch between consecutive step kernels (PERCNN_NO_PDL=1 disables)
  int tz_override = 0, grid_override = 0;   // experiment knobs (PERCNN_TMA_TY / _TZ / _GRID environment variables)
  PrepBlock* d_prep = nullptr;
  float* d_k5w = nullptr;
  EncodeTiledFn encode = nullptr;
  TmaMapPair maps[4];
  int map_rr = 0;
  // host-path scratch
  void* h_params_dev = nullptr;
  void* h_states = nullptr;
  size_t h_states_bytes = 0;
  cudaStream_t h_stream = nullptr;
};

namespace {

size_t state_bytes(const percnn_plan* p) { return size_t(p->state_elems) * p->elt; }

bool is_k5(const percnn_plan* p) { return p->desc.cell == PERCNN_CELL_PI && p->desc.ksize == 5; }
int k5_blocks(const percnn_plan* p) {
  return ((p->g.W + k5::BT_X - 1) / k5::BT_X) * ((p->g.H + k5::BT_Y - 1) / k5::BT_Y);
}
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

// Work decomposition of the persistent TMA kernel.  A tile is 128 x ty cells, an item is a tile marched over
// tz planes (+4 halo planes).  Measured on B200 (profiles/r01_sweep_tma.txt): it pays to keep every CTA on the
// same planes at the same time (one item per CTA, all items in one round) -- 128 CTAs in lock-step beat 148
// CTAs on staggered z-chunks by 23 % -- so the cost model charges extra for multi-round schedules.
struct TmaTiling {
  int ty, tz, nyt, nzc;
};
double tiling_cost(int nxt, int H, int depth, int nsm, int ty, int nzc, TmaTiling* out) {
  const int nyt = (H + ty - 1) / ty;
  const int tz = (depth + nzc - 1) / nzc;
  const int nz_chunks = (depth + tz - 1) / tz;
  const long items = long(nxt) * nyt * nz_chunks;
  const long rounds = (items + nsm - 1) / nsm;
  double cost = double(rounds) * (tz + 4) * (8.0 + 2.0 * ty + 4.0);   // fixed per-plane latency + rows in + rows out
  if (rounds > 1) cost *= 1.25;
  if (out) *out = TmaTiling{ty, tz, nyt, nz_chunks};
  return cost;
}
TmaTiling choose_tiling(int nxt, int H, int depth, int nsm, int fixed_ty, int max_ty = tma3d::BWD_WARPS) {
  TmaTiling best{max_ty, depth, (H + max_ty - 1) / max_ty, 1};
  double best_cost = 1e300;
  for (int ty = (fixed_ty ? fixed_ty : 1); ty <= (fixed_ty ? fixed_ty : max_ty); ++ty) {
    if (ty > H) break;
    for (int nzc = 1; nzc <= depth; ++nzc) {
      TmaTiling t;
      const double c = tiling_cost(nxt, H, depth, nsm, ty, nzc, &t);
      if (c < best_cost) {
        best_cost = c;
        best = t;
      }
      if ((depth + nzc - 1) / nzc <= 2) break;
    }
  }
  return best;
}

int get_maps(percnn_plan* p, const void* src, const CUtensorMap** main_map, const CUtensorMap** halo_map) {
  for (auto& m : p->maps)
    if (m.base == src) {
      *main_map = &m.main_map;
      *halo_map = &m.halo_map;
      return PERCNN_OK;
    }
  TmaMapPair& m = p->maps[p->map_rr];
  p->map_rr = (p->map_rr + 1) % 4;
  const Geom& g = p->g;
  const cuuint64_t planes = cuuint64_t(g.D + 2 * g.ghost);
  cuuint64_t gdim[4] = {cuuint64_t(g.W), cuuint64_t(g.H), planes, 2};
  cuuint64_t gstr[3] = {cuuint64_t(g.W) * 4, cuuint64_t(g.plane) * 4, cuuint64_t(g.field) * 4};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  cuuint32_t box_main[4] = {tma3d::TX, cuuint32_t(p->ty), 1, 1};
  cuuint32_t box_halo[4] = {tma3d::TX, 1, 1, 1};   // halo rows go one by one so any tile origin wraps correctly
  CUresult r = p->encode(&m.main_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(src), gdim, gstr, box_main,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = p->encode(&m.halo_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(src), gdim, gstr, box_halo, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    m.base = nullptr;
    return fail(PERCNN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  }
  m.base = src;
  *main_map = &m.main_map;
  *halo_map = &m.halo_map;
  return PERCNN_OK;
}

// Launch with programmatic stream serialization allowed (the kernels call griddepcontrol.wait themselves).
template <typename... KArgs>
cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, KArgs... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

struct SlabLink {   // mirrors percnn_slab_link_t
  float* peer_lo_dst = nullptr;
  float* peer_hi_dst = nullptr;
  const uint32_t* my_flags = nullptr;
  uint32_t* post_lo_flag = nullptr;
  uint32_t* post_hi_flag = nullptr;
  uint32_t* scratch = nullptr;
  uint32_t epoch_wait = 0, epoch_post = 0;
};

int launch_tma_fwd(percnn_plan* p, const float* src, float* dst, int z_lo, int z_hi, cudaStream_t st,
                   const SlabLink* link = nullptr, const tma3d::BwdExtra* bwd = nullptr) {
  const CUtensorMap *mm, *hm;
  int rc = get_maps(p, src, &mm, &hm);
  if (rc) return rc;
  CUtensorMap h_map;   // warp-specialised adjoint: the stored state streams through TMA as well (copied: the cache may recycle the slot)
  memset(&h_map, 0, sizeof(h_map));
  if (bwd && p->bwd_mw && !p->bwd_split_mono) {
    const CUtensorMap main_copy = *mm, halo_copy = *hm;   // get_maps may evict these when it encodes the map for h
    const CUtensorMap *h_main, *h_halo;
    rc = get_maps(p, bwd->h, &h_main, &h_halo);
    if (rc) return rc;
    h_map = *h_main;
    static thread_local CUtensorMap keep_main, keep_halo;
    keep_main = main_copy;
    keep_halo = halo_copy;
    mm = &keep_main;
    hm = &keep_halo;
  }
  const Geom& g = p->g;
  tma3d::Params prm;
  memset(&prm, 0, sizeof(prm));
  prm.src = src;
  prm.dst = dst;
  prm.D = g.D;
  prm.H = g.H;
  prm.W = g.W;
  prm.src_planes = g.D + 2 * g.ghost;
  prm.src_field = g.field;
  prm.dst_field = g.field;
  prm.src_zoff = g.ghost ? 0 : -2;
  prm.dst_zoff = g.ghost;
  prm.wrap_z = g.ghost ? 0 : 1;
  prm.nxt = g.W / tma3d::TX;
  prm.ty = p->ty;
  prm.nyt = (g.H + p->ty - 1) / p->ty;
  auto add_segment = [&](int lo, int hi) {
    const int depth = hi - lo;
    TmaTiling til = (lo == 0 && hi == g.D) ? TmaTiling{p->ty, p->tz, prm.nyt, (g.D + p->tz - 1) / p->tz}
                                           : choose_tiling(prm.nxt, g.H, depth, p->sm_count, p->ty);
    if (p->tz_override > 0) {
      til.tz = p->tz_override < depth ? p->tz_override : depth;
      til.nzc = (depth + til.tz - 1) / til.tz;
    }
    const int s = prm.nseg++;
    prm.seg_lo[s] = lo;
    prm.seg_hi[s] = hi;
    prm.seg_tz[s] = til.tz;
    prm.seg_nzc[s] = til.nzc;
  };
  if (link) {
    add_segment(0, 2);
    add_segment(g.D - 2, g.D);
    add_segment(2, g.D - 2);
    prm.fused = 1;
    prm.peer_lo_dst = link->peer_lo_dst;
    prm.peer_hi_dst = link->peer_hi_dst;
    prm.my_flags = link->my_flags;
    prm.post_lo_flag = link->post_lo_flag;
    prm.post_hi_flag = link->post_hi_flag;
    prm.scratch = link->scratch;
    prm.epoch_wait = link->epoch_wait;
    prm.epoch_post = link->epoch_post;
    if (const char* e = getenv("PERCNN_FUSED_DEBUG")) prm.debug = atoi(e);
  } else if (getenv("PERCNN_KERNEL_DEBUG")) {
    prm.debug = atoi(getenv("PERCNN_KERNEL_DEBUG"));
    add_segment(z_lo, z_hi);
  } else if (p->debug_split && z_lo == 0 && z_hi == g.D && g.D >= 5) {
    add_segment(0, 2);            // debugging aid: the fused step's three-segment schedule without any flags
    add_segment(g.D - 2, g.D);
    add_segment(2, g.D - 2);
  } else {
    add_segment(z_lo, z_hi);
  }
  prm.slot = p->slot;
  int nitems = 0;
  for (int s = 0; s < prm.nseg; ++s) nitems += prm.nxt * prm.nyt * prm.seg_nzc[s];
  int grid = nitems < p->sm_count ? nitems : p->sm_count;
  if (p->grid_override > 0 && p->grid_override < grid) grid = p->grid_override;
  cudaError_t le = cudaSuccess;
  switch (p->slot) {
#define PERCNN_TMA_CASE(S) \
  case S:                                                                                                   \
    if (bwd && p->bwd_mw && !p->bwd_split_mono) {                                                           \
      if (prm.fused) le = launch_pdl(tma3d::k_gs3d_bwd_tma_mw<S, true>, grid, tma3d::MW_THREADS, tma3d::SMEM_BYTES_BWD_MW, st, p->pdl, *mm, *hm, h_map, prm, *bwd); \
      else le = launch_pdl(tma3d::k_gs3d_bwd_tma_mw<S, false>, grid, tma3d::MW_THREADS, tma3d::SMEM_BYTES_BWD_MW, st, p->pdl, *mm, *hm, h_map, prm, *bwd); \
    } else if (bwd) {                                                                                       \
      if (p->bwd_split_mono) {                                                                              \
        if (prm.fused) le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, true, false>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD, st, p->pdl, *mm, *hm, prm, *bwd); \
        else le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, false, false>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD, st, p->pdl, *mm, *hm, prm, *bwd); \
      } else {                                                                                              \
        if (prm.fused) le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, true, true>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD, st, p->pdl, *mm, *hm, prm, *bwd); \
        else le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, false, true>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD, st, p->pdl, *mm, *hm, prm, *bwd); \
      }                                                                                                     \
    } else if (prm.fused) le = launch_pdl(tma3d::k_gs3d_fwd_tma<S, true>, grid, tma3d::THREADS, tma3d::SMEM_BYTES, st, p->pdl, *mm, *hm, prm); \
    else le = launch_pdl(tma3d::k_gs3d_fwd_tma<S, false>, grid, tma3d::THREADS, tma3d::SMEM_BYTES, st, p->pdl, *mm, *hm, prm); \
    break;
    PERCNN_TMA_CASE(0) PERCNN_TMA_CASE(1) PERCNN_TMA_CASE(2) PERCNN_TMA_CASE(3) PERCNN_TMA_CASE(4) PERCNN_TMA_CASE(5)
#undef PERCNN_TMA_CASE
    default: return fail(PERCNN_ERR_INVALID, "bad parameter slot");
  }
  if (le != cudaSuccess) return fail(PERCNN_ERR_CUDA, std::string("TMA kernel launch: ") + cudaGetErrorString(le));
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
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
bool multi_step_eligible(const percnn_plan* p) {
  if (p->use_tma || is_k5(p) || p->desc.slab_ghost) return false;
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

// Type-erased description of the fused data-loss gradient of one state (see percnn_data_loss_t).
struct InjectHost {
  const void* target = nullptr;   // that state's low-res frame
  const void* gscale = nullptr;
  int stride = 1;
  int64_t n_total = 0;
};
int lowres(int n, int s) { return (n + s - 1) / s; }
int64_t lowres_field_elems(const percnn_plan* p, int s) {
  const Geom& g = p->g;
  return int64_t(g.ndim == 3 ? lowres(g.D, s) : 1) * lowres(g.H, s) * lowres(g.W, s);
}
template <typename T>
Inject<T> make_inject(const percnn_plan* p, const InjectHost* ih) {
  Inject<T> j;
  memset(&j, 0, sizeof(j));
  j.s = 1;
  if (!ih || !ih->target) return j;
  j.target = static_cast<const T*>(ih->target);
  j.gscale = static_cast<const T*>(ih->gscale);
  j.two_over_n = 2.0 / double(ih->n_total);
  j.s = ih->stride;
  j.lh = lowres(p->g.H, ih->stride);
  j.lw = lowres(p->g.W, ih->stride);
  j.lfield = lowres_field_elems(p, ih->stride);
  return j;
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
  if (p->use_tma) return launch_tma_fwd(p, static_cast<const float*>(src), static_cast<float*>(dst), 0, p->g.D, st);
  if (p->desc.cell == PERCNN_CELL_PI && p->desc.ksize == 5) {
    dim3 grid((p->g.W + k5::TILE_X - 1) / k5::TILE_X, (p->g.H + k5::TILE_Y - 1) / k5::TILE_Y);
    k5::k_pi_k5_fwd<<<grid, k5::THREADS, k5::smem_bytes(p->desc.hidden), st>>>(
        p->g, p->slot, p->desc.hidden, static_cast<const float*>(src), static_cast<float*>(dst), p->d_k5w);
    PERCNN_CUDA(cudaGetLastError());
    p->launches++;
    return PERCNN_OK;
  }
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
  if (p->use_tma) {
    char* w = static_cast<char*>(ws);
    tma3d::BwdExtra x;
    x.h = static_cast<const float*>(h);
    x.gadd = static_cast<const float*>(gadd);
    x.partials = reinterpret_cast<double*>(w + kWsPartials);
    x.counter = reinterpret_cast<unsigned*>(w + kWsCounter);
    x.acc = reinterpret_cast<double*>(w + kWsAcc);
    x.inj = make_inject<float>(p, ih);
    if (p->bwd_split_mono) {  // the 20 stencil-free monomial sums as a separate streaming pass over h and G
      const Geom& g = p->g;
      const int64_t n4 = int64_t(g.D) * g.plane / 4;
      int64_t blocks = (n4 + 255) / 256;
      if (blocks > int64_t(p->sm_count) * 8) blocks = int64_t(p->sm_count) * 8;
      tma3d::k_monomial_sums<<<int(blocks), 256, 0, st>>>(
          x.h, static_cast<const float*>(gout), g.field, int64_t(g.ghost) * g.plane, n4, float(p->desc.dt),
          reinterpret_cast<double*>(w + kWsPartials + 8192), reinterpret_cast<unsigned*>(w + kWsCounter + 64), x.acc);
      PERCNN_CUDA(cudaGetLastError());
      p->launches++;
    }
    return launch_tma_fwd(p, static_cast<const float*>(gout), static_cast<float*>(gin), 0, p->g.D, st, link, &x);
  }
  if (link) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo adjoint needs a TMA plan");
  if (is_k5(p)) {
    char* w = static_cast<char*>(ws);
    double* acc = reinterpret_cast<double*>(w);
    float* partials = reinterpret_cast<float*>(w + ws_header_bytes(p));
    dim3 grid((p->g.W + k5::BT_X - 1) / k5::BT_X, (p->g.H + k5::BT_Y - 1) / k5::BT_Y);
    const int np = int(p->nparams);
    k5::k_pi_k5_bwd<<<grid, k5::BTHREADS, k5::bwd_smem_floats(p->desc.hidden, np) * sizeof(float), st>>>(
        p->g, p->slot, p->desc.hidden, np, static_cast<const float*>(h), static_cast<const float*>(gout),
        static_cast<const float*>(gadd), static_cast<float*>(gin), p->d_k5w, partials);
    PERCNN_CUDA(cudaGetLastError());
    if (ih && ih->target) {
      // the 5x5 adjoint sits at its register limit; the (rare, 1/s^2-sized) loss injection runs as its own pass
      k_inject_only<float><<<generic_grid(p), kGenericThreads, 0, st>>>(p->g, static_cast<const float*>(h),
                                                                         static_cast<float*>(gin), make_inject<float>(p, ih));
      PERCNN_CUDA(cudaGetLastError());
      p->launches++;
    }
    k5::k5_reduce_partials<<<(np + 255) / 256, 256, 0, st>>>(partials, int(grid.x * grid.y), np, acc);
    PERCNN_CUDA(cudaGetLastError());
    p->launches += 2;
    return PERCNN_OK;
  }
  return p->elt == 4 ? step_bwd_t<float>(p, static_cast<const float*>(h), static_cast<const float*>(gout),
                                         static_cast<const float*>(gadd), static_cast<float*>(gin),
                                         static_cast<char*>(ws), st, ih)
                     : step_bwd_t<double>(p, static_cast<const double*>(h), static_cast<const double*>(gout),
                                          static_cast<const double*>(gadd), static_cast<double*>(gin),
                                          static_cast<char*>(ws), st, ih);
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
  do {
    if (cudaSetDevice(d->device) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaSetDevice failed"); break; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaGetDeviceProperties failed"); break; }
    p->sm_count = prop.multiProcessorCount;
    {
      std::lock_guard<std::mutex> lk(g_slot_mutex);
      for (int s = 0; s < kPrepSlots; ++s)
        if (!g_slot_used[s]) { g_slot_used[s] = true; p->slot = s; break; }
    }
    if (p->slot < 0) { rc = fail(PERCNN_ERR_INVALID, "too many live plans (6 parameter slots)"); break; }
    if (cudaMalloc(&p->d_prep, sizeof(PrepBlock)) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaMalloc(prep) failed"); break; }
    if (d->cell == PERCNN_CELL_PI && d->ksize == 5) {
      if (cudaMalloc(&p->d_k5w, size_t(k5_total_floats(d->hidden)) * 4) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaMalloc(k5w) failed"); break; }
      if (cudaFuncSetAttribute(k5::k_pi_k5_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               int(k5::smem_bytes(d->hidden))) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(k5) failed"); break; }
      if (cudaFuncSetAttribute(k5::k_pi_k5_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               int(k5::bwd_smem_floats(d->hidden, int(p->nparams)) * sizeof(float))) != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(k5 bwd) failed"); break; }
    }
    if (!(d->cell == PERCNN_CELL_PI && d->ksize == 5) && !d->slab_ghost && !getenv("PERCNN_NO_MULTISTEP")) {
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
    p->use_tma = d->cell == PERCNN_CELL_PI && d->ksize == 1 && d->ndim == 3 && d->dtype == PERCNN_F32 &&
                 !(d->flags & (PERCNN_FLAG_NO_TMA | PERCNN_FLAG_EVAL_BRANCH)) && g.W % tma3d::TX == 0 &&
                 g.H >= 4 && g.D >= 4;
    if (p->use_tma) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        rc = fail(PERCNN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        break;
      }
      p->encode = reinterpret_cast<EncodeTiledFn>(fn);
      cudaError_t ae = cudaSuccess;
      switch (p->slot) {
#define PERCNN_TMA_ATTR(S) \
  case S:                                                                                                              \
    ae = cudaFuncSetAttribute(tma3d::k_gs3d_fwd_tma<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_fwd_tma<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma_mw<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD_MW); \
    if (ae == cudaSuccess)                                                                                               \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma_mw<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD_MW); \
    break;
        PERCNN_TMA_ATTR(0) PERCNN_TMA_ATTR(1) PERCNN_TMA_ATTR(2) PERCNN_TMA_ATTR(3) PERCNN_TMA_ATTR(4) PERCNN_TMA_ATTR(5)
#undef PERCNN_TMA_ATTR
      }
      if (ae != cudaSuccess) { rc = fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(tma) failed"); break; }
      int fixed_ty = 0;
      if (const char* e = getenv("PERCNN_TMA_TY")) fixed_ty = atoi(e);
      if (const char* e = getenv("PERCNN_BWD_MW")) p->bwd_mw = atoi(e) != 0;
      const int max_ty = p->bwd_mw ? tma3d::MW_STENCIL : tma3d::BWD_WARPS;   // the tiling is shared by fwd and adjoint kernels
      if (fixed_ty < 0 || fixed_ty > max_ty || fixed_ty > g.H) fixed_ty = 0;
      const TmaTiling til = choose_tiling(g.W / tma3d::TX, g.H, g.D, p->sm_count, fixed_ty, max_ty);
      p->ty = til.ty;
      p->tz = til.tz;
      if (const char* e = getenv("PERCNN_NO_PDL")) p->pdl = atoi(e) == 0;
      if (const char* e = getenv("PERCNN_TMA_SPLIT")) p->debug_split = atoi(e) != 0;
      if (const char* e = getenv("PERCNN_BWD_SPLIT_MONO")) p->bwd_split_mono = atoi(e) != 0;
      if (const char* e = getenv("PERCNN_TMA_TZ")) p->tz_override = atoi(e);
      if (const char* e = getenv("PERCNN_TMA_GRID")) p->grid_override = atoi(e);
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
  if (p->d_prep) cudaFree(p->d_prep);
  if (p->d_k5w) cudaFree(p->d_k5w);
  if (p->d_sync) cudaFree(p->d_sync);
  if (p->h_params_dev) cudaFree(p->h_params_dev);
  if (p->h_states) cudaFree(p->h_states);
  if (p->h_stream) cudaStreamDestroy(p->h_stream);
  if (p->slot >= 0) {
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    g_slot_used[p->slot] = false;
  }
  delete p;
  return PERCNN_OK;
}

int64_t percnn_param_count(const percnn_plan_t* p) { return p ? p->nparams : -1; }
int64_t percnn_state_elems(const percnn_plan_t* p) { return p ? p->state_elems : -1; }
int percnn_plan_uses_tma(const percnn_plan_t* p) { return p && p->use_tma ? 1 : 0; }
int64_t percnn_plan_launch_count(const percnn_plan_t* p) { return p ? p->launches : -1; }

size_t percnn_workspace_bytes(const percnn_plan_t* p, int nsteps) {
  (void)nsteps;
  if (!p) return 0;
  return ws_states_off(p) + 2 * ((state_bytes(p) + 255) / 256 * 256);
}

int percnn_params_load(percnn_plan_t* p, const void* params, void* stream) {
  if (!p || !params) return fail(PERCNN_ERR_INVALID, "null plan or params");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->elt == 4)
    k_prep<float><<<1, 256, 0, st>>>(static_cast<const float*>(params), p->pd, p->d_prep, p->d_k5w);
  else
    k_prep<double><<<1, 256, 0, st>>>(static_cast<const double*>(params), p->pd, p->d_prep, p->d_k5w);
  PERCNN_CUDA(cudaGetLastError());
  PERCNN_CUDA(cudaMemcpyToSymbolAsync(c_prep, p->d_prep, sizeof(PrepBlock), size_t(p->slot) * sizeof(PrepBlock),
                                      cudaMemcpyDeviceToDevice, st));
  p->launches++;
  return PERCNN_OK;
}

int percnn_step_fwd(percnn_plan_t* p, const void* h_in, void* h_out, void* stream) {
  if (!p || !h_in || !h_out) return fail(PERCNN_ERR_INVALID, "null argument");
  if (h_in == h_out) return fail(PERCNN_ERR_INVALID, "step_fwd cannot run in place");
  return step_fwd_any(p, h_in, h_out, static_cast<cudaStream_t>(stream));
}

// Forward step restricted to interior planes [z_lo, z_hi) of a 3-D TMA plan (slab mode overlap: the
// planes that need ghosts are launched after the halo exchange, the rest before).  Not part of the
// reference surface; used by percnn_b200.halo.
int percnn_step_fwd_range(percnn_plan_t* p, const void* h_in, void* h_out, int z_lo, int z_hi, void* stream) {
  if (!p || !h_in || !h_out) return fail(PERCNN_ERR_INVALID, "null argument");
  if (!p->use_tma) return fail(PERCNN_ERR_UNSUPPORTED, "step_fwd_range needs a TMA plan");
  if (z_lo < 0 || z_hi > p->g.D || z_lo >= z_hi) return fail(PERCNN_ERR_INVALID, "bad plane range");
  return launch_tma_fwd(p, static_cast<const float*>(h_in), static_cast<float*>(h_out), z_lo, z_hi,
                        static_cast<cudaStream_t>(stream));
}

// One fused slab step: boundary planes first (each also stored into the neighbour's ghost planes through the
// peer mapping), flags raised from inside the kernel, interior last.  See percnn_slab_link_t.
int percnn_step_fwd_fused_halo(percnn_plan_t* p, const void* h_in, void* h_out, const percnn_slab_link_t* link, void* stream) {
  if (!p || !h_in || !h_out || !link) return fail(PERCNN_ERR_INVALID, "null argument");
  if (!p->use_tma || !p->desc.slab_ghost) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo steps need a slab-mode TMA plan");
  if (p->g.D < 5) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo steps need at least 5 planes per rank");
  if (!link->peer_lo_out || !link->peer_hi_out || !link->my_flags || !link->peer_lo_flags || !link->peer_hi_flags || !link->scratch)
    return fail(PERCNN_ERR_INVALID, "incomplete slab link");
  SlabLink l;
  l.peer_lo_dst = static_cast<float*>(link->peer_lo_out);
  l.peer_hi_dst = static_cast<float*>(link->peer_hi_out);
  l.my_flags = link->my_flags;
  l.post_lo_flag = link->peer_lo_flags + 1;   // I provide the lower neighbour's UPPER ghosts
  l.post_hi_flag = link->peer_hi_flags + 0;   // and the upper neighbour's LOWER ghosts
  l.scratch = link->scratch;
  l.epoch_wait = link->epoch;
  l.epoch_post = link->epoch + 1;
  return launch_tma_fwd(p, static_cast<const float*>(h_in), static_cast<float*>(h_out), 0, p->g.D,
                        static_cast<cudaStream_t>(stream), &l);
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
  if (!link) return step_bwd_any(p, h_in, g_out, g_add, g_in, ws, static_cast<cudaStream_t>(stream), nullptr, &ih);
  if (!p->use_tma || !p->desc.slab_ghost) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo steps need a slab-mode TMA plan");
  if (p->g.D < 5) return fail(PERCNN_ERR_UNSUPPORTED, "fused halo steps need at least 5 planes per rank");
  if (!link->peer_lo_out || !link->peer_hi_out || !link->my_flags || !link->peer_lo_flags || !link->peer_hi_flags || !link->scratch)
    return fail(PERCNN_ERR_INVALID, "incomplete slab link");
  SlabLink l;
  l.peer_lo_dst = static_cast<float*>(link->peer_lo_out);
  l.peer_hi_dst = static_cast<float*>(link->peer_hi_out);
  l.my_flags = link->my_flags;
  l.post_lo_flag = link->peer_lo_flags + 1;
  l.post_hi_flag = link->peer_hi_flags + 0;
  l.scratch = link->scratch;
  l.epoch_wait = link->epoch;
  l.epoch_post = link->epoch + 1;
  return step_bwd_any(p, h_in, g_out, g_add, g_in, ws, static_cast<cudaStream_t>(stream), &l, &ih);
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
  PERCNN_CUDA(cudaMemsetAsync(ws, 0, ws_header_bytes(p), static_cast<cudaStream_t>(stream)));
  return PERCNN_OK;
}

int percnn_step_bwd(percnn_plan_t* p, const void* h_in, const void* g_out, const void* g_add, void* g_in, void* ws,
                    void* stream) {
  if (!p || !h_in || !g_out || !g_in || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  if (g_out == g_in) return fail(PERCNN_ERR_INVALID, "step_bwd cannot run in place");
  return step_bwd_any(p, h_in, g_out, g_add, g_in, ws, static_cast<cudaStream_t>(stream));
}

int percnn_param_grads_finish(percnn_plan_t* p, const void* params, void* param_grads, void* ws, void* stream) {
  if (!p || !params || !param_grads || !ws) return fail(PERCNN_ERR_INVALID, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double* acc = reinterpret_cast<const double*>(static_cast<char*>(ws) + kWsAcc);
  if (is_k5(p)) {
    const int np = int(p->nparams);
    k5::k5_finish<<<(np + 255) / 256, 256, 0, st>>>(static_cast<const float*>(params), acc, p->pd, np,
                                                     static_cast<float*>(param_grads));
    PERCNN_CUDA(cudaGetLastError());
    p->launches++;
    return PERCNN_OK;
  }
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
  if (p->desc.slab_ghost) return fail(PERCNN_ERR_UNSUPPORTED, "slab-mode rollouts are driven step by step (halo exchange between steps)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sb = state_bytes(p);
  char* tp = static_cast<char*>(tape);
  char* pp[2] = {nullptr, nullptr};
  if (ws) {
    pp[0] = static_cast<char*>(ws) + ws_states_off(p);
    pp[1] = pp[0] + (sb + 255) / 256 * 256;
  }
  if (tp && tp != h0) PERCNN_CUDA(cudaMemcpyAsync(tp, h0, sb, cudaMemcpyDeviceToDevice, st));
  // small grids: one persistent cooperative launch for the whole rollout (tape mode, or final-state-only mode)
  if (nsteps >= 2 && multi_step_eligible(p) && !traj && (tp || (h_final && pp[0]))) {
    return p->elt == 4 ? launch_multi_step<float>(p, h0, tp, pp[0], pp[1], h_final, nsteps, st)
                       : launch_multi_step<double>(p, h0, tp, pp[0], pp[1], h_final, nsteps, st);
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
  PERCNN_CUDA(cudaSetDevice(p->desc.device));
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
