// Host-side plan object and the launcher interface between the translation units of libpercnn_b200.so.
//
// The library is built from several .cu files compiled in parallel (percnn_b200/build.py).  Each of them holds its
// own copy of the __constant__ parameter block (common.cuh: `static __constant__`), so percnn_params_load copies
// the digested block into every copy through the *_load_prep functions declared here.
#pragma once
#include <cstring>
#include <string>

#include "common.cuh"
#include "kernels_prep.cuh"

typedef CUresult (*percnn_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                           CUtensorMapFloatOOBfill);

namespace percnn {

struct TmaMapPair {
  const void* base = nullptr;
  CUtensorMap main_map, halo_map;
};

struct TmaPairMap {
  const void* base = nullptr;
  CUtensorMap map;
};

constexpr int kMaxBlocks = 2048;
constexpr size_t kWsAcc = 0;            // kRedMaxSmall doubles
constexpr size_t kWsCounter = 256;      // one unsigned
constexpr size_t kWsPartials = 512;     // kMaxBlocks * kRedMaxSmall doubles
constexpr size_t kWsStates = kWsPartials + size_t(kMaxBlocks) * kRedMaxSmall * sizeof(double);

// Type-erased description of the fused data-loss gradient of one state (see percnn_data_loss_t).
struct InjectHost {
  const void* target = nullptr;   // that state's low-res frame
  const void* gscale = nullptr;
  int stride = 1;
  int64_t n_total = 0;
};

// Resolved percnn_slab_link_t of one fused-halo step.
struct SlabLink {
  float* peer_lo_dst = nullptr;
  float* peer_hi_dst = nullptr;
  const uint32_t* my_flags = nullptr;
  uint32_t* post_lo_flag = nullptr;
  uint32_t* post_hi_flag = nullptr;
  uint32_t* scratch = nullptr;
  uint32_t epoch_wait = 0, epoch_post = 0;
  float* peer_lo_src = nullptr;   // the neighbours' mappings of the buffers that correspond to my input (FLUSH_PREV)
  float* peer_hi_src = nullptr;
  bool flush_prev = false, defer_late = false;
};

}  // namespace percnn

struct percnn_plan {
  percnn_desc_t desc;
  percnn::Geom g;
  percnn::PrepDesc pd;
  int slot = -1;
  int elt = 4;
  int nred = 0;
  int sm_count = 148;
  int64_t nparams = 0;
  int64_t state_elems = 0;
  int64_t launches = 0;
  bool use_tma = false;
  bool use_tile2d = false;      // 2-D shared-memory tiled kernels with temporal blocking (kernels_tile2d.cuh)
  int t2_K = 0, t2_TH = 0, t2_TW = 0, t2_hy = 0, t2_hx = 0, t2_nty = 0, t2_ntx = 0, t2_smem = 0;
  int ty = 16, tz = 0;
  unsigned* d_sync = nullptr;   // grid-barrier counter of the persistent multi-step kernels
  int multi_grid = 0;           // co-resident grid size of that kernel (0 = not available)
  int multi_bwd_grid = 0;       // same for the persistent adjoint kernel
  int multi_slab_grid = 0;      // slab plans: co-resident grid of the persistent small-slab rollout kernel
  bool pdl = true;   // programmatic dependent launch between consecutive step kernels (PERCNN_NO_PDL=1 disables)
  int tz_override = 0, grid_override = 0;   // experiment knobs (PERCNN_TMA_TY / _TZ / _GRID environment variables)
  uint32_t flag_spin_limit = 1u << 26;      // PERCNN_FLAG_SPINS: spins before a fused-halo wait traps (tests shorten it)
  percnn::PrepBlock* d_prep = nullptr;
  float* d_k5w = nullptr;
  percnn_encode_tiled_fn encode = nullptr;
  percnn::TmaMapPair maps[4];
  int map_rr = 0;
  percnn::TmaPairMap pair_maps[8];   // slab mode: per buffer, box = one field's boundary pair (halo helper)
  int pair_rr = 0;
  // host-path scratch
  void* h_params_dev = nullptr;
  void* h_states = nullptr;
  size_t h_states_bytes = 0;
  cudaStream_t h_stream = nullptr;
};

namespace percnn {

int fail(int code, const std::string& msg);   // sets the thread-local error message, returns `code`

// Makes a device current for the duration of an entry point and restores the caller's device.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

#define PERCNN_CUDA(call)                                                                                   \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess)                                                                                 \
      return ::percnn::fail(PERCNN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
  } while (0)

inline size_t state_bytes(const percnn_plan* p) { return size_t(p->state_elems) * p->elt; }
inline bool is_k5(const percnn_plan* p) { return p->desc.cell == PERCNN_CELL_PI && p->desc.ksize == 5; }
inline int lowres(int n, int s) { return (n + s - 1) / s; }
inline int64_t lowres_field_elems(const percnn_plan* p, int s) {
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

// ---- tu_tma_fwd.cu: 3-D k=1 forward step (TMA z-marching kernel; slab variant with the fused halo exchange) ----
cudaError_t tma_fwd_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st);
int tma_fwd_setup(percnn_plan* p);          // tiling + function attributes (plan creation)
int tma_fwd_launch(percnn_plan* p, const float* src, float* dst, int z_lo, int z_hi, cudaStream_t st,
                   const SlabLink* link);
// ---- tu_tma_bwd.cu: its adjoint ----
cudaError_t tma_bwd_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st);
int tma_bwd_setup(percnn_plan* p);
int tma_bwd_launch(percnn_plan* p, const float* h, const float* gout, const float* gadd, float* gin, char* ws,
                   cudaStream_t st, const SlabLink* link, const InjectHost* ih);
// ---- tu_k5.cu: 5x5 Pi-block cell ----
cudaError_t k5_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st);
int k5_setup(percnn_plan* p);
int k5_blocks(const percnn_plan* p);
int k5_step_fwd(percnn_plan* p, const float* src, float* dst, cudaStream_t st);
int k5_step_bwd(percnn_plan* p, const float* h, const float* gout, const float* gadd, float* gin, double* acc,
                float* partials, cudaStream_t st);
int k5_grads_finish(percnn_plan* p, const float* params, const double* acc, float* grads, cudaStream_t st);
// ---- tu_tile2d.cu: 2-D cells on shared-memory tiles, several time steps per launch ----
cudaError_t tile2d_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st);
bool tile2d_eligible(const percnn_plan* p);
int tile2d_setup(percnn_plan* p);
// Advances `nsteps` (>= 1, <= 4096) steps from h0 in ONE cooperative launch.  tape != nullptr: every state s+1 goes to
// tape + (s+1) * state_elems (tape slot 0 must hold h0).  Otherwise: emitted states (emit[s] != 0, traj != nullptr) go to
// consecutive traj slots, the last state to final_state (nullable only with a tape); ping / pong are scratch states.
int tile2d_rollout(percnn_plan* p, const void* h0, void* final_state, void* tape, void* traj, const uint8_t* emit, void* ping,
                   void* pong, int nsteps, cudaStream_t st);

// ---- shared TMA host helpers (tma_host.cpp part of tu_tma_fwd.cu) ----
struct TmaTiling {
  int ty, tz, nyt, nzc;
};
TmaTiling choose_tiling(int nxt, int H, int depth, int nsm, int fixed_ty, int max_ty, int min_chunk = 1);
int get_maps(percnn_plan* p, const void* src, const CUtensorMap** main_map, const CUtensorMap** halo_map);
int get_pair_map(percnn_plan* p, const void* base, CUtensorMap* out);   // box {128, ty, 2 planes, 1 field}

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

}  // namespace percnn
