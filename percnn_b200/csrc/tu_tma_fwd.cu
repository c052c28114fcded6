// Translation unit: forward step of the 3-D Pi-block cell (k = 1, fp32) -- the TMA z-marching kernel and its
// slab-mode variant with the fused halo exchange -- plus the host helpers shared with the adjoint TU
// (tiling choice, tensor-map cache).
#include <cstdlib>

#include "kernels_gs3d_slab.cuh"
#include "plan.h"

namespace percnn {

// Work decomposition of the persistent TMA kernels.  A tile is 128 x ty cells, an item is a tile marched over
// tz planes (+4 halo planes).  Measured on B200 (profiles/r01_sweep_tma.txt): it pays to keep every CTA on the
// same planes at the same time (one item per CTA, all items in one round) -- 128 CTAs in lock-step beat 148
// CTAs on staggered z-chunks by 23 % -- so the cost model charges extra for multi-round schedules.
static double tiling_cost(int nxt, int H, int depth, int nsm, int ty, int nzc, TmaTiling* out) {
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

// min_chunk: smallest admissible z-chunk (first AND last).  The fused slab kernels need both boundary pairs
// (planes 0,1 and D-2,D-1) inside one item each, i.e. min_chunk = 2.
TmaTiling choose_tiling(int nxt, int H, int depth, int nsm, int fixed_ty, int max_ty, int min_chunk) {
  TmaTiling best{max_ty < H ? max_ty : H, depth, 0, 1};
  best.nyt = (H + best.ty - 1) / best.ty;
  double best_cost = 1e300;
  for (int ty = (fixed_ty ? fixed_ty : 1); ty <= (fixed_ty ? fixed_ty : max_ty); ++ty) {
    if (ty > H) break;
    for (int nzc = 1; nzc <= depth; ++nzc) {
      TmaTiling t;
      const double c = tiling_cost(nxt, H, depth, nsm, ty, nzc, &t);
      const int last = depth - (t.nzc - 1) * t.tz;
      if (t.tz >= min_chunk && last >= min_chunk && c < best_cost) {
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

int get_pair_map(percnn_plan* p, const void* base, CUtensorMap* out) {
  for (auto& m : p->pair_maps)
    if (m.base == base) {
      *out = m.map;
      return PERCNN_OK;
    }
  TmaPairMap& m = p->pair_maps[p->pair_rr];
  p->pair_rr = (p->pair_rr + 1) % 8;
  const Geom& g = p->g;
  cuuint64_t gdim[4] = {cuuint64_t(g.W), cuuint64_t(g.H), cuuint64_t(g.D + 2 * g.ghost), 2};
  cuuint64_t gstr[3] = {cuuint64_t(g.W) * 4, cuuint64_t(g.plane) * 4, cuuint64_t(g.field) * 4};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  cuuint32_t box[4] = {tma3d::TX, cuuint32_t(p->ty), 2, 1};
  const CUresult r = p->encode(&m.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    m.base = nullptr;
    return fail(PERCNN_ERR_CUDA, "cuTensorMapEncodeTiled (pair map) failed with CUresult " + std::to_string(int(r)));
  }
  m.base = base;
  *out = m.map;
  return PERCNN_OK;
}

// The halo helper's tensor maps for one fused step (see tma3d::SlabMaps).
int slab_fill_maps(percnn_plan* p, const SlabLink* link, const float* src, float* dst, bool down, tma3d::SlabMaps* sm) {
  (void)link;
  (void)down;
  int rc = get_pair_map(p, dst, &sm->dst);
  if (!rc) rc = get_pair_map(p, src, &sm->src);
  return rc;
}

cudaError_t tma_fwd_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_prep, d_prep, sizeof(PrepBlock), size_t(slot) * sizeof(PrepBlock),
                                 cudaMemcpyDeviceToDevice, st);
}

// Tiling (shared by the forward and adjoint kernels of the plan) and function attributes.
int tma_fwd_setup(percnn_plan* p) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return fail(PERCNN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  p->encode = reinterpret_cast<percnn_encode_tiled_fn>(fn);
  cudaError_t ae = cudaSuccess;
  switch (p->slot) {
#define PERCNN_TMA_ATTR(S)                                                                                              \
  case S:                                                                                                               \
    ae = cudaFuncSetAttribute(tma3d::k_gs3d_fwd_tma<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES); \
    if (ae == cudaSuccess)                                                                                              \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_fwd_slab<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_SLAB); \
    if (ae == cudaSuccess)                                                                                              \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_fwd_slab<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_SLAB); \
    break;
    PERCNN_TMA_ATTR(0) PERCNN_TMA_ATTR(1) PERCNN_TMA_ATTR(2) PERCNN_TMA_ATTR(3) PERCNN_TMA_ATTR(4) PERCNN_TMA_ATTR(5)
#undef PERCNN_TMA_ATTR
    default: return fail(PERCNN_ERR_INVALID, "bad parameter slot");
  }
  if (ae != cudaSuccess) return fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(tma fwd) failed");
  int fixed_ty = 0;
  if (const char* e = getenv("PERCNN_TMA_TY")) fixed_ty = atoi(e);
  // the adjoint kernel runs 15 consumer warps (tma3d::BWD_WARPS) and shares the tiling; slab plans keep one more
  // warp free for the halo helper (tma3d::SLAB_MAX_TY)
  const int max_ty = p->desc.slab_ghost ? tma3d::SLAB_MAX_TY : 15;
  if (fixed_ty < 0 || fixed_ty > max_ty || fixed_ty > p->g.H) fixed_ty = 0;
  // slab plans: both boundary pairs must sit inside one z-chunk each (fused halo kernels)
  const TmaTiling til = choose_tiling(p->g.W / tma3d::TX, p->g.H, p->g.D, p->sm_count, fixed_ty, max_ty,
                                      p->desc.slab_ghost ? 2 : 1);
  p->ty = til.ty;
  p->tz = til.tz;
  if (const char* e = getenv("PERCNN_NO_PDL")) p->pdl = atoi(e) == 0;
  if (const char* e = getenv("PERCNN_TMA_TZ")) p->tz_override = atoi(e);
  if (const char* e = getenv("PERCNN_TMA_GRID")) p->grid_override = atoi(e);
  if (const char* e = getenv("PERCNN_FLAG_SPINS")) {
    const long v = atol(e);
    if (v > 0) p->flag_spin_limit = uint32_t(v);
  }
  return PERCNN_OK;
}

// Fills the geometry part of the kernel parameters and the z-schedule; returns the grid size.
int tma_fill_params(percnn_plan* p, tma3d::Params& prm, const float* src, float* dst, int z_lo, int z_hi,
                    const SlabLink* link) {
  const Geom& g = p->g;
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
  const int depth = z_hi - z_lo;
  TmaTiling til = (z_lo == 0 && z_hi == g.D) ? TmaTiling{p->ty, p->tz, prm.nyt, (g.D + p->tz - 1) / p->tz}
                                             : choose_tiling(prm.nxt, g.H, depth, p->sm_count, p->ty, p->ty);
  if (p->tz_override > 0 && !link) {
    til.tz = p->tz_override < depth ? p->tz_override : depth;
    til.nzc = (depth + til.tz - 1) / til.tz;
  }
  prm.nseg = 1;
  prm.seg_lo[0] = z_lo;
  prm.seg_hi[0] = z_hi;
  prm.seg_tz[0] = til.tz;
  prm.seg_nzc[0] = til.nzc;
  if (link) {
    prm.fused = 1;
    prm.peer_lo_dst = link->peer_lo_dst;
    prm.peer_hi_dst = link->peer_hi_dst;
    prm.my_flags = link->my_flags;
    prm.post_lo_flag = link->post_lo_flag;
    prm.post_hi_flag = link->post_hi_flag;
    prm.scratch = link->scratch;
    prm.epoch_wait = link->epoch_wait;
    prm.epoch_post = link->epoch_post;
    prm.spin_limit = p->flag_spin_limit;
    prm.peer_lo_src = link->peer_lo_src;
    prm.peer_hi_src = link->peer_hi_src;
    prm.flush_prev = link->flush_prev ? 1 : 0;
    prm.defer_late = link->defer_late ? 1 : 0;
    if (const char* e = getenv("PERCNN_FUSED_DEBUG")) prm.debug = atoi(e);
  }
  prm.slot = p->slot;
  const int nitems = prm.nxt * prm.nyt * prm.seg_nzc[0];
  int grid = nitems < p->sm_count ? nitems : p->sm_count;
  if (p->grid_override > 0 && p->grid_override < grid) grid = p->grid_override;
  return grid;
}

int tma_fwd_launch(percnn_plan* p, const float* src, float* dst, int z_lo, int z_hi, cudaStream_t st,
                   const SlabLink* link) {
  const CUtensorMap *mm, *hm;
  int rc = get_maps(p, src, &mm, &hm);
  if (rc) return rc;
  tma3d::Params prm;
  const int grid = tma_fill_params(p, prm, src, dst, z_lo, z_hi, link);
  bool down = link && (link->epoch_wait & 1u);   // the march direction alternates from step to step
  tma3d::SlabMaps sm;
  if (link) {
    rc = slab_fill_maps(p, link, src, dst, down, &sm);
    if (rc) return rc;
  }
  cudaError_t le = cudaSuccess;
  switch (p->slot) {
#define PERCNN_TMA_CASE(S)                                                                                              \
  case S:                                                                                                               \
    if (!link) le = launch_pdl(tma3d::k_gs3d_fwd_tma<S>, grid, tma3d::THREADS, tma3d::SMEM_BYTES, st, p->pdl, *mm, *hm, prm); \
    else if (down) le = launch_pdl(tma3d::k_gs3d_fwd_slab<S, true>, grid, tma3d::THREADS, tma3d::SMEM_BYTES_SLAB, st, p->pdl, *mm, *hm, prm, sm); \
    else le = launch_pdl(tma3d::k_gs3d_fwd_slab<S, false>, grid, tma3d::THREADS, tma3d::SMEM_BYTES_SLAB, st, p->pdl, *mm, *hm, prm, sm); \
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

}  // namespace percnn
