// Translation unit: 2-D cells on shared-memory tiles with temporal blocking (kernels_tile2d.cuh).
#include <cstdio>
#include <cstdlib>

#include "kernels_tile2d.cuh"
#include "plan.h"

namespace percnn {

cudaError_t tile2d_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_prep, d_prep, sizeof(PrepBlock), size_t(slot) * sizeof(PrepBlock),
                                 cudaMemcpyDeviceToDevice, st);
}

namespace {
template <typename T>
const void* kernel_for(const percnn_plan* p) {
  const int cell = p->desc.cell;
  if (cell == PERCNN_CELL_PI) return (const void*)tile2d::k_tile2d<T, 0>;
  if (cell == PERCNN_CELL_BURGERS) return (const void*)tile2d::k_tile2d<T, 1>;
  return (const void*)tile2d::k_tile2d<T, 2>;
}
const void* kernel_of(const percnn_plan* p) { return p->elt == 4 ? kernel_for<float>(p) : kernel_for<double>(p); }
int roundup(int a, int b) { return (a + b - 1) / b * b; }
}  // namespace

bool tile2d_eligible(const percnn_plan* p) {
  const percnn_desc_t& d = p->desc;
  if (d.ndim != 2 || d.slab_ghost || getenv("PERCNN_NO_TILE2D")) return false;
  if (d.cell == PERCNN_CELL_PI && (d.ksize != 1 || (d.flags & PERCNN_FLAG_EVAL_BRANCH))) return false;
  const int vw = 16 / p->elt;
  return p->g.W % vw == 0 && p->g.W >= 4 * vw && p->g.H >= 8;
}

// Tile shape and steps per pass.  Model (microseconds per time step on one SM): the region of a pass is advanced
// at ~c_cell per cell and every pass pays a fixed cost F (load latency + grid barrier), amortised over K steps.
int tile2d_setup(percnn_plan* p) {
  const int vw = 16 / p->elt, H = p->g.H, W = p->g.W, nsm = p->sm_count;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->desc.device);
  if (!coop) return PERCNN_OK;
  const double c_cell = p->elt == 4 ? 0.00035 : 0.0009, F = 2.5;
  const size_t smem_cap = 200 * 1024;
  double best = 1e300;
  int fixedK = 0, fixedTH = 0, fixedTW = 0;   // experiment knobs
  if (const char* e = getenv("PERCNN_TILE2D_K")) fixedK = atoi(e);
  if (const char* e = getenv("PERCNN_TILE2D_TH")) fixedTH = atoi(e);
  if (const char* e = getenv("PERCNN_TILE2D_TW")) fixedTW = atoi(e);
  for (int K = (fixedK > 0 ? fixedK : 2); K <= (fixedK > 0 ? fixedK : 8); ++K) {
    const int hy = 2 * K, hx = roundup(vw + 2 * (K - 1), vw);
    for (int nty = 1; nty <= nsm && nty <= H / 4; ++nty) {
      const int TH = (H + nty - 1) / nty;
      if (fixedTH > 0 && TH != fixedTH) continue;
      for (int ntx = 1; ntx * nty <= nsm; ++ntx) {
        int TW = roundup((W + ntx - 1) / ntx, vw);
        if (TW > W) TW = W;
        if (TW < 4 * vw) break;
        if (fixedTW > 0 && TW != fixedTW) continue;
        const int nx = (W + TW - 1) / TW, ny = (H + TH - 1) / TH;
        if (nx * ny > nsm) continue;
        const int RH = TH + 2 * hy, RW = TW + 2 * hx;
        const size_t smem = size_t(4) * RH * RW * p->elt;
        if (smem > smem_cap) continue;
        double cells = 0;   // cells computed per pass: rows shrink by 2 per side per step, columns do not
        for (int j = 1; j <= K; ++j) cells += double(RH - 4 * j) * (RW - 2 * vw);
        const double t = (cells * c_cell + double(RH) * RW * 0.0002 + F) / K;
        if (t < best) {
          best = t;
          p->t2_K = K;
          p->t2_TH = TH;
          p->t2_TW = TW;
          p->t2_hy = hy;
          p->t2_hx = hx;
          p->t2_nty = ny;
          p->t2_ntx = nx;
          p->t2_smem = int(smem);
        }
      }
    }
  }
  if (best >= 1e300) return PERCNN_OK;   // no admissible tiling: the gather kernels serve this plan
  if (getenv("PERCNN_TILE2D_VERBOSE"))
    fprintf(stderr, "tile2d: %dx%d elt %d -> K=%d tile %dx%d (%dx%d tiles) halo %d/%d smem %d model %.2f us/step\n", H, W, p->elt,
            p->t2_K, p->t2_TH, p->t2_TW, p->t2_nty, p->t2_ntx, p->t2_hy, p->t2_hx, p->t2_smem, best);
  // (the attribute belongs to the FUNCTION, which every plan of this cell/dtype shares: always ask for the cap, or a
  // later plan with a smaller region would take the shared memory away from an earlier one)
  if (cudaFuncSetAttribute(kernel_of(p), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)) != cudaSuccess) {
    cudaGetLastError();
    return PERCNN_OK;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel_of(p), tile2d::THREADS, p->t2_smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return PERCNN_OK;
  }
  if (!p->d_sync && cudaMalloc(&p->d_sync, 256) != cudaSuccess) return fail(PERCNN_ERR_CUDA, "cudaMalloc(sync) failed");
  p->use_tile2d = true;
  return PERCNN_OK;
}

template <typename T>
static int rollout_t(percnn_plan* p, const void* h0, void* final_state, void* tape, void* traj, const uint8_t* emit, void* ping,
                     void* pong, int nsteps, cudaStream_t st) {
  static_assert(sizeof(tile2d::Args<T>) < 3900, "kernel parameter space");
  tile2d::Args<T> a;
  memset(&a, 0, sizeof(a));
  a.h0 = static_cast<const T*>(h0);
  a.tape = static_cast<T*>(tape);
  a.traj = static_cast<T*>(traj);
  a.ping = static_cast<T*>(ping);
  a.pong = static_cast<T*>(pong);
  a.final_state = static_cast<T*>(final_state);
  a.nsteps = nsteps;
  a.K = p->t2_K;
  a.TH = p->t2_TH;
  a.TW = p->t2_TW;
  a.hy = p->t2_hy;
  a.hx = p->t2_hx;
  a.nty = p->t2_nty;
  a.ntx = p->t2_ntx;
  a.stride = p->state_elems;
  if (traj)
    for (int s = 0; s < nsteps; ++s)
      if (emit[s]) a.emit[s >> 5] |= 1u << (s & 31);
  PERCNN_CUDA(cudaMemsetAsync(p->d_sync, 0, sizeof(unsigned), st));
  Geom g = p->g;
  int slot = p->slot;
  unsigned* counter = p->d_sync;
  void* args[] = {&g, &slot, &a, &counter};
  int grid = p->t2_nty * p->t2_ntx;
  if (grid > p->sm_count) grid = p->sm_count;
  PERCNN_CUDA(cudaLaunchCooperativeKernel(kernel_of(p), dim3(grid), dim3(tile2d::THREADS), args, size_t(p->t2_smem), st));
  p->launches++;
  return PERCNN_OK;
}

int tile2d_rollout(percnn_plan* p, const void* h0, void* final_state, void* tape, void* traj, const uint8_t* emit, void* ping,
                   void* pong, int nsteps, cudaStream_t st) {
  if (nsteps > tile2d::kMaxSteps) return fail(PERCNN_ERR_INVALID, "tile2d rollouts hold at most 4096 steps");
  return p->elt == 4 ? rollout_t<float>(p, h0, final_state, tape, traj, emit, ping, pong, nsteps, st)
                     : rollout_t<double>(p, h0, final_state, tape, traj, emit, ping, pong, nsteps, st);
}

}  // namespace percnn
