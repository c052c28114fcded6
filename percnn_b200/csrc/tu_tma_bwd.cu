// Translation unit: adjoint of the 3-D Pi-block step (k = 1, fp32), TMA z-marching kernel, plain and slab mode.
#include "kernels_gs3d_tma_bwd.cuh"
#include "plan.h"

namespace percnn {

int tma_fill_params(percnn_plan* p, tma3d::Params& prm, const float* src, float* dst, int z_lo, int z_hi,
                    const SlabLink* link);   // tu_tma_fwd.cu
int slab_fill_maps(percnn_plan* p, const SlabLink* link, const float* src, float* dst, bool down, tma3d::SlabMaps* sm);

cudaError_t tma_bwd_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_prep, d_prep, sizeof(PrepBlock), size_t(slot) * sizeof(PrepBlock),
                                 cudaMemcpyDeviceToDevice, st);
}

int tma_bwd_setup(percnn_plan* p) {
  cudaError_t ae = cudaSuccess;
  switch (p->slot) {
#define PERCNN_TMA_ATTR(S)                                                                                              \
  case S:                                                                                                               \
    ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD); \
    if (ae == cudaSuccess)                                                                                              \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD_SLAB); \
    if (ae == cudaSuccess)                                                                                              \
      ae = cudaFuncSetAttribute(tma3d::k_gs3d_bwd_tma<S, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma3d::SMEM_BYTES_BWD_SLAB); \
    break;
    PERCNN_TMA_ATTR(0) PERCNN_TMA_ATTR(1) PERCNN_TMA_ATTR(2) PERCNN_TMA_ATTR(3) PERCNN_TMA_ATTR(4) PERCNN_TMA_ATTR(5)
#undef PERCNN_TMA_ATTR
    default: return fail(PERCNN_ERR_INVALID, "bad parameter slot");
  }
  if (ae != cudaSuccess) return fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(tma bwd) failed");
  return PERCNN_OK;
}

// G (= gout) streams through the TMA ring; h, gadd and the output gin share its layout.
int tma_bwd_launch(percnn_plan* p, const float* h, const float* gout, const float* gadd, float* gin, char* ws,
                   cudaStream_t st, const SlabLink* link, const InjectHost* ih) {
  const CUtensorMap *mm, *hm;
  int rc = get_maps(p, gout, &mm, &hm);
  if (rc) return rc;
  tma3d::Params prm;
  const int grid = tma_fill_params(p, prm, gout, gin, 0, p->g.D, link);
  tma3d::BwdExtra x;
  x.h = h;
  x.gadd = gadd;
  x.partials = reinterpret_cast<double*>(ws + kWsPartials);
  x.counter = reinterpret_cast<unsigned*>(ws + kWsCounter);
  x.acc = reinterpret_cast<double*>(ws + kWsAcc);
  x.inj = make_inject<float>(p, ih);
  const bool down = link && (link->epoch_wait & 1u);
  tma3d::SlabMaps sm;
  if (link) {
    rc = slab_fill_maps(p, link, gout, gin, down, &sm);
    if (rc) return rc;
  } else {
    memset(&sm, 0, sizeof(sm));
  }
  cudaError_t le = cudaSuccess;
  switch (p->slot) {
#define PERCNN_TMA_CASE(S)                                                                                              \
  case S:                                                                                                               \
    if (!link) le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, false, false>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD, st, p->pdl, *mm, *hm, prm, x, sm); \
    else if (down) le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, true, true>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD_SLAB, st, p->pdl, *mm, *hm, prm, x, sm); \
    else le = launch_pdl(tma3d::k_gs3d_bwd_tma<S, true, false>, grid, tma3d::BWD_THREADS, tma3d::SMEM_BYTES_BWD_SLAB, st, p->pdl, *mm, *hm, prm, x, sm); \
    break;
    PERCNN_TMA_CASE(0) PERCNN_TMA_CASE(1) PERCNN_TMA_CASE(2) PERCNN_TMA_CASE(3) PERCNN_TMA_CASE(4) PERCNN_TMA_CASE(5)
#undef PERCNN_TMA_CASE
    default: return fail(PERCNN_ERR_INVALID, "bad parameter slot");
  }
  if (le != cudaSuccess) return fail(PERCNN_ERR_CUDA, std::string("TMA adjoint kernel launch: ") + cudaGetErrorString(le));
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

}  // namespace percnn
