// Translation unit: the 5x5 Pi-block cell (BUR1 / LO1), forward and adjoint.
#include "kernels_pi_k5.cuh"
#include "plan.h"

namespace percnn {

cudaError_t k5_load_prep(const PrepBlock* d_prep, int slot, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_prep, d_prep, sizeof(PrepBlock), size_t(slot) * sizeof(PrepBlock),
                                 cudaMemcpyDeviceToDevice, st);
}

int k5_blocks(const percnn_plan* p) {
  return ((p->g.W + k5::BT_X - 1) / k5::BT_X) * ((p->g.H + k5::BT_Y - 1) / k5::BT_Y);
}

int k5_setup(percnn_plan* p) {
  const int hc = p->desc.hidden;
  if (cudaMalloc(&p->d_k5w, size_t(k5_total_floats(hc)) * 4) != cudaSuccess) return fail(PERCNN_ERR_CUDA, "cudaMalloc(k5w) failed");
  if (cudaFuncSetAttribute(k5::k_pi_k5_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, int(k5::smem_bytes(hc))) != cudaSuccess)
    return fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(k5) failed");
  if (cudaFuncSetAttribute(k5::k_pi_k5_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           int(k5::bwd_smem_floats(hc, int(p->nparams)) * sizeof(float))) != cudaSuccess)
    return fail(PERCNN_ERR_CUDA, "cudaFuncSetAttribute(k5 bwd) failed");
  return PERCNN_OK;
}

int k5_step_fwd(percnn_plan* p, const float* src, float* dst, cudaStream_t st) {
  dim3 grid((p->g.W + k5::TILE_X - 1) / k5::TILE_X, (p->g.H + k5::TILE_Y - 1) / k5::TILE_Y, 2);   // z = output field
  k5::k_pi_k5_fwd<<<grid, k5::THREADS, k5::smem_bytes(p->desc.hidden), st>>>(p->g, p->slot, p->desc.hidden, src, dst,
                                                                              p->d_k5w);
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

// Adjoint step: per-block partial parameter sums (raw packing) -> `partials`, folded in fp64 into `acc`.
int k5_step_bwd(percnn_plan* p, const float* h, const float* gout, const float* gadd, float* gin, double* acc,
                float* partials, cudaStream_t st) {
  dim3 grid((p->g.W + k5::BT_X - 1) / k5::BT_X, (p->g.H + k5::BT_Y - 1) / k5::BT_Y);
  const int np = int(p->nparams);
  k5::k_pi_k5_bwd<<<grid, k5::BTHREADS, k5::bwd_smem_floats(p->desc.hidden, np) * sizeof(float), st>>>(
      p->g, p->slot, p->desc.hidden, np, h, gout, gadd, gin, p->d_k5w, partials);
  PERCNN_CUDA(cudaGetLastError());
  k5::k5_reduce_partials<<<(np + 255) / 256, 256, 0, st>>>(partials, int(grid.x * grid.y), np, acc);
  PERCNN_CUDA(cudaGetLastError());
  p->launches += 2;
  return PERCNN_OK;
}

int k5_grads_finish(percnn_plan* p, const float* params, const double* acc, float* grads, cudaStream_t st) {
  const int np = int(p->nparams);
  k5::k5_finish<<<(np + 255) / 256, 256, 0, st>>>(params, acc, p->pd, np, grads);
  PERCNN_CUDA(cudaGetLastError());
  p->launches++;
  return PERCNN_OK;
}

}  // namespace percnn
