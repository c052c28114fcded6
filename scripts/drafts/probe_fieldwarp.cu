// ptxas probe for the draft kernel (register footprint only; nothing here is ever launched):
//   cd scripts/drafts && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -c probe_fieldwarp.cu -o /tmp/probe.o
#include "kernels_gs3d_bwd_fieldwarp.cuh"
void* percnn_probe_fieldwarp_symbols[] = {(void*)percnn::tma3d::k_gs3d_bwd_tma_fw<0, false>,
                                          (void*)percnn::tma3d::k_gs3d_bwd_tma_fw<0, true>};
