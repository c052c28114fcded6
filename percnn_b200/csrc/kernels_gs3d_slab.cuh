// Slab-mode forward step with the halo exchange fused into the kernel (multi-GPU).
//
// One process per GPU owns D planes of the grid plus 2 ghost planes per side ([2][D+4][H][W]); the ring
// neighbours' buffers are peer-mapped over NVLink.  ONE kernel per time step:
//
//   * every CTA marches its tile through ALL D planes in a single pass (same consumer code as
//     k_gs3d_fwd_tma: TMA ring, register z-window, shuffle seam) -- D + 4 plane loads for D output planes;
//   * output planes 0,1 are stored locally AND into the lower neighbour's upper ghost planes, planes D-2,D-1 into
//     the upper neighbour's lower ghost planes (plain st.global through the peer mapping);
//   * when all tiles have stored a boundary pair, the last arriving CTA raises that neighbour's flag
//     (st.release.sys); the producer lane waits (ld.acquire.sys) on its own flag right before the first TMA that
//     touches the corresponding ghost planes.
//
// The march direction ALTERNATES from step to step (DOWN = odd epochs).  An upward step produces planes 0,1
// first and D-2,D-1 last; the following downward step consumes the upper ghosts first and the lower ghosts
// last.  So every ghost plane is produced almost a full step before it is consumed, on both sides, and neither
// the NVLink latency nor the neighbours' skew is ever on the critical path.  (Round 1 ran three z-segments per
// step -- both boundary pairs first, then the interior -- which cost 12 extra warm-up planes per step, three
// pipeline fills per CTA and 2-plane items that under-filled the machine: 0.65 scaling efficiency at 8 GPUs.)
//
// Write-after-read on the ghost planes needs no extra handshake: a rank overwrites a neighbour's ghost planes of
// buffer A while computing the boundary planes that depend on ITS OWN ghosts of buffer B, and the flag it waits on
// for those is raised by that neighbour after the very planes that read A's ghosts were completed.
// The result is bit-identical to the single-GPU rollout for either direction (the z-window is handed to the
// stencil in ascending-z order both ways).
#pragma once
#include "kernels_gs3d_tma.cuh"

namespace percnn {
namespace tma3d {

// Consumer-side: all consumer warps of the CTA have stored (locally and to the peer) the boundary pair of one
// tile.  Every storing warp fences at system scope itself -- its peer (NVLink) stores must be performed before the
// flag can be observed; relying on one thread's fence after the CTA barrier to cover the other warps' in-flight
// peer stores produced stale ghost planes on a neighbour in round 1 (caught by the 2-GPU bitwise test).
__device__ __forceinline__ void slab_post(const Params& p, int warp, int lane, uint32_t* counter, uint32_t* flag, int ntiles) {
  __threadfence_system();
  asm volatile("bar.sync 1, %0;" ::"r"(p.ty * 32) : "memory");
  if (warp == 0 && lane == 0) {
    __threadfence_system();
    const unsigned old = atomicAdd(counter, 1u);
    if (old == unsigned(ntiles) - 1u) {
      atomicExch(counter, 0u);
      __threadfence_system();
      st_release_sys(flag, p.epoch_post);
    }
  }
}

template <int SLOT, bool DOWN>
__global__ void __launch_bounds__(THREADS, 1)
k_gs3d_fwd_slab(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
                const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], p.ty);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see k_gs3d_fwd_tma
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int nitems = total_items(p);

  if (warp >= TY) {
    // ===== producer warp-group: one elected lane issues every TMA =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
    if (warp == TY && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_main)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      uint32_t it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const ItemCoord ic = decode_item(p, item);
        bool wait_lo = !(p.debug & 2), wait_hi = wait_lo;
        int yh[4] = {ic.y0 - 2, ic.y0 - 1, ic.y0 + p.ty, ic.y0 + p.ty + 1};   // periodic halo rows
#pragma unroll
        for (int h = 0; h < 4; ++h) yh[h] = yh[h] < 0 ? yh[h] + p.H : (yh[h] >= p.H ? yh[h] - p.H : yh[h]);
        const uint32_t bytes_main = 2u * uint32_t(p.ty) * TX * 4u, bytes_halo = 2u * 4u * TX * 4u;
        for (int k = 0; k < ic.nz + 4; ++k, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          // interior index of local plane k: ascending from z0 - 2, or descending from z0 + nz + 1
          const int zi = DOWN ? ic.z0 + ic.nz + 1 - k : ic.z0 + k - 2;
          // ghost planes are written by a neighbour GPU: wait for its flag, then order the TMA (async proxy)
          // reads after the acquire
          if (zi < 0 && wait_lo) {
            wait_flag(p.my_flags + 0, p.epoch_wait, p.scratch + 1, p.spin_limit);
            asm volatile("fence.proxy.async.global;" ::: "memory");
            wait_lo = false;
          }
          if (zi >= p.D && wait_hi) {
            wait_flag(p.my_flags + 1, p.epoch_wait, p.scratch + 1, p.spin_limit);
            asm volatile("fence.proxy.async.global;" ::: "memory");
            wait_hi = false;
          }
          const bool with_halo = (k >= 2) && (k < ic.nz + 2);
          const int pz = zi + 2;   // ghosted buffer: plane index = interior index + 2
          float* st = ring + s * STAGE_FLOATS;
          mbar_expect_tx(&full[s], with_halo ? bytes_main + bytes_halo : bytes_main);
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            float* sf = st + f * ROWS * TX;
            tma_load_4d(sf + 2 * TX, &tm_main, &full[s], ic.x0, ic.y0, pz, f);
            if (with_halo) {
              tma_load_4d(sf, &tm_halo, &full[s], ic.x0, yh[0], pz, f);
              tma_load_4d(sf + TX, &tm_halo, &full[s], ic.x0, yh[1], pz, f);
              tma_load_4d(sf + (p.ty + 2) * TX, &tm_halo, &full[s], ic.x0, yh[2], pz, f);
              tma_load_4d(sf + (p.ty + 3) * TX, &tm_halo, &full[s], ic.x0, yh[3], pz, f);
            }
          }
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
  if (warp >= p.ty) return;
  Consumer c;
  c.P = c_prep[SLOT].f;
  c.ring = ring;
  c.full = full;
  c.empty = empty;
  c.s = 0;
  c.parity = 0;
  c.row = warp;
  c.lane = lane;
  c.toff = uint32_t(warp) * uint32_t(p.W) + 4u * uint32_t(lane);
  c.is_seam = (lane == 0) || (lane == 31);
  const int64_t plane = int64_t(p.H) * p.W;
  const int64_t zstep = DOWN ? -plane : plane;
  const int ntiles = p.nxt * p.nyt;
  float4 wu[5], wv[5];
  float2 seam_next[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    const int64_t tile_off = int64_t(ic.y0) * p.W + ic.x0;
    const int zfirst = DOWN ? ic.z0 + ic.nz - 1 : ic.z0;   // interior index of the first output plane
    float* out = p.dst + int64_t(zfirst + 2) * plane + tile_off + c.toff;
    // boundary planes are mirrored into the neighbours' ghost planes: interior plane z < 2 is the lower
    // neighbour's ghost plane D + 2 + z, interior plane z >= D - 2 the upper neighbour's ghost plane z - (D - 2)
    float* const mir_lo = (p.debug & 1) ? nullptr : p.peer_lo_dst + int64_t(p.D + 2) * plane + tile_off + c.toff;
    float* const mir_hi = (p.debug & 1) ? nullptr : p.peer_hi_dst - int64_t(p.D - 2) * plane + tile_off + c.toff;
    int xs = (lane == 0) ? ic.x0 - 2 : ic.x0 + TX;
    xs = xs < 0 ? xs + p.W : (xs >= p.W ? xs - p.W : xs);
    const int seam_off = warp * p.W + xs - ic.x0;
    // seam cells of the plane that is the in-plane source next (the first output plane first)
    const float* seam_ptr = p.src + int64_t(zfirst + 2) * plane + tile_off + seam_off;

    warm_plane<0>(c, true, wu, wv);
    warm_plane<1>(c, true, wu, wv);
    warm_plane<2>(c, false, wu, wv);
    warm_plane<3>(c, false, wu, wv);
    ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
    ldg_f2_if(c.is_seam, seam_ptr + p.src_field, seam_next[1]);

    const int nk = ic.nz + 4;   // local planes 0 .. nz+3; outputs for k = 4 .. nz+3
    int zi = zfirst;            // interior index of the output plane of the current iteration
#define PERCNN_SLAB_STEADY(RR)                                                                                \
  {                                                                                                           \
    seam_ptr += zstep;                                                                                        \
    float* mirror = nullptr;                                                                                  \
    if (zi < 2) mirror = mir_lo == nullptr ? nullptr : mir_lo + int64_t(zi) * plane;                          \
    else if (zi >= p.D - 2) mirror = mir_hi == nullptr ? nullptr : mir_hi + int64_t(zi) * plane;              \
    steady_plane<RR, true, DOWN>(c, k >= ic.nz + 2, k <= ic.nz + 2, seam_ptr, p.src_field, out, mirror,       \
                                 p.dst_field, wu, wv, seam_next);                                             \
    out += zstep;                                                                                             \
    if (!(p.debug & 4)) {                                                                                     \
      if (zi == (DOWN ? 0 : 1)) slab_post(p, warp, lane, p.scratch + 0, p.post_lo_flag, ntiles);              \
      if (zi == (DOWN ? p.D - 2 : p.D - 1)) slab_post(p, warp, lane, p.scratch + 2, p.post_hi_flag, ntiles);  \
    }                                                                                                         \
    zi += DOWN ? -1 : 1;                                                                                      \
    ++k;                                                                                                      \
  }
    int k = 4;
    PERCNN_SLAB_STEADY(4)
    while (k + 5 <= nk) {
      PERCNN_SLAB_STEADY(0) PERCNN_SLAB_STEADY(1) PERCNN_SLAB_STEADY(2) PERCNN_SLAB_STEADY(3) PERCNN_SLAB_STEADY(4)
    }
    if (k < nk) PERCNN_SLAB_STEADY(0)
    if (k < nk) PERCNN_SLAB_STEADY(1)
    if (k < nk) PERCNN_SLAB_STEADY(2)
    if (k < nk) PERCNN_SLAB_STEADY(3)
#undef PERCNN_SLAB_STEADY
  }
}

}  // namespace tma3d
}  // namespace percnn
