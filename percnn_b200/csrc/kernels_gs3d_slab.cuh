// Slab-mode forward step with the halo exchange fused into the kernel (multi-GPU).
//
// One process per GPU owns D planes of the grid plus 2 ghost planes per side ([2][D+4][H][W]); the ring
// neighbours' buffers are peer-mapped over NVLink.  ONE kernel per time step:
//
//   * every CTA marches its tile through ALL D planes in a single pass -- D + 4 plane loads for D output planes.
//     The consumer warps run the very loop of k_gs3d_fwd_tma (TMA ring, register z-window, shuffle seam); they
//     know nothing about the exchange except one non-blocking `bar.arrive` after their second output plane;
//   * a HELPER warp (a spare consumer-group warp) moves the boundary planes: once the consumers have stored the
//     first boundary pair of the tile (planes 0,1 when marching up, D-1,D-2 when marching down) it copies them
//     from the local output buffer (L2-resident, just written) into the neighbour's ghost planes with plain
//     st.global through the peer mapping, fences at system scope, and the last tile to arrive raises that
//     neighbour's flag (red.release.sys max).  Neither the NVLink stores nor their fences ever sit in the
//     consumers' instruction stream;
//   * the producer lane waits (ld.acquire.sys) on its own flag right before the first TMA that touches the
//     corresponding ghost planes -- for the pair it needs first, already during the previous kernel's tail
//     (before griddepcontrol.wait).
//
// The march direction ALTERNATES from step to step (DOWN = odd epochs).  An upward step produces planes 0,1
// first and D-2,D-1 last; the following downward step consumes the upper ghosts first and the lower ghosts
// last.  So every ghost plane is produced almost a full step before it is consumed, on both sides, and neither
// the NVLink latency nor the neighbours' skew is ever on the critical path.
// The pair produced LAST would put copy + fence + flag into the kernel's tail, where nothing overlaps it; inside a
// rollout it is therefore DEFERRED: the next step's helper copies it from its (complete) input buffer right at
// its start, before the early pair, and one flag covers both.  Only the last step of a rollout call publishes
// its own late pair.  (Round 1 ran three z-segments per step -- both boundary pairs first, then the interior --
// 12 extra warm-up planes per step, three pipeline fills per CTA, fences in every consumer warp: 0.65 scaling
// efficiency at 8 GPUs.)
//
// Write-after-read on the ghost planes needs no extra handshake: a rank overwrites a neighbour's ghost planes of
// buffer A only after it has seen the flag of the pair it needs from that neighbour for the same planes' inputs,
// and that flag is raised after the neighbour's reads of A's ghosts were completed.
// The result is bit-identical to the single-GPU rollout for either direction (the z-window is handed to the
// stencil in ascending-z order both ways).
#pragma once
#include "kernels_gs3d_tma.cuh"

namespace percnn {
namespace tma3d {

constexpr int SLAB_MAX_TY = 14;   // (the adjoint slab kernel keeps consumer warp 14 free for its helper)
// 512 consumer-group threads x 112 + 128 producer-group threads x 32 = 640 x 96: the whole CTA allocation
#ifndef PERCNN_EXP_PRODUCER_REGS
#define PERCNN_EXP_PRODUCER_REGS 32
#endif
constexpr int SLAB_PRODUCER_REGS = PERCNN_EXP_PRODUCER_REGS;

__device__ __forceinline__ void red_release_sys_max(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ---- the helper's tensor maps: my two buffers once more with a box of one field's boundary PAIR ----
// box = {128 cells, ty rows, 2 planes, 1 field}: one TMA load brings a field's pair of this tile from the local
// buffer (L2-resident, just written) into shared memory; the helper warp then sends it to the neighbour with plain
// 128-bit stores through the peer mapping.  Measured alternatives (2 GPUs, 64 planes of 512^2 per rank, 52 us of
// compute per step): TMA *stores* to the peer (cp.async.bulk.tensor ... bulk_group + wait_group) stretched every
// step by 11 us -- the kernel cannot end before the NVLink writes of its bulk group have been acknowledged, and
// they trickle; per-row bulk copies without a tensor map (56 + 56 instructions per pair from one lane) and a
// register copy straight from global memory (two loads in flight per lane) both took 20-30 us per step.
struct alignas(64) SlabMaps {
  CUtensorMap src;   // my input buffer   (deferred pair of the previous step)
  CUtensorMap dst;   // my output buffer  (early and late pair of this step)
};

// Staging area of the halo helper in shared memory (behind the ring and its barriers): one mbarrier + NBUF
// field-pairs of 2 x ty x 128 floats.  The forward kernel stages both fields at once, the adjoint (less shared
// memory left) one after the other.
constexpr int SLAB_HELPER_OFF = STAGES * STAGE_BYTES + 2 * STAGES * 8;   // the helper's mbarrier (8 B), then padding
constexpr int SLAB_STAGE_OFF = SLAB_HELPER_OFF + 128;
constexpr int SLAB_FIELD_PAIR_BYTES = 2 * SLAB_MAX_TY * TX * 4;           // capacity per staged field-pair (ty = 14)
constexpr int SMEM_BYTES_SLAB = SLAB_STAGE_OFF + 2 * SLAB_FIELD_PAIR_BYTES;

// Move one boundary pair of this tile from a local buffer into a neighbour's ghost planes (whole helper warp).
//   from_plane: buffer plane index of the pair's first plane in the local buffer (tensor map `from`)
//   to        : the pair's first plane in the peer buffer at this tile's (y0, x0), this lane's quad; `field` /
//               `plane` / W are the peer buffer's strides (same layout as mine)
template <int NBUF>
__device__ __forceinline__ void slab_copy_pair(const CUtensorMap* from, int from_plane, float* __restrict__ to, int64_t field,
                                               int64_t plane, int W, int x0, int y0, int ty, int lane, uint64_t* bar,
                                               float* stage, uint32_t& phase, int debug) {
  if (debug & 8) return;   // timing experiment: no copies
  const uint32_t bytes = 2u * uint32_t(ty) * TX * 4u;   // one field's pair as it lands in shared memory
#pragma unroll
  for (int f0 = 0; f0 < 2; f0 += NBUF) {
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the staging rows were last read through the generic proxy
      mbar_expect_tx(bar, NBUF * bytes);
#pragma unroll
      for (int i = 0; i < NBUF; ++i) tma_load_4d(stage + i * (SLAB_FIELD_PAIR_BYTES / 4), from, bar, x0, y0, from_plane, f0 + i);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (!(debug & 32)) {
#pragma unroll
      for (int i = 0; i < NBUF; ++i) {
        const float* sp = stage + i * (SLAB_FIELD_PAIR_BYTES / 4) + 4 * lane;
        float* gp = to + int64_t(f0 + i) * field;
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll 2
          for (int r = 0; r < ty; ++r)
            *reinterpret_cast<float4*>(gp + int64_t(pl) * plane + int64_t(r) * W) = lds128(sp + (pl * ty + r) * TX);
      }
    }
    __syncwarp();
  }
}

// Every lane of the helper warp has issued its peer stores: each lane makes its own visible system-wide (round 1: a
// single thread's fence did not cover other threads' in-flight NVLink stores), then the warp counts this tile and
// the last tile raises the neighbour's flag.
__device__ __forceinline__ void slab_publish(int lane, uint32_t* counter, uint32_t* flag, uint32_t epoch, int ntiles, int debug) {
  if (debug & 16) return;   // timing experiment: no fences / atomics / flag
  __threadfence_system();
  __syncwarp();
  if (lane == 0) {
    const unsigned old = atomicAdd(counter, 1u);
    if (old == unsigned(ntiles) - 1u) {
      atomicExch(counter, 0u);
      __threadfence_system();
      red_release_sys_max(flag, epoch);
    }
  }
}

// The helper warp's whole job for one kernel (shared by the forward and the adjoint slab kernels).
template <bool DOWN, int NBUF>
__device__ __forceinline__ void slab_helper(const Params& p, const SlabMaps& m, int lane, int nitems, uint64_t* bar, float* stage) {
  const int ntiles = p.nxt * p.nyt;
  const int nbar = (p.ty + 1) * 32;
  const int64_t plane = int64_t(p.H) * p.W;
  // side 0: lower neighbour (my buffer planes 2,3 -> its ghost planes D+2,D+3); side 1: upper neighbour (my buffer
  // planes D,D+1 -> its ghost planes 0,1).  Buffer plane = interior plane + 2.
  constexpr int early = DOWN ? 1 : 0, late = DOWN ? 0 : 1;
  const int from_e = early == 0 ? 2 : p.D, to_e = early == 0 ? p.D + 2 : 0;
  const int from_l = late == 0 ? 2 : p.D, to_l = late == 0 ? p.D + 2 : 0;
  float* const peer_e_src = early == 0 ? p.peer_lo_src : p.peer_hi_src;
  float* const peer_e_dst = early == 0 ? p.peer_lo_dst : p.peer_hi_dst;
  float* const peer_l_dst = late == 0 ? p.peer_lo_dst : p.peer_hi_dst;
  uint32_t phase = 0;
  if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&m.dst)) : "memory");
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    const bool own[2] = {ic.z0 == 0, ic.z0 + ic.nz == p.D};
    const int64_t off = int64_t(ic.y0) * p.W + ic.x0 + 4 * lane;
    if (own[early]) {
      if (p.flush_prev)   // the previous step left this pair (its LAST one) in what is now my input buffer
        slab_copy_pair<NBUF>(&m.src, from_e, peer_e_src + int64_t(to_e) * plane + off, p.src_field, plane, p.W, ic.x0, ic.y0,
                             p.ty, lane, bar, stage, phase, p.debug);
      asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");   // the consumers have stored the pair ...
      if (lane == 0) asm volatile("fence.proxy.async.global;" ::: "memory");   // ... through the generic proxy; TMA reads it through the async one
      slab_copy_pair<NBUF>(&m.dst, from_e, peer_e_dst + int64_t(to_e) * plane + off, p.dst_field, plane, p.W, ic.x0, ic.y0, p.ty,
                           lane, bar, stage, phase, p.debug);
      slab_publish(lane, p.scratch + (early == 0 ? 0 : 2), early == 0 ? p.post_lo_flag : p.post_hi_flag, p.epoch_post, ntiles, p.debug);
    }
    if (own[late] && !p.defer_late) {
      asm volatile("bar.sync 2, %0;" ::"r"(nbar) : "memory");
      if (lane == 0) asm volatile("fence.proxy.async.global;" ::: "memory");
      slab_copy_pair<NBUF>(&m.dst, from_l, peer_l_dst + int64_t(to_l) * plane + off, p.dst_field, plane, p.W, ic.x0, ic.y0, p.ty,
                           lane, bar, stage, phase, p.debug);
      slab_publish(lane, p.scratch + (late == 0 ? 0 : 2), late == 0 ? p.post_lo_flag : p.post_hi_flag, p.epoch_post, ntiles, p.debug);
    }
  }
}

// Consumer side of the two named barriers: WHICH = 1 after the item's second output plane (its first boundary pair
// is stored), WHICH = 2 after its last plane.  With one item per CTA the consumers only signal (bar.arrive) and
// run on; a CTA that processes several items must not lap the helper, so there they wait for it (bar.sync).
// Everything is recomputed from the kernel parameters here, inside a branch taken once per item: the plane loop
// sits exactly at the kernel's register budget and must not carry any extra live value for this.
// A real call (noinline): inlined, its address arithmetic and integer divisions pushed the register allocator into
// spilling inside the plane loop (ptxas -v: 340 B of spill stores); as a call, the save/restore of the caller's
// live registers happens inside the once-per-item branch only.
static __device__ __noinline__ void slab_consumer_signal_fn(const Params* pp, int item, int which, int down) {
  const Params& p = *pp;
  if (p.debug & 1) return;
  const ItemCoord ic = decode_item(p, item);
  const bool lo = ic.z0 == 0, hi = ic.z0 + ic.nz == p.D;
  const bool own = which == 1 ? (down ? hi : lo) : (!p.defer_late && (down ? lo : hi));
  if (!own) return;
  const int nbar = (p.ty + 1) * 32;
  if (total_items(p) > int(gridDim.x)) asm volatile("bar.sync %0, %1;" ::"r"(which), "r"(nbar) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"r"(which), "r"(nbar) : "memory");
}
template <bool DOWN, int WHICH>
__device__ __forceinline__ void slab_consumer_signal(const Params& p, int item) {
  slab_consumer_signal_fn(&p, item, WHICH, DOWN ? 1 : 0);
}

// Producer side: interior index of local plane k, and the flag waits that guard the ghost planes.
template <bool DOWN>
__device__ __forceinline__ int slab_plane_index(const Params& p, const ItemCoord& ic, int k, bool& wait_lo, bool& wait_hi) {
  const int zi = DOWN ? ic.z0 + ic.nz + 1 - k : ic.z0 + k - 2;
  // ghost planes are written by a neighbour GPU: wait for its flag, then order the TMA (async proxy) reads after
  // the acquire
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
  return zi;
}

// ONE kernel body for the periodic single-GPU step (MODE 0) and the two slab-mode marches (MODE 1: ascending z,
// MODE 2: descending z).  The consumer loop of MODE 0 sits exactly at its register/schedule optimum (DESIGN.md 3.3:
// a 12-instruction clean-up once cost 28 %), and a separately written slab kernel with "the same" loop compiled to
// a schedule 10-25 % slower.  So the slab modes are derived from the very same source text: everything they add
// lives in the producer warp-group (flag waits, the halo helper warp) or in once-per-item calls.
// SLOT is a template parameter so that every coefficient is a compile-time constant-bank address
// (c[3][imm] / hoisted LDCU) instead of an indexed LDC per use.
template <int SLOT, int MODE>
__device__ __forceinline__ void gs3d_fwd_body(const CUtensorMap& tm_main, const CUtensorMap& tm_halo, const Params& p,
                                              const SlabMaps* sm) {
  constexpr bool SLAB = MODE != 0, DOWN = MODE == 2;
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
    if constexpr (SLAB) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + SLAB_HELPER_OFF), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // Programmatic dependent launch: the next step's grid may be scheduled while this one drains (its CTAs take
  // the SMs our CTAs leave and run their prologue), and this grid touches global memory only after the previous
  // one has completed (it wrote our input and reads the buffer we overwrite).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int nitems = total_items(p);

  if (warp >= TY) {
    // ===== producer warp-group: one elected lane issues every TMA (slab modes: one more warp is the halo helper) =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SLAB ? SLAB_PRODUCER_REGS : PRODUCER_REGS));
    if (warp == TY && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_main)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      uint32_t it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const ItemCoord ic = decode_item(p, item);
        bool wait_lo = SLAB && !(p.debug & 2), wait_hi = wait_lo;
        int yh[4] = {ic.y0 - 2, ic.y0 - 1, ic.y0 + p.ty, ic.y0 + p.ty + 1};   // periodic halo rows
#pragma unroll
        for (int h = 0; h < 4; ++h) yh[h] = yh[h] < 0 ? yh[h] + p.H : (yh[h] >= p.H ? yh[h] - p.H : yh[h]);
        const uint32_t bytes_main = 2u * uint32_t(p.ty) * TX * 4u, bytes_halo = 2u * 4u * TX * 4u;
        for (int k = 0; k < ic.nz + 4; ++k, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          const bool with_halo = (k >= 2) && (k < ic.nz + 2);
          int pz;
          if constexpr (SLAB) pz = slab_plane_index<DOWN>(p, ic, k, wait_lo, wait_hi) + 2;   // ghosted buffer: plane = interior + 2
          else pz = src_plane(p, ic.z0, k);
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
    } else if (SLAB && warp == TY + 1) {
      if (!(p.debug & 1))
        slab_helper<DOWN, 2>(p, *sm, lane, nitems, reinterpret_cast<uint64_t*>(smem_raw + SLAB_HELPER_OFF),
                             reinterpret_cast<float*>(smem_raw + SLAB_STAGE_OFF));
    }
    return;
  }

  // ===== consumer warps =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
  if (warp >= p.ty) return;   // tile shorter than 16 rows: the spare warps are done (after the aligned setmaxnreg)
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
  float4 wu[5], wv[5];
  float2 seam_next[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    // uniform per-item bases; the per-lane part (toff / seam_off) never changes
    const float* src_xy = p.src + int64_t(ic.y0) * p.W + ic.x0;
    float* out = p.dst + (int64_t((DOWN ? ic.z0 + ic.nz - 1 : ic.z0) + p.dst_zoff) * p.H + ic.y0) * p.W + ic.x0 + c.toff;
    int xs = (lane == 0) ? ic.x0 - 2 : ic.x0 + TX;
    xs = xs < 0 ? xs + p.W : (xs >= p.W ? xs - p.W : xs);
    const int seam_off = warp * p.W + xs - ic.x0;
    // source plane whose seam cells are fetched next (local plane 2 first = the first output plane's own plane)
    int pz = DOWN ? ic.z0 + ic.nz + 1 : src_plane(p, ic.z0, 2);
    const float* seam_ptr = src_xy + int64_t(pz) * plane + seam_off;   // advanced by one plane per iteration
    const int64_t wrap_back = int64_t(p.D) * plane;

    warm_plane<0>(c, true, wu, wv);
    warm_plane<1>(c, true, wu, wv);
    warm_plane<2>(c, false, wu, wv);
    warm_plane<3>(c, false, wu, wv);
    ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
    ldg_f2_if(c.is_seam, seam_ptr + p.src_field, seam_next[1]);

    const int nk = ic.nz + 4;   // local planes 0 .. nz+3; outputs for k = 4 .. nz+3
#define PERCNN_STEADY(RR)                                                                                     \
  {                                                                                                           \
    seam_ptr += DOWN ? -plane : plane;                                                                        \
    if (p.wrap_z && ++pz >= p.D) {                                                                            \
      pz -= p.D;                                                                                              \
      seam_ptr -= wrap_back;                                                                                  \
    }                                                                                                         \
    steady_plane<RR, false, DOWN>(c, k >= ic.nz + 2, k <= ic.nz + 2, seam_ptr, p.src_field, out, nullptr,     \
                                  p.dst_field, wu, wv, seam_next);                                            \
    out += DOWN ? -plane : plane;                                                                             \
    ++k;                                                                                                      \
  }
    int k = 4;
    PERCNN_STEADY(4)
    while (k + 5 <= nk) {
      PERCNN_STEADY(0)
      if constexpr (SLAB)
        if (k == 6) slab_consumer_signal<DOWN, 1>(p, item);   // first boundary pair stored: over to the helper
      PERCNN_STEADY(1) PERCNN_STEADY(2) PERCNN_STEADY(3) PERCNN_STEADY(4)
    }
    if (k < nk) {
      PERCNN_STEADY(0)
      if constexpr (SLAB)
        if (k == 6) slab_consumer_signal<DOWN, 1>(p, item);
    }
    if (k < nk) PERCNN_STEADY(1)
    if (k < nk) PERCNN_STEADY(2)
    if (k < nk) PERCNN_STEADY(3)
#undef PERCNN_STEADY
    if constexpr (SLAB) slab_consumer_signal<DOWN, 2>(p, item);
  }
}

template <int SLOT>
__global__ void __launch_bounds__(THREADS, 1)
k_gs3d_fwd_tma(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
               const __grid_constant__ Params p) {
  gs3d_fwd_body<SLOT, 0>(tm_main, tm_halo, p, nullptr);
}

template <int SLOT, bool DOWN>
__global__ void __launch_bounds__(THREADS, 1)
k_gs3d_fwd_slab(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
                const __grid_constant__ Params p, const __grid_constant__ SlabMaps sm) {
  gs3d_fwd_body<SLOT, DOWN ? 2 : 1>(tm_main, tm_halo, p, &sm);
}

}  // namespace tma3d
}  // namespace percnn
