// Adjoint of the fused 3-D Pi-block step (k = 1), same streaming structure as kernels_gs3d_tma.cuh:
//
//   g_u = g_add_u + Gu + dt (alpha_u Lap^T Gu + Gu dRu/du + Gv dRv/du)          
//   g_v = g_add_v + Gv + dt (alpha_v Lap^T Gv + Gu dRu/dv + Gv dRv/dv)
//   22 reductions per step: sum q dt Lap^T(Gq) (-> dL/dalpha_q) and sum dt G_f u^a v^b (-> the folded cubic)
//
// G (the incoming gradient) streams through the TMA ring exactly like the state does in the forward kernel
// (y-neighbours from shared memory, x-neighbours by shuffle, seam lanes from global; the z-neighbours are read
// from the ring too instead of a register window);
// the stored state h_t (centre only, no halo) and the injected loss gradient g_add are read with coalesced
// 128-bit loads one plane ahead.  Algorithmic traffic: 24 B/cell (+8 with g_add).
// Reductions: per-lane fp32 partial sums in registers, flushed every 32 planes into per-warp fp64 accumulators in
// shared memory; per-CTA results go to global memory and the last CTA folds them in fixed order (deterministic).
// Round-2 history of this kernel (813 -> 707-714 us per 512^3 step; every step A/B-timed, profiles/
// r02_adjoint_variants.txt): monomial sums from shared memory into registers, an aligned barrier at the consumers'
// entry and no CALL / top-level spin loop on their path (keeps the loop on the uniform datapath), state loads
// before the ring wait, last-CTA fold shared by all warps.  Tried and dropped: compile-time ring stages through a
// switch (8x the loop body, instruction-cache bound, +36 %), no L2 prefetch (neutral).
#pragma once
#include "kernels_gs3d_slab.cuh"

namespace percnn {
namespace tma3d {

constexpr int BWD_FLUSH = 32;
// The adjoint holds 22 running sums and the state on top of the 5-plane window: 15 consumer warps + 1 producer
// warp = 512 threads = 128 registers per thread (a 17th warp would round the allocation down to 96).
constexpr int BWD_WARPS = 15;
constexpr int BWD_THREADS = (BWD_WARPS + 1) * 32;
constexpr int SMEM_BYTES_BWD = SMEM_BYTES + 16 * kRedPiK1 * 8 + 64;
// slab mode: the halo helper's staging rows come after everything else (half a boundary pair: two chunks per pair)
constexpr int BWD_STAGE_OFF = (SMEM_BYTES_BWD + 127) / 128 * 128;   // TMA destinations are 128-byte aligned
constexpr int SMEM_BYTES_BWD_SLAB = BWD_STAGE_OFF + SLAB_FIELD_PAIR_BYTES;
static_assert(SMEM_BYTES_BWD_SLAB <= 227 * 1024, "shared memory budget of the slab adjoint");

struct BwdExtra {
  const float* h;        // stored state of this step, same layout as the G buffers
  const float* gadd;     // injected gradient for this step (nullable)
  double* partials;      // [gridDim.x][2 or 22]
  unsigned* counter;
  double* acc;           // [22] running sums over steps
  Inject<float> inj;     // fused data-loss gradient of this step's state (target == nullptr: none)
};

typedef float2 (&MonoAcc)[10];   // the 20 per-lane monomial sums: (sum_u, sum_v) per monomial, in registers

__device__ __forceinline__ float2 quad2(const float* __restrict__ d, float2 u, float2 v) {
  float2 a0 = fma2(u, fma2(u, d[3], d[1]), d[0]);
  float2 a1 = fma2(u, d[4], d[2]);
  return fma2(v, fma2(v, d[5], a1), a0);
}
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void slab_signal_inline(int which, bool sync_mode, int nbar) {
  if (sync_mode)
    asm volatile("bar.sync %0, %1;" ::"r"(which), "r"(nbar) : "memory");
  else
    asm volatile("bar.arrive %0, %1;" ::"r"(which), "r"(nbar) : "memory");
}

// One output plane of the adjoint.  Unlike the forward kernel there is no register window: the five plane
// centres (z-2..z+2) are all read from the ring (planes k-4..k stay resident), which frees 32 registers for the
// state, the injected gradient and the pointwise Jacobian -- the windowed version spilled.
// `valid`: this warp's row is not a duplicate of the previous tile's rows (last tile of a column is shifted
// back), so it contributes to the reductions.
// `DOWN`: the item is marched towards decreasing z (slab kernel, odd steps); `zstep` = +-plane accordingly.
template <bool FUSED, bool DOWN>
__device__ __forceinline__ void adjoint_plane(Consumer& c, const float* __restrict__ TP, bool prefetch_seam,
                                              const float* seam_ptr, int64_t field, int64_t zstep, int64_t off,
                                              float* __restrict__ dst, float* mirror, const float* __restrict__ hbase,
                                              const float* __restrict__ gadd, bool prefetch_next, bool valid,
                                              float2 (&seam_next)[2], float (&aacc)[2], MonoAcc macc,
                                              const Inject<float>& inj, int64_t inj_row, int xq) {
  const float* P = c.P;
  const uint32_t cs = c.s;
  // stored state of this step: issued BEFORE the wait for the ring (it does not depend on it: ~740 -> 707 us per 512^3
  // step), L2-prefetched one plane ahead
  const float4 hu = ldg128(hbase + off);
  const float4 hv = ldg128(hbase + off + field);
  mbar_wait(&c.full[cs], c.parity);   // plane k has landed; planes k-4 .. k-1 are still resident
  const float2 seam_u = seam_next[0], seam_v = seam_next[1];
  if (prefetch_seam) {
    ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
    ldg_f2_if(c.is_seam, seam_ptr + field, seam_next[1]);
  }
  if (prefetch_next && (c.lane & 7) == 0) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off + zstep));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off + zstep + field));
    if (gadd != nullptr) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gadd + off + zstep));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gadd + off + zstep + field));
    }
  }
  const uint32_t lane_off = (c.row + 2) * TX + 4 * c.lane;
  float2 Lu_lo, Lu_hi, Lv_lo, Lv_hi;
  float4 Gu, Gv;
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    float4 win[5];
#pragma unroll
    for (int j = 0; j < 5; ++j)   // local plane k-4+j sits in stage (s + STAGES - 4 + j) % STAGES; win[] is ascending in z
      win[j] = lds128(c.ring + ((cs + STAGES - (DOWN ? j : 4 - j)) & (STAGES - 1)) * STAGE_FLOATS + f * ROWS * TX + lane_off);
    const float* sp = c.ring + ((cs + STAGES - 2) & (STAGES - 1)) * STAGE_FLOATS + f * ROWS * TX + c.row * TX + 4 * c.lane;
    const float4 y[4] = {lds128(sp), lds128(sp + TX), lds128(sp + 3 * TX), lds128(sp + 4 * TX)};
    const float4 ctr = win[2];
    float Lz = __shfl_up_sync(0xffffffffu, ctr.z, 1), Lw = __shfl_up_sync(0xffffffffu, ctr.w, 1);
    float Rx = __shfl_down_sync(0xffffffffu, ctr.x, 1), Ry = __shfl_down_sync(0xffffffffu, ctr.y, 1);
    const float2 seam = f == 0 ? seam_u : seam_v;
    if (c.lane == 0) { Lz = seam.x; Lw = seam.y; }
    if (c.lane == 31) { Rx = seam.x; Ry = seam.y; }
    if (f == 0) {
      lap_quad(TP, win, y, Lz, Lw, Rx, Ry, Lu_lo, Lu_hi);
      Gu = ctr;
    } else {
      lap_quad(TP, win, y, Lz, Lw, Rx, Ry, Lv_lo, Lv_hi);
      Gv = ctr;
    }
  }
  // plane k-4 is no longer needed by this warp (release only once the loads have completed, see mbar_arrive_after)
  __syncwarp();
  // (the release must NOT wait for the long-latency global loads of h)
  if (c.lane == 0) mbar_arrive_after(&c.empty[(cs + STAGES - 4) & (STAGES - 1)], Lu_lo.x + Lv_lo.x);
  float4 au4 = make_float4(0.f, 0.f, 0.f, 0.f), av4 = au4;
  if (gadd != nullptr) {
    au4 = ldg128(gadd + off);
    av4 = ldg128(gadd + off + field);
  }
  const float alpha_u = P[P_ALPHA + 0], alpha_v = P[P_ALPHA + 1], dt = P[P_DT];
  const float* D = P + P_DPOLY;
  float4 ou, ov;
#define PERCNN_BWD_PAIR(U2, V2, GU2, GV2, LU2, LV2, OU0, OU1, OV0, OV1, AU0, AU1, AV0, AV1)                     \
  {                                                                                                            \
    const float2 gdu = mul2(GU2, dt), gdv = mul2(GV2, dt);                                                     \
    const float2 su = fma2(gdu, quad2(D + 0, U2, V2), __fmul2_rn(gdv, quad2(D + 12, U2, V2)));                 \
    const float2 sv = fma2(gdu, quad2(D + 6, U2, V2), __fmul2_rn(gdv, quad2(D + 18, U2, V2)));                 \
    const float2 lu = mul2(LU2, dt), lv = mul2(LV2, dt);                                                       \
    float2 gu = __fadd2_rn(GU2, fma2(lu, alpha_u, su));                                                        \
    float2 gv = __fadd2_rn(GV2, fma2(lv, alpha_v, sv));                                                        \
    gu = __fadd2_rn(gu, make_float2(AU0, AU1));                                                                \
    gv = __fadd2_rn(gv, make_float2(AV0, AV1));                                                                \
    OU0 = gu.x; OU1 = gu.y; OV0 = gv.x; OV1 = gv.y;                                                            \
    if (valid) {                                                                                               \
      aacc[0] = fmaf(U2.x, lu.x, fmaf(U2.y, lu.y, aacc[0]));                                                   \
      aacc[1] = fmaf(V2.x, lv.x, fmaf(V2.y, lv.y, aacc[1]));                                                   \
    }                                                                                                          \
  }
  PERCNN_BWD_PAIR(lo(hu), lo(hv), lo(Gu), lo(Gv), Lu_lo, Lv_lo, ou.x, ou.y, ov.x, ov.y, au4.x, au4.y, av4.x, av4.y)
  PERCNN_BWD_PAIR(hi(hu), hi(hv), hi(Gu), hi(Gv), Lu_hi, Lv_hi, ou.z, ou.w, ov.z, ov.w, au4.z, au4.w, av4.z, av4.w)
#undef PERCNN_BWD_PAIR
  if (valid) {
    // 20 monomial sums  sum G_f u^a v^b  for this lane's 4 cells, added to per-lane accumulators: 10 float2 registers
    // (round 1 kept them in shared memory because the windowed kernel spilled with them; the windowless one holds them
    // at exactly 128 registers, no spills, and sheds 10 LDS.64 + 10 STS.64 per plane: 813 -> 760 us per 512^3 step,
    // profiles/r02_adjoint_variants.txt).  Everything stays in the NATURAL register pairs
    // of the 128-bit loads -- (cell0, cell1) and (cell2, cell3) -- so that no operand has to be re-packed: monomials
    // by FMUL2, products by FMUL2/FFMA2, then one FADD per field folds the pair and (sum_u, sum_v) is the float2
    // that is accumulated.  (Packing (G_u, G_v) per cell instead cost ~80 MOVs per plane: ncu r01b_ncu_bwd_512.)
    const float2 ul = lo(hu), uh = hi(hu), vl = lo(hv), vh = hi(hv);
    const float2 gul = lo(Gu), guh = hi(Gu), gvl = lo(Gv), gvh = hi(Gv);
    const float2 uul = __fmul2_rn(ul, ul), uuh = __fmul2_rn(uh, uh);
    const float2 uvl = __fmul2_rn(ul, vl), uvh = __fmul2_rn(uh, vh);
    const float2 vvl = __fmul2_rn(vl, vl), vvh = __fmul2_rn(vh, vh);
#define PERCNN_MONO(M, EL, EH)                                                                      \
  {                                                                                                 \
    const float2 el = EL, eh = EH;                                                                  \
    const float2 tu = fma2(guh, eh, __fmul2_rn(gul, el));                                           \
    const float2 tv = fma2(gvh, eh, __fmul2_rn(gvl, el));                                           \
    macc[M] = __fadd2_rn(macc[M], make_float2(tu.x + tu.y, tv.x + tv.y));                           \
  }
    {
      const float2 tu = __fadd2_rn(gul, guh), tv = __fadd2_rn(gvl, gvh);
      macc[0] = __fadd2_rn(macc[0], make_float2(tu.x + tu.y, tv.x + tv.y));
    }
    PERCNN_MONO(1, ul, uh)
    PERCNN_MONO(2, vl, vh)
    PERCNN_MONO(3, uul, uuh)
    PERCNN_MONO(4, uvl, uvh)
    PERCNN_MONO(5, vvl, vvh)
    PERCNN_MONO(6, __fmul2_rn(uul, ul), __fmul2_rn(uuh, uh))
    PERCNN_MONO(7, __fmul2_rn(uul, vl), __fmul2_rn(uuh, vh))
    PERCNN_MONO(8, __fmul2_rn(ul, vvl), __fmul2_rn(uh, vvh))
    PERCNN_MONO(9, __fmul2_rn(vvl, vl), __fmul2_rn(vvh, vh))
#undef PERCNN_MONO
  }
  if (inj_row >= 0) {
    // This (plane, row) lies on the sampling lattice of the fused data loss (warp-uniform branch, taken on the
    // selected steps only): add coef * (h - target) at the lane's cells whose x is a multiple of the stride.
    const float icoef = inject_coef(inj);
    const float hus[4] = {hu.x, hu.y, hu.z, hu.w}, hvs[4] = {hv.x, hv.y, hv.z, hv.w};
    float iu[4] = {ou.x, ou.y, ou.z, ou.w}, iv[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xg = xq + j;
      if (xg % inj.s == 0) {
        const int64_t i = inj_row + xg / inj.s;
        iu[j] = fmaf(icoef, hus[j] - __ldg(inj.target + i), iu[j]);
        iv[j] = fmaf(icoef, hvs[j] - __ldg(inj.target + inj.lfield + i), iv[j]);
      }
    }
    ou = make_float4(iu[0], iu[1], iu[2], iu[3]);
    ov = make_float4(iv[0], iv[1], iv[2], iv[3]);
  }
  *reinterpret_cast<float4*>(dst + off) = ou;
  *reinterpret_cast<float4*>(dst + off + field) = ov;
  if (FUSED && mirror != nullptr) {
    *reinterpret_cast<float4*>(mirror) = ou;
    *reinterpret_cast<float4*>(mirror + field) = ov;
  }
}

// FUSED: slab mode with the halo exchange of the gradient fused in (single z-march, direction DOWN on odd steps,
// boundary pairs moved by a helper warp, the pair produced last deferred to the next step's kernel -- see
// kernels_gs3d_slab.cuh; the protocol is the forward kernel's).  Slab plans keep ty <= SLAB_MAX_TY, so consumer warp
// BWD_WARPS - 1 is free to be the helper.
template <int SLOT, bool FUSED, bool DOWN>
__global__ void __launch_bounds__(BWD_THREADS, 1)
k_gs3d_bwd_tma(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
               const __grid_constant__ Params p, const __grid_constant__ BwdExtra x, const __grid_constant__ SlabMaps sm) {
  static_assert(FUSED || !DOWN, "only the slab kernel marches downwards");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  double* wacc = reinterpret_cast<double*>(smem_raw + STAGES * STAGE_BYTES + 2 * STAGES * 8 + 64);   // [TY][22]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], p.ty);
    }
    if (FUSED) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + SLAB_HELPER_OFF), 1);   // (sits in the padding before wacc)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < TY * kRedPiK1; i += BWD_THREADS) wacc[i] = 0.0;
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see the forward kernel
  const int nitems = total_items(p);

  if (warp >= BWD_WARPS) {
    if (lane == 0) {
      // the ghost pair this march starts from: its flag wait overlaps the previous kernel's tail.  (The spin loop lives
      // inside the producer's branch on purpose: at top level, on every warp's path, it made ptxas treat the consumer
      // loop as divergent -- 83 R2UR / 27 BRA.DIV in the slab kernels against 15 / 1 in the periodic one.)
      if (FUSED) wait_flag(p.my_flags + (DOWN ? 1 : 0), p.epoch_wait, p.scratch + 1, p.spin_limit);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_main)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      uint32_t it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const ItemCoord ic = decode_item(p, item);
        bool wait_lo = FUSED, wait_hi = FUSED;
        int yh[4] = {ic.y0 - 2, ic.y0 - 1, ic.y0 + p.ty, ic.y0 + p.ty + 1};
#pragma unroll
        for (int h = 0; h < 4; ++h) yh[h] = yh[h] < 0 ? yh[h] + p.H : (yh[h] >= p.H ? yh[h] - p.H : yh[h]);
        const uint32_t bytes_main = 2u * uint32_t(p.ty) * TX * 4u, bytes_halo = 2u * 4u * TX * 4u;
        for (int k = 0; k < ic.nz + 4; ++k, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          const bool with_halo = (k >= 2) && (k < ic.nz + 2);
          const int pz = FUSED ? slab_plane_index<DOWN>(p, ic, k, wait_lo, wait_hi) + 2 : src_plane(p, ic.z0, k);
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

  asm volatile("griddepcontrol.wait;" ::: "memory");
  // ===== consumer warps =====
  if (FUSED && warp == BWD_WARPS - 1) {   // the helper warp moves the boundary pairs to the neighbours
    slab_helper<DOWN, 1>(p, sm, lane, nitems, reinterpret_cast<uint64_t*>(smem_raw + SLAB_HELPER_OFF),
                         reinterpret_cast<float*>(smem_raw + BWD_STAGE_OFF));
    return;
  }
  if (warp >= p.ty) return;
  // An ALIGNED barrier among the consumer warps: every thread of a warp executes it together, which tells the compiler
  // that the warps are converged from here on.  Without it ptxas treats the whole consumer loop as potentially
  // divergent (the role split above branches on threadIdx): loop state lives in vector registers, every LDG / STG
  // re-materialises its memory descriptor with two R2UR (34 per plane), the shuffles sit behind BRA.DIV -- 89 R2UR and
  // 24 BRA.DIV in the kernel without it, 15 and 0 with it, 5 registers fewer, 783 -> 743 us per 512^3 adjoint step
  // (profiles/r02_adjoint_variants.txt).  The forward kernel gets the same guarantee from setmaxnreg.sync.aligned.
  asm volatile("bar.sync 4, %0;" ::"r"(p.ty * 32) : "memory");
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
  // lap_quad indexes its table as P[P_LAP_C0], P[P_LAP_AX + i]; the mirrored taps sit at P_LAPT in the same order
  const float* TP = c.P + (P_LAPT - P_LAP_C0);
  const int64_t plane = int64_t(p.H) * p.W;
  const int64_t zstep = DOWN ? -plane : plane;
  const int64_t field = p.dst_field;
  float2 seam_next[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  float aacc[2] = {0.f, 0.f};
  int since_flush = 0;
  float2 macc[10];
#pragma unroll
  for (int m = 0; m < 10; ++m) macc[m] = make_float2(0.f, 0.f);
  auto flush = [&]() {   // per-lane fp32 partial sums -> per-warp fp64 accumulators (every BWD_FLUSH planes)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float t = aacc[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
      if (lane == 0) wacc[warp * kRedPiK1 + i] += double(t);
      aacc[i] = 0.f;
    }
    const float dt = c.P[P_DT];
#pragma unroll
    for (int m = 0; m < 10; ++m) {
      float2 t = macc[m];
      macc[m] = make_float2(0.f, 0.f);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        t.x += __shfl_down_sync(0xffffffffu, t.x, off);
        t.y += __shfl_down_sync(0xffffffffu, t.y, off);
      }
      if (lane == 0) {
        wacc[warp * kRedPiK1 + 2 + m] += double(dt) * double(t.x);
        wacc[warp * kRedPiK1 + 12 + m] += double(dt) * double(t.y);
      }
    }
    since_flush = 0;
  };
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    // Hand-over to the halo helper (slab_consumer_signal_fn's logic, inline): a real CALL inside this loop made ptxas
    // treat everything after it as divergent again -- 127 R2UR and 29 BRA.DIV in the slab kernels against 15 and 1 in
    // the periodic one, +9 % instructions per plane (ncu: 64.5 M vs 59.2 M for 64 planes of 512^2).  `ic` is already
    // decoded here, so the inline form costs a handful of uniform instructions per item.
    const bool s_lo = ic.z0 == 0, s_hi = ic.z0 + ic.nz == p.D;
    const bool own_early = FUSED && !(p.debug & 1) && (DOWN ? s_hi : s_lo);
    const bool own_late = FUSED && !(p.debug & 1) && !p.defer_late && (DOWN ? s_lo : s_hi);
    const bool sync_mode = nitems > int(gridDim.x);
    const int nbar = (p.ty + 1) * 32;
    // rows below the natural start of this tile are duplicates of the previous tile (last tile shifted back)
    const bool valid = (ic.y0 + warp) >= ic.ytile * p.ty;
    // fused data loss: is this warp's row on the sampling lattice, and where does it start in the low-res frame
    const int inj_ly = (x.inj.target != nullptr && (ic.y0 + warp) % x.inj.s == 0) ? (ic.y0 + warp) / x.inj.s : -1;
    const int64_t tile_off = int64_t(ic.y0) * p.W + ic.x0;
    const int zfirst = DOWN ? ic.z0 + ic.nz - 1 : ic.z0;   // interior index of the first output plane
    int64_t off = int64_t(zfirst + p.dst_zoff) * plane + tile_off + c.toff;   // this lane's quad, first output plane
    int xs = (lane == 0) ? ic.x0 - 2 : ic.x0 + TX;
    xs = xs < 0 ? xs + p.W : (xs >= p.W ? xs - p.W : xs);
    const int seam_off = warp * p.W + xs - ic.x0;
    // seam cells come from planes [z0, z0 + nz) of the item itself: no periodic wrap needed (see the forward kernel)
    const float* seam_ptr = p.src + int64_t(FUSED ? zfirst + 2 : src_plane(p, ic.z0, 2)) * plane + tile_off + seam_off;

    const int nk = ic.nz + 4;   // local planes 0 .. nz+3 arrive in order; output plane k-2 is produced when plane k lands
    int zi = zfirst;            // interior index of the output plane
    for (int k = 0; k < nk; ++k) {
      if (k < 4) {
        mbar_wait(&c.full[c.s], c.parity);
        if (k == 3) {
          ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
          ldg_f2_if(c.is_seam, seam_ptr + p.src_field, seam_next[1]);
        }
      } else {
        seam_ptr += zstep;
        int64_t inj_row = -1;
        if (inj_ly >= 0 && zi % x.inj.s == 0) inj_row = (int64_t(zi / x.inj.s) * x.inj.lh + inj_ly) * x.inj.lw;
        adjoint_plane<false, DOWN>(c, TP, k <= ic.nz + 2, seam_ptr, field, zstep, off, p.dst, nullptr, x.h, x.gadd,
                                   k + 1 < nk, valid, seam_next, aacc, macc, x.inj, inj_row, ic.x0 + 4 * lane);
        off += zstep;
        if (FUSED && k == 5 && own_early) slab_signal_inline(1, sync_mode, nbar);   // first boundary pair stored: over to the helper
        zi += DOWN ? -1 : 1;
        if (++since_flush >= BWD_FLUSH) flush();
      }
      advance_stage(c);
    }
    // the last four planes of the item are still held: hand their stages back (their loads fed the outputs that
    // were already stored, so they have completed)
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int j = 1; j <= 4; ++j) mbar_arrive(&c.empty[(c.s + STAGES - j) & (STAGES - 1)]);
    }
    if (FUSED && own_late) slab_signal_inline(2, sync_mode, nbar);
  }
  flush();
  // ---- CTA result -> global partials; the last CTA folds all CTAs in fixed order ----
  // The fold sits in the kernel's tail, where nothing overlaps it (the next step's kernel needs this one's complete
  // output): all consumer warps of the last CTA share it, one value per warp at a time, lanes striding over the CTAs
  // and a fixed shuffle tree (fold_partials) -- about five L2 loads deep instead of a 74-deep chain in one warp,
  // which cost ~9 us per step (8 % of a 64-plane slab step).
  asm volatile("bar.sync 3, %0;" ::"r"(p.ty * 32) : "memory");   // (ids 1, 2: slab helper hand-over)
  __shared__ bool s_last;
  constexpr int NR = kRedPiK1;
  if (warp == 0) {
    if (lane < NR) {
      double s = 0;
      for (int w = 0; w < p.ty; ++w) s += wacc[w * kRedPiK1 + lane];
      x.partials[size_t(blockIdx.x) * NR + lane] = s;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) s_last = (atomicAdd(x.counter, 1u) == gridDim.x - 1);
  }
  asm volatile("bar.sync 3, %0;" ::"r"(p.ty * 32) : "memory");
  if (s_last) {
    __threadfence();
    for (int i = warp; i < NR; i += p.ty) {
      double s = 0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(x.partials + size_t(b) * NR + i);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
      if (lane == 0) x.acc[i] += s;
    }
    if (warp == 0 && lane == 0) *x.counter = 0;
  }
}

}  // namespace tma3d
}  // namespace percnn
