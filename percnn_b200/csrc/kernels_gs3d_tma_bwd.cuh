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
// Reductions: per-lane fp32 partial sums, flushed every 32 planes into per-warp fp64 accumulators in shared
// memory; per-CTA results go to global memory and the last CTA folds them in fixed order (deterministic).
#pragma once
#include "kernels_gs3d_tma.cuh"

namespace percnn {
namespace tma3d {

constexpr int BWD_FLUSH = 32;
// The adjoint holds 22 running sums and the state on top of the 5-plane window: 15 consumer warps + 1 producer
// warp = 512 threads = 128 registers per thread (a 17th warp would round the allocation down to 96).
constexpr int BWD_WARPS = 15;
constexpr int BWD_THREADS = (BWD_WARPS + 1) * 32;
constexpr int SMEM_BYTES_BWD = SMEM_BYTES + 16 * kRedPiK1 * 8 + 64 + 10 * BWD_THREADS * 8;

struct BwdExtra {
  const float* h;        // stored state of this step, same layout as the G buffers
  const float* gadd;     // injected gradient for this step (nullable)
  double* partials;      // [gridDim.x][2 or 22]
  unsigned* counter;
  double* acc;           // [22] running sums over steps ([0..1] from this kernel, [2..21] from k_monomial_sums)
  Inject<float> inj;     // fused data-loss gradient of this step's state (target == nullptr: none)
};

__device__ __forceinline__ float2 quad2(const float* __restrict__ d, float2 u, float2 v) {
  float2 a0 = fma2(u, fma2(u, d[3], d[1]), d[0]);
  float2 a1 = fma2(u, d[4], d[2]);
  return fma2(v, fma2(v, d[5], a1), a0);
}
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// One output plane of the adjoint.  Unlike the forward kernel there is no register window: the five plane
// centres (z-2..z+2) are all read from the ring (planes k-4..k stay resident), which frees 32 registers for the
// state, the injected gradient and the pointwise Jacobian -- the windowed version spilled.
// `valid`: this warp's row is not a duplicate of the previous tile's rows (last tile of a column is shifted
// back), so it contributes to the reductions.
template <bool FUSED, bool MONO>
__device__ __forceinline__ void adjoint_plane(Consumer& c, const float* __restrict__ TP, bool prefetch_seam,
                                              const float* seam_ptr, int64_t field, int64_t plane, int64_t off,
                                              float* __restrict__ dst, float* mirror, const float* __restrict__ hbase,
                                              const float* __restrict__ gadd, bool prefetch_next, bool valid,
                                              float2 (&seam_next)[2], float (&aacc)[2], float2* __restrict__ macc,
                                              const Inject<float>& inj, int64_t inj_row, int xq,
                                              const float* __restrict__ hsm = nullptr, int hsm_field = 0) {
  const float* P = c.P;
  mbar_wait(&c.full[c.s], c.parity);   // plane k has landed; planes k-4 .. k-1 are still resident
  const float2 seam_u = seam_next[0], seam_v = seam_next[1];
  if (prefetch_seam) {
    ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
    ldg_f2_if(c.is_seam, seam_ptr + field, seam_next[1]);
  }
  // stored state of this step: from global memory (L2-prefetched one plane ahead), or -- warp-specialised kernel --
  // from the shared-memory ring the producer fills with TMA (hsm = this lane's quad in the current h stage)
  const float4 hu = hsm != nullptr ? lds128(hsm) : ldg128(hbase + off);
  const float4 hv = hsm != nullptr ? lds128(hsm + hsm_field) : ldg128(hbase + off + field);
  if (prefetch_next && (c.lane & 7) == 0) {
    if (hsm == nullptr) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off + plane));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off + plane + field));
    }
    if (gadd != nullptr) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gadd + off + plane));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gadd + off + plane + field));
    }
  }
  const uint32_t lane_off = (c.row + 2) * TX + 4 * c.lane;
  float2 Lu_lo, Lu_hi, Lv_lo, Lv_hi;
  float4 Gu, Gv;
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    float4 win[5];
#pragma unroll
    for (int j = 0; j < 5; ++j)   // plane k-4+j sits in stage (s + STAGES - 4 + j) % STAGES
      win[j] = lds128(c.ring + ((c.s + STAGES - 4 + j) & (STAGES - 1)) * STAGE_FLOATS + f * ROWS * TX + lane_off);
    const float* sp = c.ring + ((c.s + STAGES - 2) & (STAGES - 1)) * STAGE_FLOATS + f * ROWS * TX + c.row * TX + 4 * c.lane;
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
  // (in the warp-specialised kernel the h stage read above is recycled together with this G stage, so its loads join
  // the dependency; with h in global memory the release must NOT wait for those long-latency loads)
  float release_dep = Lu_lo.x + Lv_lo.x;
  if (hsm != nullptr) release_dep += hu.w + hv.w;
  if (c.lane == 0) mbar_arrive_after(&c.empty[(c.s + STAGES - 4) & (STAGES - 1)], release_dep);
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
  if (MONO && valid) {
    // 20 monomial sums  sum G_f u^a v^b  for this lane's 4 cells, added to per-lane accumulators in shared memory
    // (keeping them in registers next to the stencil state spilled).  Everything stays in the NATURAL register pairs
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
    macc[M * BWD_THREADS] = __fadd2_rn(macc[M * BWD_THREADS], make_float2(tu.x + tu.y, tv.x + tv.y)); \
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

template <int SLOT, bool FUSED, bool MONO>
__global__ void __launch_bounds__(BWD_THREADS, 1)
k_gs3d_bwd_tma(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
               const __grid_constant__ Params p, const __grid_constant__ BwdExtra x) {
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
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float2* macc_all = reinterpret_cast<float2*>(wacc + 16 * kRedPiK1);   // [10][BWD_THREADS] per-lane monomial sums
  for (int i = threadIdx.x; i < TY * kRedPiK1; i += BWD_THREADS) wacc[i] = 0.0;
  if (MONO)
    for (int m = 0; m < 10; ++m) macc_all[m * BWD_THREADS + threadIdx.x] = make_float2(0.f, 0.f);
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see the forward kernel
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int nitems = total_items(p);

  if (warp >= BWD_WARPS) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_main)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      uint32_t it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const ItemCoord ic = decode_item(p, item);
        if (FUSED && ic.seg < 2) {
          wait_flag(p.my_flags + ic.seg, p.epoch_wait, p.scratch + 1);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        int yh[4] = {ic.y0 - 2, ic.y0 - 1, ic.y0 + p.ty, ic.y0 + p.ty + 1};
#pragma unroll
        for (int h = 0; h < 4; ++h) yh[h] = yh[h] < 0 ? yh[h] + p.H : (yh[h] >= p.H ? yh[h] - p.H : yh[h]);
        const uint32_t bytes_main = 2u * uint32_t(p.ty) * TX * 4u, bytes_halo = 2u * 4u * TX * 4u;
        for (int k = 0; k < ic.nz + 4; ++k, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          const bool with_halo = (k >= 2) && (k < ic.nz + 2);
          const int pz = src_plane(p, ic.z0, k);
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
  // lap_quad indexes its table as P[P_LAP_C0], P[P_LAP_AX + i]; the mirrored taps sit at P_LAPT in the same order
  const float* TP = c.P + (P_LAPT - P_LAP_C0);
  const int64_t plane = int64_t(p.H) * p.W;
  const int64_t field = p.dst_field;
  float2 seam_next[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  float aacc[2] = {0.f, 0.f};
  int since_flush = 0;
  float2* macc = macc_all + threadIdx.x;
  auto flush = [&]() {   // per-lane fp32 partial sums -> per-warp fp64 accumulators (every BWD_FLUSH planes)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float t = aacc[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
      if (lane == 0) wacc[warp * kRedPiK1 + i] += double(t);
      aacc[i] = 0.f;
    }
    if (MONO) {
      const float dt = c.P[P_DT];
      for (int m = 0; m < 10; ++m) {
        float2 t = macc[m * BWD_THREADS];
        macc[m * BWD_THREADS] = make_float2(0.f, 0.f);
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
    }
    since_flush = 0;
  };
  bool posted = !FUSED;
  auto post_boundary_done = [&]() {
    // Every storing warp fences at system scope itself: its peer (NVLink) stores must be performed before the
    // flag can be observed.  Relying on one thread's fence after the CTA barrier to cover the other warps'
    // in-flight peer stores produced stale ghost planes on a neighbour (caught by the 2-GPU bitwise test).
    __threadfence_system();
    asm volatile("bar.sync 1, %0;" ::"r"(p.ty * 32) : "memory");
    if (warp == 0 && lane == 0) {
      __threadfence_system();
      const unsigned old = atomicAdd(p.scratch, 1u);
      if (old == gridDim.x - 1) {
        atomicExch(p.scratch, 0u);
        __threadfence_system();
        st_release_sys(p.post_lo_flag, p.epoch_post);
        st_release_sys(p.post_hi_flag, p.epoch_post);
      }
    }
  };
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    if (FUSED && !posted && ic.seg == 2) {
      post_boundary_done();
      posted = true;
    }
    float* mirror = nullptr;
    if (FUSED && ic.seg < 2) {
      float* base = ic.seg == 0 ? p.peer_lo_dst : p.peer_hi_dst;
      const int mz = ic.seg == 0 ? p.D + 2 + ic.z0 : ic.z0 - (p.D - 2);
      mirror = base + (int64_t(mz) * p.H + ic.y0) * p.W + ic.x0 + c.toff;
    }
    // rows below the natural start of this tile are duplicates of the previous tile (last tile shifted back)
    const bool valid = (ic.y0 + warp) >= ic.ytile * p.ty;
    // fused data loss: is this warp's row on the sampling lattice, and where does it start in the low-res frame
    const int inj_ly = (x.inj.target != nullptr && (ic.y0 + warp) % x.inj.s == 0) ? (ic.y0 + warp) / x.inj.s : -1;
    const float* src_xy = p.src + int64_t(ic.y0) * p.W + ic.x0;
    int64_t off = (int64_t(ic.z0 + p.dst_zoff) * p.H + ic.y0) * p.W + ic.x0 + c.toff;   // this lane's quad, first output plane
    int xs = (lane == 0) ? ic.x0 - 2 : ic.x0 + TX;
    xs = xs < 0 ? xs + p.W : (xs >= p.W ? xs - p.W : xs);
    const int seam_off = warp * p.W + xs - ic.x0;
    // seam cells come from planes [z0, z0 + nz) of the item itself: no periodic wrap needed (see the forward kernel)
    const float* seam_ptr = src_xy + int64_t(src_plane(p, ic.z0, 2)) * plane + seam_off;

    const int nk = ic.nz + 4;   // local planes 0 .. nz+3 arrive in order; output plane k-2 is produced when plane k lands
    for (int k = 0; k < nk; ++k) {
      if (k < 4) {
        mbar_wait(&c.full[c.s], c.parity);
        if (k == 3) {
          ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
          ldg_f2_if(c.is_seam, seam_ptr + p.src_field, seam_next[1]);
        }
      } else {
        seam_ptr += plane;
        int64_t inj_row = -1;
        if (inj_ly >= 0) {
          const int zg = ic.z0 + k - 4;   // interior index of the output plane
          if (zg % x.inj.s == 0) inj_row = (int64_t(zg / x.inj.s) * x.inj.lh + inj_ly) * x.inj.lw;
        }
        adjoint_plane<FUSED, MONO>(c, TP, k <= ic.nz + 2, seam_ptr, field, plane, off, p.dst, mirror, x.h, x.gadd,
                                   k + 1 < nk, valid, seam_next, aacc, macc, x.inj, inj_row, ic.x0 + 4 * lane);
        off += plane;
        if (FUSED && mirror != nullptr) mirror += plane;
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
  }
  if (FUSED && !posted) post_boundary_done();
  flush();
  // ---- CTA result -> global partials; last CTA folds all CTAs in fixed order ----
  asm volatile("bar.sync 2, %0;" ::"r"(p.ty * 32) : "memory");
  __shared__ bool s_last;
  constexpr int NR = MONO ? kRedPiK1 : 2;
  if (warp == 0) {
    if (lane < NR) {
      double s = 0;
      for (int w = 0; w < p.ty; ++w) s += wacc[w * kRedPiK1 + lane];
      x.partials[size_t(blockIdx.x) * NR + lane] = s;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) s_last = (atomicAdd(x.counter, 1u) == gridDim.x - 1);
    __syncwarp();
    if (s_last) {
      __threadfence();
      if (lane < NR) {
        double s0 = 0, s1 = 0;   // two chains keep the loads in flight; fixed order
        unsigned b = 0;
        for (; b + 2 <= gridDim.x; b += 2) {
          s0 += __ldcg(x.partials + size_t(b) * NR + lane);
          s1 += __ldcg(x.partials + size_t(b + 1) * NR + lane);
        }
        if (b < gridDim.x) s0 += __ldcg(x.partials + size_t(b) * NR + lane);
        x.acc[lane] += s0 + s1;
      }
      if (lane == 0) *x.counter = 0;
    }
  }
}

// =====================================================================================================
// Warp-specialised variant ("mono warps").  ncu on k_gs3d_bwd_tma at 512^3 (profiles/r01_ncu_tma_bwd_512_details.txt):
// the shared-memory pipe is the busiest unit (70 %), and 36 % of its wavefronts are the read-modify-write of the
// per-lane monomial accumulators, which live in shared memory only because the stencil warps have no registers
// left for them.  Here the 20 monomial sums move to two dedicated warps that do nothing else: they read the centre
// rows of G from the ring and h from global memory and keep all 40 partial sums in registers.  The 14 stencil warps
// lose ~100 instructions and 20 shared-memory accesses per plane.  Layout (640 threads, as the forward kernel):
//   warps 0..13  stencil warps, one tile row each (adjoint_plane<FUSED, false>)
//   warps 14,15  monomial warps, set A: the six monomials of degree >= 2 except u^2 (uv v^2 u^3 u^2v uv^2 v^3), even / odd rows
//   warp  16     producer: one lane issues the TMA loads
//   warps 17,18  monomial warps, set B: 1 u v u^2 (few registers: they live in the producer warp-group), even / odd rows
//   warp  19     idle
// With only two monomial warps doing all ten monomials those warps were the critical path (profiles/
// r01_adjoint_mw_experiment.txt): 63 instructions per row against the stencil warps' 330 per plane.
// Every consumer warp -- stencil or monomial -- waits on the same full barriers and releases plane k-4 after
// iteration k, so the ring protocol is unchanged (empty barriers count ty + 4 arrivals).
// =====================================================================================================
constexpr int MW_STENCIL = 14;
constexpr int MW_MONO = 2;
constexpr int MW_CONSUMERS = MW_STENCIL + MW_MONO;
constexpr int MW_THREADS = MW_CONSUMERS * 32 + 128;
constexpr int MW_CONSUMER_REGS = 104;   // 512 * 104 + 128 * 48 <= 640 * 96 (setmaxnreg only redistributes the CTA's allocation)
constexpr int MW_PRODUCER_REGS = 48;    // producer warp-group: the TMA lane + two light monomial warps (set B)
constexpr int MW_WACC_ROWS = 20;        // one row of fp64 sums per warp of the CTA
constexpr int MW_FLUSH = 8;             // planes between fp32 -> fp64 flushes of the monomial warps (<= 224 terms per lane sum)
// The stored state h_t streams through its own small ring: plane (k - 4) of h travels with plane k of G (same full
// barrier), is read during consumer iteration k and recycled with G plane k - 4, i.e. after that same iteration.
// The producer reuses an h stage HSTAGES planes later, by which time it has waited for the G stage released in
// iteration k: 4 stages are exactly enough.
constexpr int HSTAGES = 4;
constexpr int HSTAGE_FLOATS = 2 * MW_STENCIL * TX;
constexpr int MW_OFF_HRING = STAGES * STAGE_BYTES;
constexpr int MW_OFF_BARS = MW_OFF_HRING + HSTAGES * HSTAGE_FLOATS * 4;
constexpr int MW_OFF_WACC = MW_OFF_BARS + 2 * STAGES * 8 + 64;
constexpr int SMEM_BYTES_BWD_MW = MW_OFF_WACC + MW_WACC_ROWS * kRedPiK1 * 8 + 64;
static_assert(SMEM_BYTES_BWD_MW <= 227 * 1024, "shared memory budget");

// Monomial warp: sums  sum dt G_f u^a v^b  for the monomials of SET over the tile rows row0, row0 + 2, ... of every
// plane.  SET 0 ("A"): uv v^2 u^3 u^2v uv^2 v^3 (reduction slots 4..9); SET 1 ("B"): 1 u v u^2 (slots 0..3).
// All partial sums stay in registers: (cells 0+2, cells 1+3) pairs per monomial and field, folded every MW_FLUSH planes.
template <int SET>
__device__ __forceinline__ void mono_warp_loop(Consumer& c, const Params& p, const float* __restrict__ hring,
                                               double* __restrict__ wacc, int nitems, int row0, int warp, int lane) {
  constexpr int NM = SET == 0 ? 6 : 4;
  constexpr int M0 = SET == 0 ? 4 : 0;
  float2 au[NM], av[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) au[m] = av[m] = make_float2(0.f, 0.f);
  int since_flush = 0;
  uint32_t hcur = 0;
  auto flush = [&]() {
    const float dt = c.P[P_DT];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      float tu = au[m].x + au[m].y, tv = av[m].x + av[m].y;
      au[m] = av[m] = make_float2(0.f, 0.f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tu += __shfl_down_sync(0xffffffffu, tu, o);
        tv += __shfl_down_sync(0xffffffffu, tv, o);
      }
      if (lane == 0) {
        wacc[warp * kRedPiK1 + 2 + M0 + m] += double(dt) * double(tu);
        wacc[warp * kRedPiK1 + 12 + M0 + m] += double(dt) * double(tv);
      }
    }
    since_flush = 0;
  };
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    // first row of this warp that is not a duplicate of the previous tile (last tile of a column is shifted back)
    int r0 = row0;
    while (r0 < p.ty && (ic.y0 + r0) < ic.ytile * p.ty) r0 += MW_MONO;
    if (p.debug & 8) r0 = p.ty;   // timing experiment without the monomial work
    const int nk = ic.nz + 4;
    for (int k = 0; k < nk; ++k) {
      mbar_wait(&c.full[c.s], c.parity);
      if (k >= 4) {
        const float* st = c.ring + ((c.s + STAGES - 2) & (STAGES - 1)) * STAGE_FLOATS + 2 * TX + 4 * lane;   // plane k-2, tile row 0
        const float* hst = hring + hcur * HSTAGE_FLOATS + 4 * lane;   // state plane of this output plane, tile row 0
        float keep = 0.f;
        for (int r = r0; r < p.ty; r += MW_MONO) {
          const float4 Gu = lds128(st + r * TX), Gv = lds128(st + ROWS * TX + r * TX);
          const float4 hu = lds128(hst + r * TX), hv = lds128(hst + MW_STENCIL * TX + r * TX);
          keep = Gv.w + hv.w;
          const float2 ul = lo(hu), uh = hi(hu), vl = lo(hv), vh = hi(hv);
          const float2 gul = lo(Gu), guh = hi(Gu), gvl = lo(Gv), gvh = hi(Gv);
          const float2 uul = __fmul2_rn(ul, ul), uuh = __fmul2_rn(uh, uh);
#define PERCNN_MW_MONO(M, EL, EH)                                    \
  {                                                                  \
    const float2 el = EL, eh = EH;                                   \
    au[M] = fma2(guh, eh, fma2(gul, el, au[M]));                     \
    av[M] = fma2(gvh, eh, fma2(gvl, el, av[M]));                     \
  }
          if (SET == 1) {
            au[0] = __fadd2_rn(au[0], __fadd2_rn(gul, guh));
            av[0] = __fadd2_rn(av[0], __fadd2_rn(gvl, gvh));
            PERCNN_MW_MONO(1, ul, uh)
            PERCNN_MW_MONO(2, vl, vh)
            PERCNN_MW_MONO(3, uul, uuh)
          } else {
            const float2 uvl = __fmul2_rn(ul, vl), uvh = __fmul2_rn(uh, vh);
            const float2 vvl = __fmul2_rn(vl, vl), vvh = __fmul2_rn(vh, vh);
            PERCNN_MW_MONO(0, uvl, uvh)
            PERCNN_MW_MONO(1, vvl, vvh)
            PERCNN_MW_MONO(2, __fmul2_rn(uul, ul), __fmul2_rn(uuh, uh))
            PERCNN_MW_MONO(3, __fmul2_rn(uul, vl), __fmul2_rn(uuh, vh))
            PERCNN_MW_MONO(4, __fmul2_rn(ul, vvl), __fmul2_rn(uh, vvh))
            PERCNN_MW_MONO(5, __fmul2_rn(vvl, vl), __fmul2_rn(vvh, vh))
          }
#undef PERCNN_MW_MONO
        }
        // release plane k-4 like the stencil warps do (this warp has finished with every plane <= k-2); the
        // data dependency on the last shared-memory loads keeps the arrive behind them (see mbar_arrive_after)
        __syncwarp();
        if (lane == 0) mbar_arrive_after(&c.empty[(c.s + STAGES - 4) & (STAGES - 1)], keep);
        hcur = (hcur + 1) & (HSTAGES - 1);
        if (++since_flush >= MW_FLUSH) flush();
      }
      advance_stage(c);
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int j = 1; j <= 4; ++j) mbar_arrive(&c.empty[(c.s + STAGES - j) & (STAGES - 1)]);
    }
  }
  flush();
}

template <int SLOT, bool FUSED>
__global__ void __launch_bounds__(MW_THREADS, 1)
k_gs3d_bwd_tma_mw(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
                  const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ Params p,
                  const __grid_constant__ BwdExtra x) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  float* hring = reinterpret_cast<float*>(smem_raw + MW_OFF_HRING);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + MW_OFF_BARS);
  uint64_t* empty = full + STAGES;
  double* wacc = reinterpret_cast<double*>(smem_raw + MW_OFF_WACC);   // [16][22]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nmono = 2 * MW_MONO;   // monomial warps: two of set A (consumer warp-groups) + two of set B (producer warp-group)
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], p.ty + nmono);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < MW_WACC_ROWS * kRedPiK1; i += MW_THREADS) wacc[i] = 0.0;
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int nitems = total_items(p);

  const int nsync = (p.ty + nmono) * 32;   // consumer threads that reach the final reduction
  if (warp >= MW_CONSUMERS) {
    // ===== producer warp-group: TMA lane (warp 16), light monomial warps (17, 18) =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MW_PRODUCER_REGS));
    if (warp == MW_CONSUMERS + 1 || warp == MW_CONSUMERS + 2) {
      Consumer c;
      c.P = c_prep[SLOT].f;
      c.ring = ring;
      c.full = full;
      c.empty = empty;
      c.s = 0;
      c.parity = 0;
      c.row = 0;
      c.lane = lane;
      c.toff = 0;
      c.is_seam = false;
      mono_warp_loop<1>(c, p, hring, wacc, nitems, warp - (MW_CONSUMERS + 1), warp, lane);   // set B: even / odd rows
      asm volatile("bar.sync 2, %0;" ::"r"(nsync) : "memory");   // joins the consumers' final reduction barrier
      return;
    }
    if (warp == MW_CONSUMERS && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_main)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_h)) : "memory");
      uint32_t it = 0, hit = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const ItemCoord ic = decode_item(p, item);
        if (FUSED && ic.seg < 2) {
          wait_flag(p.my_flags + ic.seg, p.epoch_wait, p.scratch + 1);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        int yh[4] = {ic.y0 - 2, ic.y0 - 1, ic.y0 + p.ty, ic.y0 + p.ty + 1};
#pragma unroll
        for (int h = 0; h < 4; ++h) yh[h] = yh[h] < 0 ? yh[h] + p.H : (yh[h] >= p.H ? yh[h] - p.H : yh[h]);
        const uint32_t bytes_main = 2u * uint32_t(p.ty) * TX * 4u, bytes_halo = 2u * 4u * TX * 4u;
        for (int k = 0; k < ic.nz + 4; ++k, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          const bool with_halo = (k >= 2) && (k < ic.nz + 2);
          const bool with_h = k >= 4;                  // state plane of output plane k - 4 travels with G plane k
          const int pz = src_plane(p, ic.z0, k);
          float* st = ring + s * STAGE_FLOATS;
          mbar_expect_tx(&full[s], bytes_main + (with_halo ? bytes_halo : 0u) + (with_h ? bytes_main : 0u));
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
          if (with_h) {
            float* hs = hring + (hit % HSTAGES) * HSTAGE_FLOATS;
            const int hz = ic.z0 + (k - 4) + p.dst_zoff;   // the state buffers share the layout of the gradient buffers
            tma_load_4d(hs, &tm_h, &full[s], ic.x0, ic.y0, hz, 0);
            tma_load_4d(hs + MW_STENCIL * TX, &tm_h, &full[s], ic.x0, ic.y0, hz, 1);
            ++hit;
          }
        }
      }
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MW_CONSUMER_REGS));
  uint32_t hcur = 0;   // h stage of the output plane of the current iteration (advances with every k >= 4)
  const bool is_mono = warp >= MW_STENCIL;
  if (!is_mono && warp >= p.ty) return;   // tile shorter than 14 rows (after the aligned setmaxnreg)
  const int64_t plane = int64_t(p.H) * p.W;
  const int64_t field = p.dst_field;
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

  if (is_mono) {
    mono_warp_loop<0>(c, p, hring, wacc, nitems, warp - MW_STENCIL, warp, lane);   // set A: even / odd rows
  } else {
    // ===== stencil warps =====
    const float* TP = c.P + (P_LAPT - P_LAP_C0);
    float2 seam_next[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float aacc[2] = {0.f, 0.f};
    int since_flush = 0;
    auto flush = [&]() {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float t = aacc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) wacc[warp * kRedPiK1 + i] += double(t);
        aacc[i] = 0.f;
      }
      since_flush = 0;
    };
    bool posted = !FUSED;
    auto post_boundary_done = [&]() {
      __threadfence_system();   // every storing warp fences its own peer stores (see the forward kernel)
      asm volatile("bar.sync 1, %0;" ::"r"(p.ty * 32) : "memory");
      if (warp == 0 && lane == 0) {
        __threadfence_system();
        const unsigned old = atomicAdd(p.scratch, 1u);
        if (old == gridDim.x - 1) {
          atomicExch(p.scratch, 0u);
          __threadfence_system();
          st_release_sys(p.post_lo_flag, p.epoch_post);
          st_release_sys(p.post_hi_flag, p.epoch_post);
        }
      }
    };
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const ItemCoord ic = decode_item(p, item);
      if (FUSED && !posted && ic.seg == 2) {
        post_boundary_done();
        posted = true;
      }
      float* mirror = nullptr;
      if (FUSED && ic.seg < 2) {
        float* base = ic.seg == 0 ? p.peer_lo_dst : p.peer_hi_dst;
        const int mz = ic.seg == 0 ? p.D + 2 + ic.z0 : ic.z0 - (p.D - 2);
        mirror = base + (int64_t(mz) * p.H + ic.y0) * p.W + ic.x0 + c.toff;
      }
      const bool valid = (ic.y0 + warp) >= ic.ytile * p.ty;
      const int inj_ly = (x.inj.target != nullptr && (ic.y0 + warp) % x.inj.s == 0) ? (ic.y0 + warp) / x.inj.s : -1;
      const float* src_xy = p.src + int64_t(ic.y0) * p.W + ic.x0;
      int64_t off = (int64_t(ic.z0 + p.dst_zoff) * p.H + ic.y0) * p.W + ic.x0 + c.toff;
      int xs = (lane == 0) ? ic.x0 - 2 : ic.x0 + TX;
      xs = xs < 0 ? xs + p.W : (xs >= p.W ? xs - p.W : xs);
      const int seam_off = warp * p.W + xs - ic.x0;
      const float* seam_ptr = src_xy + int64_t(src_plane(p, ic.z0, 2)) * plane + seam_off;
      const int nk = ic.nz + 4;
      for (int k = 0; k < nk; ++k) {
        if (k < 4) {
          mbar_wait(&c.full[c.s], c.parity);
          if (k == 3) {
            ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
            ldg_f2_if(c.is_seam, seam_ptr + p.src_field, seam_next[1]);
          }
        } else {
          seam_ptr += plane;
          int64_t inj_row = -1;
          if (inj_ly >= 0) {
            const int zg = ic.z0 + k - 4;
            if (zg % x.inj.s == 0) inj_row = (int64_t(zg / x.inj.s) * x.inj.lh + inj_ly) * x.inj.lw;
          }
          adjoint_plane<FUSED, false>(c, TP, k <= ic.nz + 2, seam_ptr, field, plane, off, p.dst, mirror, x.h, x.gadd,
                                      k + 1 < nk, valid, seam_next, aacc, nullptr, x.inj, inj_row, ic.x0 + 4 * lane,
                                      hring + hcur * HSTAGE_FLOATS + warp * TX + 4 * lane, MW_STENCIL * TX);
          hcur = (hcur + 1) & (HSTAGES - 1);
          off += plane;
          if (FUSED && mirror != nullptr) mirror += plane;
          if (++since_flush >= BWD_FLUSH) flush();
        }
        advance_stage(c);
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int j = 1; j <= 4; ++j) mbar_arrive(&c.empty[(c.s + STAGES - j) & (STAGES - 1)]);
      }
    }
    if (FUSED && !posted) post_boundary_done();
    flush();
  }
  // ---- CTA result -> global partials; last CTA folds all CTAs in fixed order ----
  asm volatile("bar.sync 2, %0;" ::"r"(nsync) : "memory");
  __shared__ bool s_last;
  constexpr int NR = kRedPiK1;
  if (warp == 0) {
    if (lane < NR) {
      double s = 0;
      for (int w = 0; w < MW_WACC_ROWS; ++w) s += wacc[w * kRedPiK1 + lane];
      x.partials[size_t(blockIdx.x) * NR + lane] = s;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) s_last = (atomicAdd(x.counter, 1u) == gridDim.x - 1);
    __syncwarp();
    if (s_last) {
      __threadfence();
      if (lane < NR) {
        double s0 = 0, s1 = 0;
        unsigned b = 0;
        for (; b + 2 <= gridDim.x; b += 2) {
          s0 += __ldcg(x.partials + size_t(b) * NR + lane);
          s1 += __ldcg(x.partials + size_t(b + 1) * NR + lane);
        }
        if (b < gridDim.x) s0 += __ldcg(x.partials + size_t(b) * NR + lane);
        x.acc[lane] += s0 + s1;
      }
      if (lane == 0) *x.counter = 0;
    }
  }
}

// The 20 monomial sums  sum_x dt G_f u^a v^b  (-> gradient of the folded cubic) need no stencil: a plain
// streaming pass over h and G (16 B/cell) with 128-bit loads.  Kept out of the stencil kernel so that the latter
// fits its register budget (its 5-plane window + 20 running sums spilled).
__global__ void __launch_bounds__(256) k_monomial_sums(const float* __restrict__ h, const float* __restrict__ g, int64_t field,
                                                       int64_t base, int64_t n4, float dt, double* __restrict__ partials,
                                                       unsigned* __restrict__ counter, double* __restrict__ acc) {
  float2 m[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = make_float2(0.f, 0.f);
  const float4* hu = reinterpret_cast<const float4*>(h + base);
  const float4* hv = reinterpret_cast<const float4*>(h + base + field);
  const float4* gu = reinterpret_cast<const float4*>(g + base);
  const float4* gv = reinterpret_cast<const float4*>(g + base + field);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 u4 = __ldg(hu + i), v4 = __ldg(hv + i), a4 = __ldg(gu + i), b4 = __ldg(gv + i);
    const float us[4] = {u4.x, u4.y, u4.z, u4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
    const float as[4] = {a4.x, a4.y, a4.z, a4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = us[j], v = vs[j];
      const float2 gg = make_float2(as[j], bs[j]);
      const float uu = u * u, uv = u * v, vv = v * v;
      m[0] = __fadd2_rn(m[0], gg);
      m[1] = fma2(gg, u, m[1]);
      m[2] = fma2(gg, v, m[2]);
      m[3] = fma2(gg, uu, m[3]);
      m[4] = fma2(gg, uv, m[4]);
      m[5] = fma2(gg, vv, m[5]);
      m[6] = fma2(gg, uu * u, m[6]);
      m[7] = fma2(gg, uu * v, m[7]);
      m[8] = fma2(gg, u * vv, m[8]);
      m[9] = fma2(gg, vv * v, m[9]);
    }
  }
  __shared__ double sm[8][20];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double a = double(m[i].x), b = double(m[i].y);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, off);
      b += __shfl_down_sync(0xffffffffu, b, off);
    }
    if (lane == 0) {
      sm[warp][i] = a;        // field u sums
      sm[warp][10 + i] = b;   // field v sums
    }
  }
  __syncthreads();
  if (threadIdx.x < 20) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    partials[size_t(blockIdx.x) * 20 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    fold_partials(partials, gridDim.x, 20, acc, double(dt), 2);
    if (threadIdx.x == 0) *counter = 0;
  }
}

}  // namespace tma3d
}  // namespace percnn
