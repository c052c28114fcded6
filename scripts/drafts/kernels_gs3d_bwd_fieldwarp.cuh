// DRAFT -- NOT PART OF THE LIBRARY BUILD, NEVER RUN ON A GPU.  Written at the end of round 1 (no GPU minutes left) as
// the starting point for the next attempt on the 3-D adjoint; only its register footprint has been checked
// (scripts/drafts/probe_fieldwarp.cu + ptxas -v, numbers in profiles/r01_adjoint_mw_experiment.txt).
//
// "One field per warp": warp w of a CTA owns tile row (w >> 1) and field (w & 1).  Compared with k_gs3d_bwd_tma:
//   * the z-neighbours of the own field live in a 5-plane REGISTER window (20 registers instead of 40 for both
//     fields), so a plane costs 1 + 4 (+1 for the other field's centre) LDS.128 instead of 18, and the ring only has
//     to keep 3 planes resident (5 planes of look-ahead instead of 3), exactly like the forward kernel;
//   * the 10 monomial sums of the own field stay in registers (10 float2, natural pairs), no shared-memory
//     accumulators;
//   * tiles are 128 x 7 rows (14 consumer warps): at 512^2 that is 296 columns = two lock-step rounds on 148 SMs, and
//     4 halo rows per 7 tile rows (shared-memory fill 1.57x the tile instead of 1.29x; DRAM traffic unchanged).
// Ring protocol, producer, seam handling and the fused-halo hooks are the forward kernel's; the arithmetic is
// adjoint_plane's, restricted to one output field.
#pragma once
#include "../../percnn_b200/csrc/kernels_gs3d_tma_bwd.cuh"

namespace percnn {
namespace tma3d {

constexpr int FW_ROWS_MAX = 7;                       // tile rows; consumer warps = 2 * rows
constexpr int FW_CONSUMERS = 2 * FW_ROWS_MAX;
constexpr int FW_THREADS = (FW_CONSUMERS + 2) * 32;  // + producer warp + one idle warp: 16 warps -> 128 registers
constexpr int FW_SMEM_BYTES = SMEM_BYTES + 16 * kRedPiK1 * 8 + 64;

// Per-warp running sums.
struct FwSums {
  float aacc;        // sum q_f * dt * Lap^T(G_f)
  float m[10];       // sum G_f * monomial_j (folded); dt is applied at the flush
};

// Laplacian^T of ONE field for the lane's 4 cells; identical arithmetic to lap_quad.
template <int R>
__device__ __forceinline__ void fw_lap(const float* __restrict__ TP, const float4 (&w)[5], const float4 (&y)[4], float Lz,
                                       float Lw, float Rx, float Ry, float2& lo_out, float2& hi_out) {
  const float4 wl[5] = {w[(R + 0) % 5], w[(R + 1) % 5], w[(R + 2) % 5], w[(R + 3) % 5], w[(R + 4) % 5]};
  lap_quad(TP, wl, y, Lz, Lw, Rx, Ry, lo_out, hi_out);
}

// One output plane (local plane k arrives, output plane k-2 is produced) for field F of this warp's row.
template <int R, int F, bool FUSED>
__device__ __forceinline__ void fw_plane(Consumer& c, const float* __restrict__ TP, bool drain, bool prefetch_seam,
                                         const float* seam_ptr, int64_t field, int64_t plane, int64_t off,
                                         float* __restrict__ dst, float* mirror, const float* __restrict__ hbase,
                                         const float* __restrict__ gadd, bool prefetch_next, bool valid, float4 (&w)[5],
                                         float2& seam_next, FwSums& sums, const Inject<float>& inj, int64_t inj_row,
                                         int xq) {
  const float* P = c.P;
  mbar_wait(&c.full[c.s], c.parity);
  {
    const float* st = c.ring + c.s * STAGE_FLOATS + F * ROWS * TX + (c.row + 2) * TX + 4 * c.lane;
    w[(R + 4) % 5] = lds128(st);
  }
  if (drain) {   // planes past the chunk end are z-neighbours only
    __syncwarp();
    if (c.lane == 0) mbar_arrive_after(&c.empty[c.s], w[(R + 4) % 5].w);
  }
  const float2 seam = seam_next;
  if (prefetch_seam) ldg_f2_if(c.is_seam, seam_ptr, seam_next);
  // stored state (both fields) of the output cell quad, injected gradient of the own field
  const float4 hu = ldg128(hbase + off), hv = ldg128(hbase + off + field);
  if (prefetch_next && (c.lane & 7) == 0) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off + plane));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off + plane + field));
    if (gadd != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(gadd + off + F * field + plane));
  }
  const uint32_t s2 = (c.s + STAGES - 2) & (STAGES - 1);
  const float* sp = c.ring + s2 * STAGE_FLOATS + F * ROWS * TX + c.row * TX + 4 * c.lane;
  const float4 Gf = w[(R + 2) % 5];
  float2 L_lo, L_hi;
  float4 Gg;   // centre of the OTHER field in the output plane
  {
    const float4 y[4] = {lds128(sp), lds128(sp + TX), lds128(sp + 3 * TX), lds128(sp + 4 * TX)};
    Gg = lds128(c.ring + s2 * STAGE_FLOATS + (1 - F) * ROWS * TX + (c.row + 2) * TX + 4 * c.lane);
    float Lz = __shfl_up_sync(0xffffffffu, Gf.z, 1), Lw = __shfl_up_sync(0xffffffffu, Gf.w, 1);
    float Rx = __shfl_down_sync(0xffffffffu, Gf.x, 1), Ry = __shfl_down_sync(0xffffffffu, Gf.y, 1);
    if (c.lane == 0) { Lz = seam.x; Lw = seam.y; }
    if (c.lane == 31) { Rx = seam.x; Ry = seam.y; }
    fw_lap<R>(TP, w, y, Lz, Lw, Rx, Ry, L_lo, L_hi);
  }
  // the rows of plane k-2 are no longer needed by this warp
  __syncwarp();
  if (c.lane == 0) mbar_arrive_after(&c.empty[s2], L_lo.x + Gg.w);
  const float4 Gu = F == 0 ? Gf : Gg, Gv = F == 0 ? Gg : Gf;
  float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gadd != nullptr) a4 = ldg128(gadd + off + F * field);
  const float alpha = P[P_ALPHA + F], dt = P[P_DT];
  const float* D = P + P_DPOLY;
  float4 o;
#define PERCNN_FW_PAIR(U2, V2, GU2, GV2, GF2, L2, O0, O1, A0, A1)                                              \
  {                                                                                                          \
    const float2 gdu = mul2(GU2, dt), gdv = mul2(GV2, dt);                                                   \
    const float2 s = fma2(gdu, quad2(D + 6 * F, U2, V2), __fmul2_rn(gdv, quad2(D + 12 + 6 * F, U2, V2)));    \
    const float2 l = mul2(L2, dt);                                                                           \
    float2 g = __fadd2_rn(GF2, fma2(l, alpha, s));                                                           \
    g = __fadd2_rn(g, make_float2(A0, A1));                                                                  \
    O0 = g.x; O1 = g.y;                                                                                      \
    if (valid) {                                                                                             \
      const float2 q = F == 0 ? U2 : V2;                                                                     \
      sums.aacc = fmaf(q.x, l.x, fmaf(q.y, l.y, sums.aacc));                                                 \
    }                                                                                                        \
  }
  PERCNN_FW_PAIR(lo(hu), lo(hv), lo(Gu), lo(Gv), lo(Gf), L_lo, o.x, o.y, a4.x, a4.y)
  PERCNN_FW_PAIR(hi(hu), hi(hv), hi(Gu), hi(Gv), hi(Gf), L_hi, o.z, o.w, a4.z, a4.w)
#undef PERCNN_FW_PAIR
  if (valid) {   // 10 monomial sums of the own field, natural pairs, registers
    const float2 ul = lo(hu), uh = hi(hu), vl = lo(hv), vh = hi(hv);
    const float2 gl = lo(Gf), gh = hi(Gf);
    const float2 uul = __fmul2_rn(ul, ul), uuh = __fmul2_rn(uh, uh);
    const float2 uvl = __fmul2_rn(ul, vl), uvh = __fmul2_rn(uh, vh);
    const float2 vvl = __fmul2_rn(vl, vl), vvh = __fmul2_rn(vh, vh);
#define PERCNN_FW_MONO(M, EL, EH) { const float2 t = fma2(gh, EH, __fmul2_rn(gl, EL)); sums.m[M] += t.x + t.y; }
    { const float2 t = __fadd2_rn(gl, gh); sums.m[0] += t.x + t.y; }
    PERCNN_FW_MONO(1, ul, uh)
    PERCNN_FW_MONO(2, vl, vh)
    PERCNN_FW_MONO(3, uul, uuh)
    PERCNN_FW_MONO(4, uvl, uvh)
    PERCNN_FW_MONO(5, vvl, vvh)
    PERCNN_FW_MONO(6, __fmul2_rn(uul, ul), __fmul2_rn(uuh, uh))
    PERCNN_FW_MONO(7, __fmul2_rn(uul, vl), __fmul2_rn(uuh, vh))
    PERCNN_FW_MONO(8, __fmul2_rn(ul, vvl), __fmul2_rn(uh, vvh))
    PERCNN_FW_MONO(9, __fmul2_rn(vvl, vl), __fmul2_rn(vvh, vh))
#undef PERCNN_FW_MONO
  }
  if (inj_row >= 0) {   // fused data loss (warp-uniform branch, selected steps only)
    const float icoef = inject_coef(inj);
    const float4 hq = F == 0 ? hu : hv;
    const float hs[4] = {hq.x, hq.y, hq.z, hq.w};
    float io[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xg = xq + j;
      if (xg % inj.s == 0) io[j] = fmaf(icoef, hs[j] - __ldg(inj.target + F * inj.lfield + inj_row + xg / inj.s), io[j]);
    }
    o = make_float4(io[0], io[1], io[2], io[3]);
  }
  *reinterpret_cast<float4*>(dst + off + F * field) = o;
  if (FUSED && mirror != nullptr) *reinterpret_cast<float4*>(mirror + F * field) = o;
  advance_stage(c);
}

// Whole per-warp loop for field F (the kernel branches once on the warp's field so that every coefficient address
// is a compile-time constant).
template <int SLOT, int F, bool FUSED>
__device__ __forceinline__ void fw_consumer(const Params& p, const BwdExtra& x, float* ring, uint64_t* full, uint64_t* empty,
                                            double* wacc, int warp, int lane, int nitems) {
  Consumer c;
  c.P = c_prep[SLOT].f;
  c.ring = ring;
  c.full = full;
  c.empty = empty;
  c.s = 0;
  c.parity = 0;
  c.row = warp >> 1;
  c.lane = lane;
  c.toff = uint32_t(c.row) * uint32_t(p.W) + 4u * uint32_t(lane);
  c.is_seam = (lane == 0) || (lane == 31);
  const float* TP = c.P + (P_LAPT - P_LAP_C0);
  const int64_t plane = int64_t(p.H) * p.W;
  const int64_t field = p.dst_field;
  float4 w[5];
  float2 seam_next = make_float2(0.f, 0.f);
  FwSums sums;
  sums.aacc = 0.f;
#pragma unroll
  for (int m = 0; m < 10; ++m) sums.m[m] = 0.f;
  int since_flush = 0;
  auto flush = [&]() {
    const float dt = c.P[P_DT];
    float t = sums.aacc;
    sums.aacc = 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) wacc[warp * kRedPiK1 + F] += double(t);
#pragma unroll
    for (int m = 0; m < 10; ++m) {
      float tm = sums.m[m];
      sums.m[m] = 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tm += __shfl_down_sync(0xffffffffu, tm, o);
      if (lane == 0) wacc[warp * kRedPiK1 + 2 + 10 * F + m] += double(dt) * double(tm);
    }
    since_flush = 0;
  };
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    float* mirror = nullptr;
    if (FUSED && ic.seg < 2) {
      float* base = ic.seg == 0 ? p.peer_lo_dst : p.peer_hi_dst;
      const int mz = ic.seg == 0 ? p.D + 2 + ic.z0 : ic.z0 - (p.D - 2);
      mirror = base + (int64_t(mz) * p.H + ic.y0) * p.W + ic.x0 + c.toff;
    }
    const bool valid = (ic.y0 + c.row) >= ic.ytile * p.ty;
    const int inj_ly = (x.inj.target != nullptr && (ic.y0 + c.row) % x.inj.s == 0) ? (ic.y0 + c.row) / x.inj.s : -1;
    const float* src_xy = p.src + F * p.src_field + int64_t(ic.y0) * p.W + ic.x0;
    int64_t off = (int64_t(ic.z0 + p.dst_zoff) * p.H + ic.y0) * p.W + ic.x0 + c.toff;
    int xs = (lane == 0) ? ic.x0 - 2 : ic.x0 + TX;
    xs = xs < 0 ? xs + p.W : (xs >= p.W ? xs - p.W : xs);
    const int seam_off = c.row * p.W + xs - ic.x0;
    const float* seam_ptr = src_xy + int64_t(src_plane(p, ic.z0, 2)) * plane + seam_off;
    // warm-up planes 0..3: only enter the register window (planes 0, 1 can be released at once)
#define PERCNN_FW_WARM(SLOTI, REL)                                                                            \
  {                                                                                                          \
    mbar_wait(&c.full[c.s], c.parity);                                                                       \
    w[SLOTI] = lds128(c.ring + c.s * STAGE_FLOATS + F * ROWS * TX + (c.row + 2) * TX + 4 * lane);            \
    if (REL) {                                                                                               \
      __syncwarp();                                                                                          \
      if (lane == 0) mbar_arrive_after(&c.empty[c.s], w[SLOTI].w);                                           \
    }                                                                                                        \
    advance_stage(c);                                                                                        \
  }
    PERCNN_FW_WARM(0, true)
    PERCNN_FW_WARM(1, true)
    PERCNN_FW_WARM(2, false)
    PERCNN_FW_WARM(3, false)
#undef PERCNN_FW_WARM
    ldg_f2_if(c.is_seam, seam_ptr, seam_next);
    const int nk = ic.nz + 4;
    int k = 4;
#define PERCNN_FW_STEADY(RR)                                                                                     \
  {                                                                                                              \
    seam_ptr += plane;                                                                                           \
    int64_t inj_row = -1;                                                                                        \
    if (inj_ly >= 0) {                                                                                           \
      const int zg = ic.z0 + k - 4;                                                                              \
      if (zg % x.inj.s == 0) inj_row = (int64_t(zg / x.inj.s) * x.inj.lh + inj_ly) * x.inj.lw;                    \
    }                                                                                                            \
    fw_plane<RR, F, FUSED>(c, TP, k >= ic.nz + 2, k <= ic.nz + 2, seam_ptr, field, plane, off, p.dst, mirror, x.h,   \
                           x.gadd, k + 1 < nk, valid, w, seam_next, sums, x.inj, inj_row, ic.x0 + 4 * lane);     \
    off += plane;                                                                                                \
    if (FUSED && mirror != nullptr) mirror += plane;                                                             \
    if (++since_flush >= BWD_FLUSH) flush();                                                                     \
    ++k;                                                                                                         \
  }
    // window slots: warm-up filled w[0..3] with planes 0..3, so the first steady plane (k = 4) uses rotation R = 0
    // (new plane into slot (R + 4) % 5 = 4, centre = slot 2 = plane 2)
    while (k + 5 <= nk) {
      PERCNN_FW_STEADY(0) PERCNN_FW_STEADY(1) PERCNN_FW_STEADY(2) PERCNN_FW_STEADY(3) PERCNN_FW_STEADY(4)
    }
    if (k < nk) PERCNN_FW_STEADY(0)
    if (k < nk) PERCNN_FW_STEADY(1)
    if (k < nk) PERCNN_FW_STEADY(2)
    if (k < nk) PERCNN_FW_STEADY(3)
#undef PERCNN_FW_STEADY
    // NOTE for whoever finishes this: after a partial last group the window rotation of the NEXT item restarts at
    // R = 0 with freshly loaded planes, so no state carries over; the ring stages of the last two planes were
    // released by the drain path above (k >= nz + 2), planes nz .. nz+1 by the ordinary release of s2.
  }
  flush();
}

template <int SLOT, bool FUSED>
__global__ void __launch_bounds__(FW_THREADS, 1)
k_gs3d_bwd_tma_fw(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_halo,
                  const __grid_constant__ Params p, const __grid_constant__ BwdExtra x) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  double* wacc = reinterpret_cast<double*>(smem_raw + STAGES * STAGE_BYTES + 2 * STAGES * 8 + 64);   // [16][22]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2 * p.ty);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 16 * kRedPiK1; i += FW_THREADS) wacc[i] = 0.0;
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int nitems = total_items(p);
  if (warp >= FW_CONSUMERS) {
    if (warp == FW_CONSUMERS && lane == 0) {
      // producer: identical to k_gs3d_fwd_tma's (3 resident planes, halo rows as separate boxes)
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
  if ((warp >> 1) >= p.ty) return;
  // TODO(next round): the fused-halo "boundary done" publication (post_boundary_done) is not wired in this draft.
  if (warp & 1)
    fw_consumer<SLOT, 1, FUSED>(p, x, ring, full, empty, wacc, warp, lane, nitems);
  else
    fw_consumer<SLOT, 0, FUSED>(p, x, ring, full, empty, wacc, warp, lane, nitems);
  // ---- CTA result -> global partials; last CTA folds all CTAs in fixed order (as in k_gs3d_bwd_tma) ----
  asm volatile("bar.sync 2, %0;" ::"r"(2 * p.ty * 32) : "memory");
  __shared__ bool s_last;
  constexpr int NR = kRedPiK1;
  if (warp == 0) {
    if (lane < NR) {
      double s = 0;
      for (int wq = 0; wq < 16; ++wq) s += wacc[wq * kRedPiK1 + lane];
      x.partials[size_t(blockIdx.x) * NR + lane] = s;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) s_last = (atomicAdd(x.counter, 1u) == gridDim.x - 1);
    __syncwarp();
    if (s_last) {
      __threadfence();
      if (lane < NR) {
        double s0 = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) s0 += __ldcg(x.partials + size_t(b) * NR + lane);
        x.acc[lane] += s0;
      }
      if (lane == 0) *x.counter = 0;
    }
  }
}

}  // namespace tma3d
}  // namespace percnn
