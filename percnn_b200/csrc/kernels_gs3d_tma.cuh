// Flagship kernel: one fused time step of the 3-D Pi-block cell (k = 1) for large fp32 grids.
//
//   u+ = u + dt * (alpha_u * Lap(u) + R_u(u, v)),  v+ likewise      (GS3D:123-139)
//
// Design (B200 / sm_100a):
//  * persistent CTAs, one per SM; a work item is an (x-tile, y-tile, z-chunk) column that the CTA
//    marches along z ("2.5-D blocking");
//  * a dedicated producer warp streams xy-planes of both fields into a shared-memory ring with TMA
//    (cp.async.bulk.tensor.4d -> SASS UTMALDG), completion on mbarriers; periodic wrap in y/z is done
//    by issuing the halo rows / wrapped planes as separate TMA boxes, so the state stays un-padded;
//  * 16 consumer warps, one per tile row, 4 cells per lane (LDS.128 / STG.128).  The z-neighbours
//    live in a 5-plane register window, the y-neighbours come from the ring, the x-neighbours from
//    the adjacent lanes by warp shuffle; only lanes 0 and 31 (the tile seam) fetch their two halo
//    cells from global memory, one plane ahead;
//  * all arithmetic is packed FFMA2 (2 x fp32 per issue slot, coefficients as broadcast uniform
//    operands from __constant__), the 1x1 Pi-block is its folded bivariate cubic (9 FMA);
//  * every cell is read once from HBM (+ halo re-reads that hit L2) and written once.
#pragma once
#include "point_ops.cuh"

namespace percnn {


namespace tma3d {

constexpr int TX = 128;        // tile width  = one warp x 4 cells per lane
constexpr int TY = 16;         // MAXIMUM tile height = consumer warps; the actual height is Params::ty (<= 15,
                               // so that the adjoint kernel, which runs 15 consumer warps, can share the tiling)
constexpr int STAGES = 8;      // planes in the ring
constexpr int ROWS = TY + 4;   // rows per field per stage (2 halo rows above and below)   // rows per field per stage (2 halo rows above and below)
constexpr int STAGE_FLOATS = 2 * ROWS * TX;
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
constexpr int CONSUMER_THREADS = TY * 32;
// 16 consumer warps + one producer warp-group (4 warps, of which one lane works).  The kernel is launched at
// the register count 640 threads allow (96); setmaxnreg then moves registers from the idle producer
// warp-group to the consumers, whose 5-plane window wants ~100.
constexpr int THREADS = CONSUMER_THREADS + 128;
constexpr int CONSUMER_REGS = 112;   // 512 * 112 + 128 * 24 <= 640 * 96: setmaxnreg only redistributes the CTA's own allocation
constexpr int PRODUCER_REGS = 24;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;

struct Params {
  const float* src;   // base of the source state (for the seam loads)
  float* dst;
  int D, H, W;        // interior extents
  int src_planes;     // planes per field in src (D, or D + 4 in slab mode)
  int64_t src_field;  // elements between fields in src
  int64_t dst_field;
  int src_zoff;       // plane index of interior plane z - 2 is (z + src_zoff) [wrapped if wrap_z]
  int dst_zoff;       // interior plane z of dst is stored at plane z + dst_zoff
  int wrap_z;         // 1: periodic along z inside this buffer; 0: ghost planes present
  int ty;             // rows per tile (<= TY); tiles start at min(j * ty, H - ty), so only the last may overlap
  int nxt, nyt;       // tiles along x, y
  // Work is a list of up to 3 z-segments, processed in order by every CTA.  A plain step has one segment
  // [0, D); the fused slab step has [0,2) (needs the lower ghosts, mirrored to the lower neighbour),
  // [D-2, D) (upper ghosts / upper neighbour) and the interior [2, D-2).
  int nseg;
  int seg_lo[3], seg_hi[3], seg_tz[3], seg_nzc[3];
  int slot;
  // ---- fused halo exchange over peer memory (slab mode; all null/0 otherwise) ----
  int fused;
  float* peer_lo_dst;        // lower neighbour's destination buffer (same layout as dst), mapped over NVLink
  float* peer_hi_dst;
  const uint32_t* my_flags;  // [0] epoch up to which my lower ghosts are valid, [1] same for the upper ghosts
  uint32_t* post_lo_flag;    // lower neighbour's flags[1] (I write its upper ghosts)
  uint32_t* post_hi_flag;    // upper neighbour's flags[0]
  uint32_t* scratch;         // [0] arrival counter (lower boundary), [1] error word (spin deadline exceeded),
                             // [2] arrival counter (upper boundary)
  uint32_t epoch_wait, epoch_post;
  int debug;                 // bisecting aid: 1 = no mirror stores, 2 = no flag wait, 4 = no flag post
  uint32_t spin_limit;       // flag-wait spins before the kernel gives up (error word + trap)
  float* peer_lo_src;        // the neighbours' mappings of the buffer that is `src` here (deferred boundary pair, see
  float* peer_hi_src;        // kernels_gs3d_slab.cuh)
  int flush_prev;            // 1: the previous step of this rollout deferred its last boundary pair to this kernel
  int defer_late;            // 1: leave the boundary pair produced last to the next step's kernel
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Release a ring stage only after the shared-memory loads of it have COMPLETED.  `last_loaded` must come from the
// last LDS issued on that stage: the arrive's address is made data-dependent on it (x * 0 cannot be folded for
// IEEE floats), and a warp issues in order, so the arrive cannot be issued while the load is still queued.
// Without this a load stuck behind peer (NVLink) stores in the LSU was overtaken by the arrive, the producer
// refilled the stage and the load returned the NEXT plane (found by the multi-GPU bitwise test).
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, float last_loaded) {
  uint32_t z;
  asm volatile("{\n\t.reg .f32 t;\n\tmul.rn.f32 t, %1, 0f00000000;\n\tcvt.rzi.u32.f32 %0, t;\n\t}" : "=r"(z) : "f"(last_loaded));
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + z) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- packed fp32x2 helpers (FFMA2 / FMUL2; scalar operands broadcast) -------------------------
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fma2(float2 a, float s, float2 c) { return __ffma2_rn(a, bc(s), c); }
__device__ __forceinline__ float2 fma2(float2 a, float s, float t) { return __ffma2_rn(a, bc(s), bc(t)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float t) { return __ffma2_rn(a, b, bc(t)); }
__device__ __forceinline__ float2 mul2(float2 a, float s) { return __fmul2_rn(a, bc(s)); }
__device__ __forceinline__ float2 lo(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi(const float4& v) { return make_float2(v.z, v.w); }

__device__ __forceinline__ float2 cubic2(const float* __restrict__ c, float2 u, float2 v) {
  float2 a0 = fma2(u, fma2(u, fma2(u, c[6], c[3]), c[1]), c[0]);
  float2 a1 = fma2(u, fma2(u, c[7], c[4]), c[2]);
  float2 a2 = fma2(u, c[8], c[5]);
  return fma2(v, fma2(v, fma2(v, c[9], a2), a1), a0);
}

// Laplacian of one field for the 4 cells of this lane.
//   win[0..4]: planes z-2..z+2 (centre win[2]);  y[0..3]: rows y-2,y-1,y+1,y+2 of plane z;
//   Lz,Lw / Rx,Ry: the two cells left / right of the lane's quad.
__device__ __forceinline__ void lap_quad(const float* __restrict__ P, const float4 (&win)[5], const float4 (&y)[4], float Lz,
                                         float Lw, float Rx, float Ry, float2& acc_lo, float2& acc_hi) {
  const float4& c = win[2];
  // two independent partial sums per register pair (z-part and in-plane part) halve the dependent-FMA chain
  float2 za_lo = mul2(lo(c), P[P_LAP_C0]);
  float2 za_hi = mul2(hi(c), P[P_LAP_C0]);
  za_lo = fma2(lo(win[0]), P[P_LAP_AX + 0], za_lo);
  za_hi = fma2(hi(win[0]), P[P_LAP_AX + 0], za_hi);
  za_lo = fma2(lo(win[1]), P[P_LAP_AX + 1], za_lo);
  za_hi = fma2(hi(win[1]), P[P_LAP_AX + 1], za_hi);
  za_lo = fma2(lo(win[3]), P[P_LAP_AX + 2], za_lo);
  za_hi = fma2(hi(win[3]), P[P_LAP_AX + 2], za_hi);
  za_lo = fma2(lo(win[4]), P[P_LAP_AX + 3], za_lo);
  za_hi = fma2(hi(win[4]), P[P_LAP_AX + 3], za_hi);
  // y taps (axis 1)
  float2 pa_lo = mul2(lo(y[0]), P[P_LAP_AX + 4]);
  float2 pa_hi = mul2(hi(y[0]), P[P_LAP_AX + 4]);
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    pa_lo = fma2(lo(y[k]), P[P_LAP_AX + 4 + k], pa_lo);
    pa_hi = fma2(hi(y[k]), P[P_LAP_AX + 4 + k], pa_hi);
  }
  // x taps (axis 2): +-2 are register-pair aligned, +-1 straddle pairs -> scalar FFMA
  pa_lo = fma2(make_float2(Lz, Lw), P[P_LAP_AX + 8], pa_lo);
  pa_hi = fma2(lo(c), P[P_LAP_AX + 8], pa_hi);
  pa_lo = fma2(hi(c), P[P_LAP_AX + 11], pa_lo);
  pa_hi = fma2(make_float2(Rx, Ry), P[P_LAP_AX + 11], pa_hi);
  const float m1 = P[P_LAP_AX + 9], p1 = P[P_LAP_AX + 10];
  pa_lo.x = fmaf(p1, c.y, fmaf(m1, Lw, pa_lo.x));
  pa_lo.y = fmaf(p1, c.z, fmaf(m1, c.x, pa_lo.y));
  pa_hi.x = fmaf(p1, c.w, fmaf(m1, c.y, pa_hi.x));
  pa_hi.y = fmaf(p1, Rx, fmaf(m1, c.z, pa_hi.y));
  acc_lo = __fadd2_rn(za_lo, pa_lo);
  acc_hi = __fadd2_rn(za_hi, pa_hi);
}

__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }

struct ItemCoord {
  int x0, y0, z0, nz, seg, ytile;
};
__device__ __forceinline__ int total_items(const Params& p) {
  int n = 0;
  for (int s = 0; s < p.nseg; ++s) n += p.nxt * p.nyt * p.seg_nzc[s];
  return n;
}
__device__ __forceinline__ ItemCoord decode_item(const Params& p, int item) {
  ItemCoord c;
  int seg = 0;
  const int tiles = p.nxt * p.nyt;
  while (seg + 1 < p.nseg && item >= tiles * p.seg_nzc[seg]) {
    item -= tiles * p.seg_nzc[seg];
    ++seg;
  }
  const int xt = item % p.nxt;
  const int r = item / p.nxt;
  const int yt = r % p.nyt;
  const int zc = r / p.nyt;
  c.seg = seg;
  c.ytile = yt;
  c.x0 = xt * TX;
  c.y0 = min(yt * p.ty, p.H - p.ty);
  c.z0 = p.seg_lo[seg] + zc * p.seg_tz[seg];
  c.nz = min(p.seg_tz[seg], p.seg_hi[seg] - c.z0);
  return c;
}

// ---- cross-GPU flags (system scope) ----
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until *flag >= epoch (wrap-safe).  Bounded: a lost peer must not hang the GPU, but carrying on with stale
// ghost planes would silently corrupt the rollout, so the deadline is FATAL: the error word is set (for a
// debugger / the peer) and the kernel traps, which surfaces as a CUDA error at the caller's next synchronisation.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch, uint32_t* err, uint32_t spin_limit) {
  for (uint32_t spins = 0; spins < spin_limit; ++spins) {
    if (int32_t(ld_acquire_sys(flag) - epoch) >= 0) return;
    __nanosleep(64);
  }
  atomicExch(err, 1u);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ int src_plane(const Params& p, int z0, int j) {
  // z0 + j + src_zoff lies in [-2, D + 1]; the host only picks this kernel for D >= 4, so one
  // conditional wrap is enough (no integer division in the plane loop).
  int pz = z0 + j + p.src_zoff;
  if (p.wrap_z) pz = pz < 0 ? pz + p.D : (pz >= p.D ? pz - p.D : pz);
  return pz;
}

// Predicated 8-byte read-only load (no branch): only the two seam lanes of a warp touch global memory.
__device__ __forceinline__ void ldg_f2_if(bool pred, const float* ptr, float2& v) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t"
      "@p ld.global.nc.v2.f32 {%0, %1}, [%3];\n\t}"
      : "+f"(v.x), "+f"(v.y)
      : "r"(uint32_t(pred)), "l"(ptr));
}

// Per-warp consumer state that survives across planes and items.
struct Consumer {
  const float* P;        // coefficients (__constant__)
  float* ring;
  uint64_t* full;
  uint64_t* empty;
  uint32_t s;            // ring stage of the plane arriving next
  uint32_t parity;       // its mbarrier phase parity
  int row, lane;
  uint32_t toff;         // row * W + 4 * lane: this lane's quad inside a tile plane (elements)
  bool is_seam;          // lane 0 or 31
};

__device__ __forceinline__ void advance_stage(Consumer& c) {
  c.s = (c.s + 1) & (STAGES - 1);
  c.parity ^= (c.s == 0);
}

// Warm-up plane (local index 0..3 of an item): only enters the register window.
template <int R>
__device__ __forceinline__ void warm_plane(Consumer& c, bool release_now, float4 (&wu)[5], float4 (&wv)[5]) {
  mbar_wait(&c.full[c.s], c.parity);
  const float* st = c.ring + c.s * STAGE_FLOATS + (c.row + 2) * TX + 4 * c.lane;
  wu[(R + 4) % 5] = lds128(st);
  wv[(R + 4) % 5] = lds128(st + ROWS * TX);
  if (release_now) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive_after(&c.empty[c.s], wv[(R + 4) % 5].w);
  }
  advance_stage(c);
}

// Steady-state plane: local plane k arrives, output plane k-2 is produced.
//   seam_ptr : this lane's seam cells in the source plane that will be the in-plane source NEXT iteration
//   out      : this lane's quad in the output plane
//   DOWN     : the item is marched towards decreasing z (slab kernel, odd steps): the window then holds the planes
//              in descending order, so it is handed to lap_quad reversed -- the arithmetic (and its rounding)
//              is the same whatever the direction
template <int R, bool FUSED, bool DOWN = false>
__device__ __forceinline__ void steady_plane(Consumer& c, bool drain, bool prefetch_seam, const float* seam_ptr,
                                             int64_t src_field, float* out, float* mirror, int64_t dst_field,
                                             float4 (&wu)[5], float4 (&wv)[5], float2 (&seam_next)[2]) {
  const float* P = c.P;
  mbar_wait(&c.full[c.s], c.parity);
  {
    const float* st = c.ring + c.s * STAGE_FLOATS + (c.row + 2) * TX + 4 * c.lane;
    wu[(R + 4) % 5] = lds128(st);
    wv[(R + 4) % 5] = lds128(st + ROWS * TX);
  }
  if (drain) {  // planes past the chunk end are z-neighbours only
    __syncwarp();
    if (c.lane == 0) mbar_arrive_after(&c.empty[c.s], wv[(R + 4) % 5].w);
  }
  const float2 seam_u = seam_next[0], seam_v = seam_next[1];
  if (prefetch_seam) {
    ldg_f2_if(c.is_seam, seam_ptr, seam_next[0]);
    ldg_f2_if(c.is_seam, seam_ptr + src_field, seam_next[1]);
  }
  const uint32_t s2 = (c.s + STAGES - 2) & (STAGES - 1);
  const float* sp = c.ring + s2 * STAGE_FLOATS + c.row * TX + 4 * c.lane;
  const float4 wl_u[5] = {wu[(R + (DOWN ? 4 : 0)) % 5], wu[(R + (DOWN ? 3 : 1)) % 5], wu[(R + 2) % 5],
                          wu[(R + (DOWN ? 1 : 3)) % 5], wu[(R + (DOWN ? 0 : 4)) % 5]};
  const float4 wl_v[5] = {wv[(R + (DOWN ? 4 : 0)) % 5], wv[(R + (DOWN ? 3 : 1)) % 5], wv[(R + 2) % 5],
                          wv[(R + (DOWN ? 1 : 3)) % 5], wv[(R + (DOWN ? 0 : 4)) % 5]};
  const float4 cu = wl_u[2], cv = wl_v[2];
  float2 Lu_lo, Lu_hi, Lv_lo, Lv_hi;
  {
    const float4 y[4] = {lds128(sp), lds128(sp + TX), lds128(sp + 3 * TX), lds128(sp + 4 * TX)};
    float Lz = __shfl_up_sync(0xffffffffu, cu.z, 1), Lw = __shfl_up_sync(0xffffffffu, cu.w, 1);
    float Rx = __shfl_down_sync(0xffffffffu, cu.x, 1), Ry = __shfl_down_sync(0xffffffffu, cu.y, 1);
    if (c.lane == 0) { Lz = seam_u.x; Lw = seam_u.y; }
    if (c.lane == 31) { Rx = seam_u.x; Ry = seam_u.y; }
    lap_quad(P, wl_u, y, Lz, Lw, Rx, Ry, Lu_lo, Lu_hi);
  }
  {
    const float* spv = sp + ROWS * TX;
    const float4 y[4] = {lds128(spv), lds128(spv + TX), lds128(spv + 3 * TX), lds128(spv + 4 * TX)};
    float Lz = __shfl_up_sync(0xffffffffu, cv.z, 1), Lw = __shfl_up_sync(0xffffffffu, cv.w, 1);
    float Rx = __shfl_down_sync(0xffffffffu, cv.x, 1), Ry = __shfl_down_sync(0xffffffffu, cv.y, 1);
    if (c.lane == 0) { Lz = seam_v.x; Lw = seam_v.y; }
    if (c.lane == 31) { Rx = seam_v.x; Ry = seam_v.y; }
    lap_quad(P, wl_v, y, Lz, Lw, Rx, Ry, Lv_lo, Lv_hi);
  }
  // the y-neighbour rows of plane k-2 are no longer needed (both Laplacians depend on every row load)
  __syncwarp();
  if (c.lane == 0) mbar_arrive_after(&c.empty[s2], Lu_lo.x + Lv_lo.x);
  const float au = P[P_ALPHA + 0], av = P[P_ALPHA + 1], dt = P[P_DT];
  const float2 ou_lo = fma2(fma2(Lu_lo, au, cubic2(P + P_POLY, lo(cu), lo(cv))), dt, lo(cu));
  const float2 ou_hi = fma2(fma2(Lu_hi, au, cubic2(P + P_POLY, hi(cu), hi(cv))), dt, hi(cu));
  const float2 ov_lo = fma2(fma2(Lv_lo, av, cubic2(P + P_POLY + 10, lo(cu), lo(cv))), dt, lo(cv));
  const float2 ov_hi = fma2(fma2(Lv_hi, av, cubic2(P + P_POLY + 10, hi(cu), hi(cv))), dt, hi(cv));
  const float4 ou = make_float4(ou_lo.x, ou_lo.y, ou_hi.x, ou_hi.y);
  const float4 ov = make_float4(ov_lo.x, ov_lo.y, ov_hi.x, ov_hi.y);
  *reinterpret_cast<float4*>(out) = ou;
  *reinterpret_cast<float4*>(out + dst_field) = ov;
  if (FUSED && mirror != nullptr) {   // boundary plane: also the neighbour GPU's ghost plane (peer store over NVLink)
    *reinterpret_cast<float4*>(mirror) = ou;
    *reinterpret_cast<float4*>(mirror + dst_field) = ov;
  }
  advance_stage(c);
}

// (the kernels themselves -- periodic and slab mode -- are in kernels_gs3d_slab.cuh: one body, three modes)

}  // namespace tma3d
}  // namespace percnn
