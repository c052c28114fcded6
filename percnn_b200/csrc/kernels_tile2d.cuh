// 2-D cells (Pi-block k = 1, Burgers and lambda-omega physics cells; fp32 and fp64) on shared-memory tiles with
// TEMPORAL BLOCKING: several time steps per pass over the grid.
//
// The 2-D BASELINE configs (128^2 .. 512^2) keep their state in L2 and hold well under a microsecond of arithmetic
// per step, so a rollout is bound by what happens BETWEEN steps: a launch (4-5 us) or, in the persistent
// gather kernel k_multi_step, one grid barrier plus an L2 round trip per step (2.5 us/step at 256^2, 7-9 us at 512^2
// fp64 -- round 1).  Here one cooperative launch runs the whole rollout in passes of K steps:
//
//   * a CTA owns one tile; per pass it loads the tile plus a halo of 2K rows / hx columns of both fields into
//     shared memory with 128-bit loads (periodic wrap by addressing; tile origins, halo width and W are multiples
//     of the vector width, so no vector straddles the wrap),
//   * advances the region K steps in shared memory, ping-ponging between two buffers: one vector of cells per
//     thread and item, the cross neighbourhood from 7 x 128-bit shared-memory loads per field, arithmetic by the
//     same per-cell functions as the gather kernels (point_ops.cuh), so results are bit-identical to them.  The
//     rows that can still be valid shrink by 2 per step; columns are not shrunk (garbage creeps in from the
//     un-computed outer vector by 2 cells per step and never reaches the tile: hx >= VW + 2 (K - 1)),
//   * writes the tile's cells of every step that must be kept (tape / emitted frames) straight from the registers
//     that computed them, and the state after the last step of the pass for the next pass,
//   * meets the other CTAs at ONE grid barrier per pass.
#pragma once
#include "kernels_generic.cuh"

namespace percnn {
namespace tile2d {

constexpr int THREADS = 512;
constexpr int kMaxSteps = 4096;   // emit mask travels as a kernel parameter

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int W = 4;
  typedef float4 type;
};
template <>
struct Vec<double> {
  static constexpr int W = 2;
  typedef double2 type;
};

template <typename T>
struct Args {
  const T* h0;          // initial state
  T* tape;              // tape mode: state s+1 -> tape + (s+1) * stride (tape[0] holds h0); else nullptr
  T* traj;              // emit mode: state after step s -> next slot of traj if bit s of emit[] is set; else nullptr
  T* ping;              // pass-to-pass state when there is no tape
  T* pong;
  T* final_state;       // state after the last step (may be nullptr in tape mode)
  int nsteps;
  int K;                // steps per pass
  int TH, TW;           // tile
  int hy, hx;           // halo rows / columns per side (hy = 2K, hx = roundup(VW + 2(K-1), VW))
  int nty, ntx;         // tiles along y / x
  int64_t stride;       // elements between tape / traj slots
  uint32_t emit[kMaxSteps / 32];
};

template <typename T>
__device__ __forceinline__ typename Vec<T>::type ld_cg(const T* p);
template <>
__device__ __forceinline__ float4 ld_cg<float>(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ double2 ld_cg<double>(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

template <typename T>
__device__ __forceinline__ void unpack(const typename Vec<T>::type& v, T* out);
template <>
__device__ __forceinline__ void unpack<float>(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <>
__device__ __forceinline__ void unpack<double>(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
template <typename T>
__device__ __forceinline__ typename Vec<T>::type pack(const T* o);
template <>
__device__ __forceinline__ float4 pack<float>(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
template <>
__device__ __forceinline__ double2 pack<double>(const double* o) { return make_double2(o[0], o[1]); }

// CELL: 0 = Pi-block k=1 (folded cubic), 1 = Burgers physics, 2 = lambda-omega physics  (as k_multi_step)
template <typename T, int CELL>
__global__ void __launch_bounds__(THREADS, 1) k_tile2d(Geom g, int slot, const __grid_constant__ Args<T> a, unsigned* counter) {
  constexpr int VW = Vec<T>::W;
  typedef typename Vec<T>::type V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const T* P = PrepView<T>::get(c_prep[slot]);
  const int RH = a.TH + 2 * a.hy, RW = a.TW + 2 * a.hx;
  const int nq = RW / VW;                       // vectors per region row
  const int64_t fsz = int64_t(RH) * RW;         // elements per field per buffer
  const int ntiles = a.nty * a.ntx;
  const int H = g.H, W = g.W;
  int emitted_before = 0;                       // emit mode: slots used by earlier passes
  unsigned barrier_no = 0;
  for (int s0 = 0; s0 < a.nsteps; s0 += a.K) {
    const int k_eff = min(a.K, a.nsteps - s0);
    const bool last_pass = s0 + k_eff >= a.nsteps;
    const int pass = s0 / a.K;
    const T* src = a.tape != nullptr ? a.tape + int64_t(s0) * a.stride : (pass == 0 ? a.h0 : ((pass & 1) ? a.ping : a.pong));
    T* pass_dst = nullptr;                      // where the state after this pass goes (besides tape / traj)
    if (a.tape == nullptr) pass_dst = last_pass ? a.final_state : ((pass & 1) ? a.pong : a.ping);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tyi = tile / a.ntx, txi = tile - tyi * a.ntx;
      const int y0 = min(tyi * a.TH, H - a.TH), x0 = min(txi * a.TW, W - a.TW);   // the last tile of a row/column is shifted back
      // ---- load the region (both fields) into buffer 0 ----
      for (int i = threadIdx.x; i < 2 * RH * nq; i += THREADS) {
        const int q = i % nq;
        const int r = (i / nq) % RH;
        const int f = i / (nq * RH);
        int gy = y0 - a.hy + r;
        gy %= H;
        if (gy < 0) gy += H;
        int gx = x0 - a.hx + q * VW;
        gx %= W;
        if (gx < 0) gx += W;
        const V v = ld_cg<T>(src + int64_t(f) * g.field + int64_t(gy) * W + gx);
        *reinterpret_cast<V*>(sm + int64_t(f) * fsz + int64_t(r) * RW + q * VW) = v;
      }
      __syncthreads();
      // ---- K steps in shared memory ----
      int slot_in_pass = 0;
      for (int j = 0; j < k_eff; ++j) {
        const int s = s0 + j;                   // this sub-step produces state s + 1
        const T* cur = sm + int64_t(j & 1) * 2 * fsz;
        T* nxt = sm + int64_t((j + 1) & 1) * 2 * fsz;
        T* keep = nullptr;                      // global destination of the tile's cells of state s + 1
        if (a.tape != nullptr) keep = a.tape + int64_t(s + 1) * a.stride;
        else if (a.traj != nullptr && ((a.emit[s >> 5] >> (s & 31)) & 1u)) keep = a.traj + int64_t(emitted_before + slot_in_pass) * a.stride;
        if (a.traj != nullptr && ((a.emit[s >> 5] >> (s & 31)) & 1u)) ++slot_in_pass;
        T* keep2 = (j == k_eff - 1) ? pass_dst : nullptr;   // the state the next pass starts from
        const int r_lo = 2 * (j + 1), r_hi = RH - 2 * (j + 1);
        const int nrow = r_hi - r_lo, ncol = nq - 2;
        const float inv_ncol = 1.0f / float(ncol);
        const int fszi = int(fsz);              // a region is far below 2^31 elements: 32-bit shared-memory indices
        for (int i = threadIdx.x; i < nrow * ncol; i += THREADS) {
          const int ri = int((float(i) + 0.5f) * inv_ncol);   // i / ncol (exact: i < 2^16, no integer division in the loop)
          const int r = r_lo + ri;
          const int q = 1 + (i - ri * ncol);
          T cu[3 * VW], cv[3 * VW];             // centre row: vectors q-1, q, q+1
          T yu[4][VW], yv[4][VW];               // rows r-2, r-1, r+1, r+2 at vector q
          const int o = r * RW + q * VW;
          const T* pu = cur + o;
          const T* pv = pu + fszi;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            unpack<T>(*reinterpret_cast<const V*>(pu + (k - 1) * VW), cu + k * VW);
            unpack<T>(*reinterpret_cast<const V*>(pv + (k - 1) * VW), cv + k * VW);
          }
          const int yo[4] = {-2, -1, 1, 2};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            unpack<T>(*reinterpret_cast<const V*>(pu + yo[k] * RW), yu[k]);
            unpack<T>(*reinterpret_cast<const V*>(pv + yo[k] * RW), yv[k]);
          }
          T ou[VW], ov[VW];
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            Cross<T, 2> U, Vc;
            U.c = cu[VW + e];
            Vc.c = cv[VW + e];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              U.n[0][k] = yu[k][e];
              Vc.n[0][k] = yv[k][e];
              U.n[1][k] = cu[VW + e + yo[k]];
              Vc.n[1][k] = cv[VW + e + yo[k]];
            }
            if (CELL == 0) pi_k1_fwd_poly<T>(U.c, Vc.c, lap_apply<T, 2>(U, P), lap_apply<T, 2>(Vc, P), P, ou[e], ov[e]);
            else if (CELL == 1) burgers_fwd<T>(U, Vc, P, ou[e], ov[e]);
            else lo_fwd<T>(U.c, Vc.c, lap_apply<T, 2>(U, P), lap_apply<T, 2>(Vc, P), P, ou[e], ov[e]);
          }
          const V vu = pack<T>(ou), vv = pack<T>(ov);
          *reinterpret_cast<V*>(nxt + o) = vu;
          *reinterpret_cast<V*>(nxt + fszi + o) = vv;
          // cells of the tile itself: to the tape / trajectory / next pass, straight from these registers
          const int ty_ = r - a.hy, tx_ = q * VW - a.hx;
          if ((keep != nullptr || keep2 != nullptr) && ty_ >= 0 && ty_ < a.TH && tx_ >= 0 && tx_ < a.TW) {
            const int64_t off = int64_t(y0 + ty_) * W + (x0 + tx_);
            if (keep != nullptr) {
              *reinterpret_cast<V*>(keep + off) = vu;
              *reinterpret_cast<V*>(keep + g.field + off) = vv;
            }
            if (keep2 != nullptr) {
              *reinterpret_cast<V*>(keep2 + off) = vu;
              *reinterpret_cast<V*>(keep2 + g.field + off) = vv;
            }
          }
        }
        __syncthreads();
      }
    }
    // emit mode: count the slots this pass filled (same for every CTA)
    if (a.traj != nullptr)
      for (int j = 0; j < k_eff; ++j) emitted_before += int((a.emit[(s0 + j) >> 5] >> ((s0 + j) & 31)) & 1u);
    if (!last_pass) grid_barrier(counter, ++barrier_no * gridDim.x);
  }
}

}  // namespace tile2d
}  // namespace percnn
