/*
 * percnn_b200.h -- C ABI of the B200-native PeRCNN recurrent-cell library (libpercnn_b200.so).
 *
 * The reference (isds-neu/PeRCNN) is 100 % Python and has no FFI layer; the boundary a replacement
 * must honour is the nn.Module surface of `RCNNCell` / `RCNN`.  This header
 * is the C-ABI that sits directly under that surface: plain pointers and sizes, no torch types, no
 * C++ exceptions.  percnn_b200/_lib.py binds it with ctypes; INTEGRATION.md shows the stub.
 *
 * Each entry point names the reference code it replaces, as <alias>:<lines> with the aliases of
 * SURVEY.md (GS2D = DataDrivenModeling/2d_gs_rd/train_2drd.py, GS3D = .../3d_gs_rd/train_3drd.py,
 * FWD = ForwardSimulationOfPDEs/2d_lambda_omega/percnn_LO_eqn.py, BUR1/LO1 = Stage-1 scripts,
 * BUR3/LO3 = Stage-3 scripts).
 *
 * Conventions
 *  - return 0 = work enqueued; non-zero = percnn_status_t, message via percnn_last_error().
 *  - every device buffer is owned by the caller; calls allocate nothing and never synchronise the
 *    host (stream-ordered, CUDA-graph capturable) except the *_host entry points, which own their
 *    staging buffers and block until the result is in host memory.
 *  - state layout: [2 fields][D][H][W] (3-D) or [2][H][W] (2-D), W fastest, contiguous -- exactly the
 *    reference's h[0] of shape [1,2,(D,)H,W].
 *  - `params` / `param_grads`: flat array of plan-dtype scalars = the cell's state_dict tensors
 *    concatenated in state_dict order (see percnn_param_count and DESIGN.md "parameter packing").
 *  - a plan is used from one host thread AND one stream at a time (it owns one constant-memory parameter slot, one
 *    grid-barrier counter and its tensor maps); distinct plans are independent.  Every entry point makes the
 *    plan's device current for the duration of the call and restores the caller's device.
 */
#ifndef PERCNN_B200_H_
#define PERCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PERCNN_ABI_VERSION 2

typedef enum {
  PERCNN_OK = 0,
  PERCNN_ERR_INVALID = 1,      /* bad descriptor / argument */
  PERCNN_ERR_UNSUPPORTED = 2,  /* valid but not implemented (e.g. non-cross Laplacian table) */
  PERCNN_ERR_CUDA = 3,         /* a CUDA runtime/driver call failed */
  PERCNN_ERR_NO_DEVICE = 4     /* no sm_100 device / driver */
} percnn_status_t;

typedef enum { PERCNN_F32 = 0, PERCNN_F64 = 1 } percnn_dtype_t;

typedef enum {
  PERCNN_CELL_PI = 0,       /* Pi-block cell: GS2D:105-121, FWD:98-112, GS3D:123-139, BUR1:142-178, LO1:142-171 */
  PERCNN_CELL_BURGERS = 1,  /* Stage-3 physics cell with d/dx, d/dy advection: BUR3:154-157,209-221 */
  PERCNN_CELL_LO = 2        /* Stage-3 lambda-omega polynomial cell: LO3:148-151,203-215 */
} percnn_cell_t;

typedef enum {
  PERCNN_COEF_RAW = 0,      /* alpha = DA (FWD:107) or nu_u (BUR3:155) */
  PERCNN_COEF_SIGMOID = 1   /* alpha = mu_up * sigmoid(CA) (GS2D:115) */
} percnn_coef_mode_t;

enum {                       /* percnn_desc_t.flags */
  PERCNN_FLAG_EVAL_BRANCH = 1,   /* k=1: evaluate the three conv branches channel by channel exactly as the
                                    reference does instead of the folded bivariate cubic (slower; parity aid) */
  PERCNN_FLAG_NO_TMA = 2,        /* never pick the TMA z-marching kernels (generic kernels only) */
  PERCNN_FLAG_LO_C6 = 4          /* PERCNN_CELL_LO carries the 13th coefficient C6_v (10 %-noise script) */
};

typedef struct percnn_desc {
  int32_t abi_version;   /* PERCNN_ABI_VERSION */
  int32_t ndim;          /* 2 | 3 */
  int64_t extent[3];     /* D, H, W (2-D: extent[0] = 1).  In slab mode D is the LOCAL depth. */
  int32_t dtype;         /* percnn_dtype_t */
  int32_t cell;          /* percnn_cell_t */
  int32_t ksize;         /* Pi conv kernel size: 1 | 5 (0 for physics cells) */
  int32_t hidden;        /* Pi hidden channels hc (<= 16) */
  int32_t coef_mode;     /* percnn_coef_mode_t */
  int32_t flags;
  double mu_up;          /* GS2D:58 mu_up / BUR1:96 nu_up; ignored for RAW */
  double dt;             /* GS2D:57 */
  double dx;             /* Stage-3 cells divide the un-scaled stencils by dx, dx^2 (BUR3:78-80) */
  int32_t device;        /* CUDA device ordinal */
  int32_t slab_ghost;    /* 0: state buffers are periodic along every axis (single GPU).
                            1: slab mode -- state buffers carry 2 ghost planes (rows in 2-D) on each side of the
                               slowest axis, i.e. [2][D+4][H][W]; the caller (halo exchange) fills them. */
} percnn_desc_t;

typedef struct percnn_plan percnn_plan_t;

/* ---- library ---------------------------------------------------------------------------------- */
int percnn_abi_version(void);
const char* percnn_last_error(void);              /* thread-local; valid until the next call on the thread */
int percnn_device_ok(int device);                 /* 1 if `device` exists and is sm_100; no context side effects beyond cudaGetDeviceProperties */

/* ---- plans ------------------------------------------------------------------------------------ */
/* Replaces RCNNCell.__init__ (GS2D:46-90): fixes geometry, dtype, variant and constants. */
int percnn_plan_create(const percnn_desc_t* desc, percnn_plan_t** out);
int percnn_plan_destroy(percnn_plan_t* plan);
/* Number of scalars in `params` (and `param_grads`). */
int64_t percnn_param_count(const percnn_plan_t* plan);
/* Scalars in one state buffer as the kernels see it (includes ghost planes in slab mode). */
int64_t percnn_state_elems(const percnn_plan_t* plan);
/* Bytes of caller-provided scratch needed by rollout_fwd / rollout_bwd for `nsteps`. */
size_t percnn_workspace_bytes(const percnn_plan_t* plan, int nsteps);
/* 1 if the plan runs the TMA z-marching kernel, 0 if the generic one (introspection for tests/bench). */
int percnn_plan_uses_tma(const percnn_plan_t* plan);
/* > 0 if rollouts of this (2-D) plan run on shared-memory tiles with temporal blocking: the number of time steps
 * per pass (one grid barrier per pass); 0 if the gather kernels serve it. */
int percnn_plan_uses_tile2d(const percnn_plan_t* plan);
/* Kernel launches issued by this plan since creation (bench "gpu_launches"). */
int64_t percnn_plan_launch_count(const percnn_plan_t* plan);
/* 1 if percnn_slab_rollout_fwd runs this (slab-mode) plan's rollouts of >= 2 steps as ONE persistent cooperative kernel
 * (small slabs: the grid barrier doubles as the halo hand-shake) instead of one fused-halo kernel per step.  The
 * persistent kernel is the gather kernel (PERCNN_FLAG_NO_TMA arithmetic), not the TMA z-march. */
int percnn_plan_slab_persistent(const percnn_plan_t* plan);

/* ---- parameters ------------------------------------------------------------------------------- */
/* Digest the raw parameter tensors (device pointer, plan dtype) into the constant block the kernels
 * read: alpha = mu_up*sigmoid(CA) (GS2D:115), the cross taps of W_laplace.weight (GS2D:66), the folded
 * cubic of the 1x1 Pi-block (GS2D:115 `Wh4(Wh1*Wh2*Wh3)`), or the repacked 5x5 filters (BUR1:108-124).
 * Must precede step/rollout calls whenever the parameters changed.  Stream-ordered. */
int percnn_params_load(percnn_plan_t* plan, const void* params, void* stream);

/* ---- one time step ---------------------------------------------------------------------------- */
/* RCNNCell.forward (GS2D:105-121 and siblings): h_out = h_in + dt * (alpha * Lap(h_in) + Pi(h_in)). */
int percnn_step_fwd(percnn_plan_t* plan, const void* h_in, void* h_out, void* stream);
/* RCNNCell.forward_rk4 (BUR3:159-206, LO3:153-200; defined by the Stage-3 scripts, never called by them): one
 * classical RK4 step h_out = h + dt (k1 + 2 k2 + 2 k3 + k4) / 6 of the physics cell's f_rhs, as four launches of the
 * fused right-hand side (the intermediate states h + a k are formed on the fly).  `ws`: percnn_workspace_bytes. */
int percnn_step_rk4(percnn_plan_t* plan, const void* h_in, void* h_out, void* ws, void* stream);
/* Same step restricted to interior planes [z_lo, z_hi) of the slowest axis (3-D TMA plans only).  Lets the
 * slab-mode driver launch the planes that do not touch ghost cells before the halo exchange lands. */
int percnn_step_fwd_range(percnn_plan_t* plan, const void* h_in, void* h_out, int z_lo, int z_hi, void* stream);
/* Slab-mode step with the halo exchange fused into the kernel (multi-GPU, peer-mapped buffers over NVLink; replaces
 * nothing in the reference, which is single-GPU -- it is north_star's "domain-decomposed ... halo exchange of the
 * ghost cells only").  ONE kernel per time step marches every tile through all local planes in a
 * single pass; a helper warp per CTA copies the boundary planes 0,1 and D-2,D-1 of the output into the neighbours'
 * ghost planes and, when all tiles of a pair have landed, raises that neighbour's flag (release, system scope).
 * Before its first TMA load of a ghost plane the kernel waits (acquire) for my_flags >= epoch.  The march direction
 * alternates with the parity of `epoch` (even: ascending z, odd: descending), so that every ghost plane is produced
 * almost a full step before it is consumed and no NVLink latency is exposed.  Inside a rollout the pair a step
 * produces LAST is published by the next step's kernel (PERCNN_SLAB_DEFER_LATE / PERCNN_SLAB_FLUSH_PREV).  A wait
 * that exceeds its spin deadline sets the error word and TRAPS (the rollout must not continue on stale ghosts): the
 * caller sees a CUDA error at its next synchronisation.  No NCCL call and no host round trip per step.  All
 * pointers are device pointers. */
enum {                       /* percnn_slab_link_t.flags */
  PERCNN_SLAB_FLUSH_PREV = 1,   /* the previous step (same rollout, epoch - 1) ran with DEFER_LATE: publish its last
                                   boundary pair from h_in first (needs peer_lo_in / peer_hi_in) */
  PERCNN_SLAB_DEFER_LATE = 2    /* leave the boundary pair produced last to the next step (which must FLUSH_PREV) */
};
typedef struct percnn_slab_link {
  void* peer_lo_out;        /* lower ring neighbour's h_out buffer (same [2][D+4][H][W] layout) */
  void* peer_hi_out;        /* upper ring neighbour's h_out buffer */
  const uint32_t* my_flags; /* [0]: epoch up to which my LOWER ghosts are valid, [1]: same for the UPPER ghosts */
  uint32_t* peer_lo_flags;  /* the lower neighbour's flags array (its [1] is raised by this step) */
  uint32_t* peer_hi_flags;  /* the upper neighbour's flags array (its [0] is raised by this step) */
  uint32_t* scratch;        /* 3 local words, zero-initialised: arrival counter (lower boundary), error word (spin
                               deadline), arrival counter (upper boundary) */
  uint32_t epoch;           /* waits for flags >= epoch, publishes epoch + 1; parity selects the march direction */
  uint32_t flags;           /* PERCNN_SLAB_* */
  void* peer_lo_in;         /* the neighbours' mappings of THEIR h_in buffers (only read with FLUSH_PREV) */
  void* peer_hi_in;
} percnn_slab_link_t;
int percnn_step_fwd_fused_halo(percnn_plan_t* plan, const void* h_in, void* h_out, const percnn_slab_link_t* link,
                               void* stream);
/* Adjoint of the fused slab step: g_in's boundary planes are mirrored into the neighbours' g_in ghost planes. */
int percnn_step_bwd_fused_halo(percnn_plan_t* plan, const void* h_in, const void* g_out, const void* g_add, void* g_in,
                               void* ws, const percnn_slab_link_t* link, void* stream);
/* The ping-pong buffers of a slab rank and their peer mappings: everything a whole slab rollout needs. */
typedef struct percnn_slab_ring {
  void* buf[2];             /* my two state (or gradient) buffers, [2][D+4][H][W] each */
  void* peer_lo_buf[2];     /* the lower neighbour's two buffers (peer-mapped) */
  void* peer_hi_buf[2];     /* the upper neighbour's */
  const uint32_t* my_flags; /* as in percnn_slab_link_t */
  uint32_t* peer_lo_flags;
  uint32_t* peer_hi_flags;
  uint32_t* scratch;
} percnn_slab_ring_t;
/* RCNN.forward's loop (GS3D:186-214) on one slab: `nsteps` fused steps issued from ONE host call; step s reads
 * buf[cur ^ (s & 1)], writes the other buffer and uses epoch + s.  The state ends in buf[cur ^ (nsteps & 1)]. */
int percnn_slab_rollout_fwd(percnn_plan_t* plan, const percnn_slab_ring_t* ring, int cur, int nsteps, uint32_t epoch,
                            void* stream);
/* Communication-avoiding variant for SMALL slabs (percnn_plan_slab_persistent(plan) != 0; cfg4-class grids): the whole
 * rollout is one persistent kernel that exchanges 2k ghost planes every k time steps instead of 2 planes every step
 * (the per-step hand-shake over NVLink costs more than such a slab's arithmetic).  `wide`: four scratch buffers
 * [2][D + 4k][H][W] of this rank (contents irrelevant) and the neighbours' mappings of theirs; 2k <= D.  Same
 * contract as percnn_slab_rollout_fwd otherwise (ring buffers in/out, flags, epochs): the two are interchangeable. */
typedef struct percnn_slab_wide {
  void* buf[4];
  void* peer_lo_buf[4];
  void* peer_hi_buf[4];
  int32_t k;
  int32_t reserved;
} percnn_slab_wide_t;
int percnn_slab_rollout_fwd_blocked(percnn_plan_t* plan, const percnn_slab_ring_t* ring, const percnn_slab_wide_t* wide,
                                    int cur, int nsteps, uint32_t epoch, void* stream);
/* Same, keeping every state: step t reads tape slot t and writes slot t+1, mirroring the boundary planes into the
 * neighbours' slot t+1 (`peer_*_tape` are the peer mappings of their tapes; slots are percnn_state_elems apart). */
int percnn_slab_rollout_tape(percnn_plan_t* plan, void* tape, void* peer_lo_tape, void* peer_hi_tape,
                             const percnn_slab_ring_t* ring, int nsteps, uint32_t epoch, void* stream);
/* Adjoint of one step (replaces autograd through GS2D:105-121): g_in = g_add + (dh_out/dh_in)^T g_out
 * (g_add may be NULL), and the step's parameter-gradient sums are ACCUMULATED into the accumulator at the
 * head of `ws` (zeroed by percnn_param_grads_begin, read by percnn_param_grads_finish). */
int percnn_step_bwd(percnn_plan_t* plan, const void* h_in, const void* g_out, const void* g_add, void* g_in,
                    void* ws, void* stream);
int percnn_param_grads_begin(percnn_plan_t* plan, void* ws, void* stream);   /* zero the accumulator */
/* Turn the accumulated partials into gradients w.r.t. the raw parameters (same packing as `params`). */
int percnn_param_grads_finish(percnn_plan_t* plan, const void* params, void* param_grads, void* ws, void* stream);

/* ---- whole rollout ---------------------------------------------------------------------------- */
/* RCNN.forward loop (GS2D:169-188).  Runs `nsteps` steps from h0.  emit[s] != 0 stores the state after
 * step s into the next slot of `traj` (slots are percnn_state_elems apart; slot order = step order).
 * `h_final` (optional) receives the state after the last step.  With `tape` != NULL every state
 * h_0..h_nsteps is kept there ([nsteps+1] slots) for rollout_bwd; `ws` may then be NULL. */
int percnn_rollout_fwd(percnn_plan_t* plan, const void* h0, void* traj, const uint8_t* emit, int nsteps,
                       void* h_final, void* tape, void* ws, void* stream);
/* Back-propagation through the unrolled rollout (replaces loss.backward() through GS2D:169-188).
 * `tape` holds h_0..h_nsteps ([nsteps+1] slots, as written by rollout_fwd).  gmask[s] != 0 (s = 0..nsteps)
 * says dL/dh_s is present; the present gradients are packed in increasing s in `g_tape` (NULL = none).
 * Writes dL/dh_0 to g_h0 and dL/dparams (same packing as `params`) to param_grads. */
int percnn_rollout_bwd(percnn_plan_t* plan, const void* params, const void* tape, const void* g_tape,
                       const uint8_t* gmask, int nsteps, void* g_h0, void* param_grads, void* ws, void* stream);

/* ---- fused data loss ---------------------------------------------------- */
/* The training scripts' data loss is a strided-subsample MSE over selected states of the rollout:
 *     mse_loss(output[:-1:15, :, ::2, ::2, ::2], truth[::15, :, ::2, ::2, ::2])          (GS3D:403)
 *     mse_loss(output[0:-1:20, :, ::4, ::4], truth[::20, :, ::4, ::4])                   (GS2D:397-401)
 *     mse_loss(output[0:-1:s_t, :, ::s, ::s], truth...)                                  (BUR1:610-614)
 * Stock autograd materialises a dense [T+1, 2, ...] gradient for it.  Here the loss is one small reduction over
 * the sampled points of the tape, and its gradient 2/N (h_s - target) is INJECTED by the adjoint kernel of step s
 * itself (it already holds h_s in registers): no dense gradient tape, 8/s^ndim extra bytes per cell on the
 * selected steps only. */
typedef struct percnn_data_loss {
  const void* target;    /* device, plan dtype: the selected states' low-res frames, packed in increasing step
                            order, each [2][ceil(D/s)][ceil(H/s)][ceil(W/s)] (2-D: [2][ceil(H/s)][ceil(W/s)]) */
  const uint8_t* sel;    /* host, [nsteps + 1]: sel[s] != 0 <=> state h_s enters the loss */
  int32_t stride;        /* spatial subsampling stride s >= 1 (`::s` on every spatial axis) */
  int32_t reserved;      /* must be 0 */
  int64_t n_total;       /* number of elements the mean runs over; 0 = this plan's own count
                            (nsel * 2 * prod ceil(extent/s)).  Slab ranks pass the global count. */
  const void* gscale;    /* device scalar (plan dtype) dL/dloss for the backward pass; NULL = 1 */
} percnn_data_loss_t;
/* loss = sum over selected states and sampled points of (h - target)^2 / n_total, written to *loss_out (device,
 * one plan-dtype scalar).  `tape` as written by percnn_rollout_fwd.  Uses the head of `ws`; deterministic
 * (fixed-order fp64 partial sums). */
int percnn_data_loss_fwd(percnn_plan_t* plan, const void* tape, int nsteps, const percnn_data_loss_t* loss,
                         void* loss_out, void* ws, void* stream);
/* percnn_step_bwd with the loss gradient of ONE selected state injected: g_in += gscale * 2/n_total * (h_in -
 * target_frame) at the sampled points.  `target_frame` is that state's low-res frame (NULL = no injection);
 * `link` (NULL = single GPU) selects the fused-halo slab step. */
int percnn_step_bwd_loss(percnn_plan_t* plan, const void* h_in, const void* g_out, const void* g_add,
                         const void* target_frame, int stride, int64_t n_total, const void* gscale, void* g_in,
                         void* ws, const percnn_slab_link_t* link, void* stream);
/* percnn_rollout_bwd with the fused data loss as an additional gradient source (`g_tape`/`gmask` may be NULL). */
int percnn_rollout_bwd_loss(percnn_plan_t* plan, const void* params, const void* tape, const void* g_tape,
                            const uint8_t* gmask, const percnn_data_loss_t* loss, int nsteps, void* g_h0,
                            void* param_grads, void* ws, void* stream);
/* Back-propagation through a taped slab rollout (loss.backward() of GS3D:407 on one slab): ring->buf[0] holds
 * dL/dh_nsteps with exchanged ghosts; step t = nsteps-1 .. 0 runs the fused-halo adjoint from buf[b] to buf[b ^ 1]
 * (b starts at 0), so dL/dh_0 ends in buf[nsteps & 1].  `g_tape` (nullable): dense dL/d(tape slot t) in the ghosted
 * layout, added at step t.  `loss` (nullable): fused data loss with the GLOBAL n_total; sel[nsteps] must be 0.
 * Parameter sums accumulate at the head of `ws` (percnn_param_grads_begin before; all-reduce them across ranks and
 * call percnn_param_grads_finish after). */
int percnn_slab_rollout_bwd(percnn_plan_t* plan, const void* tape, const void* g_tape, const percnn_data_loss_t* loss,
                            const percnn_slab_ring_t* ring, int nsteps, uint32_t epoch, void* ws, void* stream);

/* ---- fused physics-residual loss ------------------------------------------ */
/* `loss_gen(output, loss_generator(dt, dx))` of the scripts (FWD:288-357, the TRAINING loss of the forward-simulation
 * script FWD:371-373; GS2D:270-353 and GS3D:286-345, where it is a validation metric):
 *     f_q = D_q * Lap(q_t) + R_q(u_t, v_t) - (q_{t+1} - q_t) / dt,   loss = mse(f_u, 0) + mse(f_v, 0)
 * over frames t = 0 .. nframes-3, with the 4th-order Laplacian / dx^2 on the periodic grid.  The reference pads 2
 * cells before and 3 after every axis, so each axis contributes extent + 1 residual points, the last being the
 * periodic image of the first; that double counting is reproduced exactly (weights, N = (nframes-2) prod(extent+1)).
 * R_q is a bivariate cubic (lambda-omega: FWD:338-339, Gray-Scott: GS2D:322-328, GS3D:319-326).
 * Stand-alone (no plan): works on any [nframes][2][(D,)H,W] contiguous device tensor. */
typedef struct percnn_phys_loss {
  int32_t ndim;          /* 2 | 3 */
  int32_t dtype;         /* percnn_dtype_t */
  int64_t extent[3];     /* D, H, W (2-D: extent[0] = 1) */
  int32_t nframes;       /* frames in `frames` (>= 3); consecutive frames are one time step dt apart */
  int32_t device;
  double diff[2];        /* D_u, D_v */
  double poly[2][10];    /* R_u, R_v as cubics in (u, v): c00 c10 c01 c20 c11 c02 c30 c21 c12 c03 */
  double dt;             /* FWD:283-286 */
  double dx;             /* FWD:276-280: Laplacian table / dx^2 */
} percnn_phys_loss_t;
size_t percnn_phys_loss_workspace_bytes(void);
/* loss -> *loss_out (device scalar of `dtype`).  `resid` (nullable): [nframes-2] frames receiving dloss/df, needed
 * by percnn_phys_loss_bwd.  `ws`: percnn_phys_loss_workspace_bytes() of device scratch. */
int percnn_phys_loss_fwd(const percnn_phys_loss_t* pl, const void* frames, void* resid, void* loss_out, void* ws,
                         void* stream);
/* g_frames[t] = gscale * dloss/dframes[t] for every frame (dense; the last frame's gradient is zero).
 * `gscale`: device scalar of `dtype` (NULL = 1). */
int percnn_phys_loss_bwd(const percnn_phys_loss_t* pl, const void* frames, const void* resid, const void* gscale,
                         void* g_frames, void* stream);

/* ---- initial-state generator ("upscaler") + IC loss ------------------------ */
/* `model.UpconvBlock(model.init_state_low)` of the training scripts (GS2D:26-41,164; GS3D:41-56,186; BUR1:38-52 =
 * LO1:38-52 = BUR3:38-52) and the autograd backward of it (the sink of dL/dh0, GS2D:407):
 *   layers = 2:  ConvTranspose(2->C, k5, s2, p2, op1) -> act -> ConvTranspose(C->C, k5, stride2, p2, op stride2-1) -> Conv1x1(C->2)
 *   layers = 1:  ConvTranspose2d(2->C, k5, s2, p2, op1) -> act -> Conv1x1(C->2)
 * `params`: the module's state_dict tensors concatenated in state_dict order (W1 [2][C][K], b1 [C], (W2 [C][C][K],
 * b2 [C],) W3 [2][C], b3 [2]; K = 5^ndim).  `low`: [2][Dl][Hl][Wl] (always the WHOLE low-resolution input).
 * `mid`: caller-owned tape of the activations, percnn_upscaler_sizes().mid_elems elements, written by _fwd and read by
 * _bwd.  h0 / g_h0: [2] fields of [out_nz][Ho][Wo], `out_field_stride` elements apart (0 = dense).
 * Slab mode (3-D): out_z0 / out_nz select the output planes this call produces, so every rank generates its own slab
 * of h0; in _bwd, g_h0 must then be addressable 2 planes beyond both ends of the range (the ghost planes of the slab
 * layout, holding the neighbours' dL/dh0; planes outside the global grid are never read), and the parameter gradients
 * are this slab's partial sums (all-reduce them).  Stand-alone: no plan needed; stream-ordered, deterministic. */
typedef struct percnn_upscaler {
  int32_t ndim;              /* 2 | 3 */
  int32_t dtype;             /* percnn_dtype_t */
  int32_t channels;          /* C: 8 (GS2D:31, GS3D:46) | 16 (BUR1:44) */
  int32_t act;               /* 0 = sigmoid (GS2D:34) | 1 = tanh (BUR1:46) */
  int32_t layers;            /* 1 | 2 */
  int32_t stride2;           /* stride of the second transposed conv: 2 (GS2D:36) | 1 (GS3D:51); ignored for layers = 1 */
  int32_t device;
  int32_t reserved;
  int64_t low_extent[3];     /* Dl, Hl, Wl (2-D: Dl = 1) */
  int64_t out_z0, out_nz;    /* output plane range; out_nz = 0: the whole grid */
  int64_t out_field_stride;  /* 0 = out_nz * Ho * Wo */
} percnn_upscaler_t;
/* Any output pointer may be NULL.  out_extent[3] = Do, Ho, Wo of the whole output grid. */
int percnn_upscaler_sizes(const percnn_upscaler_t* up, int64_t* nparams, int64_t* mid_elems, int64_t* out_extent,
                          size_t* ws_bytes);
int percnn_upscaler_fwd(const percnn_upscaler_t* up, const void* params, const void* low, void* mid, void* h0, void* ws,
                        void* stream);
/* g_params (same packing as `params`) = dL/dparams for the upstream gradient g_h0; accumulate != 0 adds to it. */
int percnn_upscaler_bwd(const percnn_upscaler_t* up, const void* params, const void* low, const void* mid,
                        const void* g_h0, void* g_params, int accumulate, void* ws, void* stream);
/* `get_ic_loss` (GS2D:331-338, GS3D:325-333, BUR1:462-471): loss_out = mean((a - b)^2) over n elements (fp64
 * partial sums, fixed order); _bwd: g = gscale * 2/n * (a - b)  (gscale: device scalar, NULL = 1; accumulate != 0 adds). */
size_t percnn_mse_workspace_bytes(void);
int percnn_mse_fwd(int dtype, int device, const void* a, const void* b, int64_t n, void* loss_out, void* ws, void* stream);
int percnn_mse_bwd(int dtype, int device, const void* a, const void* b, int64_t n, const void* gscale, void* g,
                   int accumulate, void* stream);

/* ---- Stage-2 library of candidate terms ------------------------------------- */
/* `Loss_generator.get_phy_residual` / `get_library` (2D_Burgers_eqn/Stage-2/derivatives.py:129-199, lambda-omega
 * stage-2/derivatives.py:128-199) on the periodically padded trajectory that `get_residual_mse` builds
 * (derivatives.py:207-208), and the 70-column matrix of PDE_FIND_u.py:185-193,246-259.
 * frames: [nframes][2][H][W].  terms: [12][nframes-2][H+1][W+1] in the order
 *     f_u f_v u v u_t v_t u_x u_y v_x v_y lap_u lap_v
 * ('ones' is implicit); point (i, j) of the (H+1) x (W+1) grid is cell (i mod H, j mod W) -- the reference's padding
 * makes the last row/column the periodic image of the first.  u_x differentiates along tensor dim 2, u_y along dim 3. */
typedef struct percnn_library {
  int32_t dtype;     /* percnn_dtype_t (the scripts run in fp32, derivatives.py:8) */
  int32_t kind;      /* residual: 0 = Burgers, nu = 1/200 (derivatives.py:189-192) | 1 = lambda-omega (stage-2/derivatives.py:188-192) */
  int64_t H, W;
  int32_t nframes;   /* >= 3 */
  int32_t device;
  double dt, dx;     /* derivatives.py:87: dy = dx */
} percnn_library_t;
int64_t percnn_library_points(const percnn_library_t* lib);   /* (nframes-2)(H+1)(W+1), -1 on a bad descriptor */
int percnn_library_terms(const percnn_library_t* lib, const void* frames, void* terms, void* stream);
/* theta[r][a*7+b] = A_a * B_b at flattened point idx[r] (fp64, from the stored terms: `to_numpy_float64` then the
 * eval'd products), A = ones u v u^2 uv v^2 u^3 u^2v uv^2 v^3, B = ones u_x u_y v_x v_y lap_u lap_v;
 * rhs[r] = (u_t, v_t).  idx: n device int64 indices into [0, points). */
int percnn_library_theta(const percnn_library_t* lib, const void* terms, const int64_t* idx, int64_t n, double* theta,
                         double* rhs, void* stream);

/* ---- host-buffer convenience (end-to-end path) ------------------------------------------------ */
/* Same as params_load + rollout_fwd but with HOST pointers: copies params and h0 to the device, runs the
 * rollout, copies the emitted frames (and h_final if non-NULL) back, blocks until done.  Scratch is
 * owned by the plan and reused across calls. */
int percnn_rollout_fwd_host(percnn_plan_t* plan, const void* params_host, const void* h0_host,
                            void* traj_host, const uint8_t* emit, int nsteps, void* h_final_host);

#ifdef __cplusplus
}
#endif
#endif /* PERCNN_B200_H_ */
