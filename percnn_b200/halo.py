"""Slab domain decomposition of the 3-D rollout over N GPUs.

One process per GPU.  The grid is split along the slowest axis D; every rank keeps two ghosted state
buffers [2][nz+4][H][W] (ping-pong) whose 2 ghost planes per side are the neighbours' boundary planes
(periodic ring: neighbours are (rank +- 1) mod N).  Per time step

    compute stream:  wait(ghosts of cur ready) -> boundary planes [0,2) and [nz-2,nz) of nxt
    comm stream:     wait(boundary done)       -> send them / receive the neighbours' into nxt's ghosts
    compute stream:  interior planes [2,nz-2) of nxt            (overlaps the exchange)

so only the ghost cells cross NVLink (2 x 2 x H x W x 4 B per side per step) and the exchange hides
behind the interior kernel.  Transport:

  * "nccl": torch.distributed batch_isend_irecv (ncclSend/ncclRecv grouped) on the comm stream;
  * "symm": peer-mapped buffers (torch.distributed._symmetric_memory): the boundary planes are copied
    straight into the neighbour's ghost planes over NVLink and a stream-ordered signal replaces the
    rendezvous -- no NCCL kernel, no host round trip;
  * "fused" (default when peer mapping works): ONE kernel per step does the compute and the exchange -- a single
    z-march whose boundary planes are stored locally AND into the neighbour's ghost planes through the peer
    mapping; the kernel raises the neighbour's flag when a boundary pair has landed and waits on its own flags
    right before it first touches a ghost plane.  The march direction alternates from step to step, so every
    ghost plane is produced almost a full step before it is needed (csrc/kernels_gs3d_slab.cuh), and the whole
    rollout is issued from ONE C call (percnn_slab_rollout_fwd) -- no per-step Python.

`exchange_ghosts` is device-agnostic (it only moves planes with torch.distributed), which is what the
world_size-2 gloo tests exercise on CPU; the step kernels themselves are CUDA-only.
"""
from __future__ import annotations

from typing import List, Optional

import os

import torch
import torch.distributed as dist

from . import engine


def slab_bounds(depth: int, rank: int, world: int):
    """[z0, z0+nz) of `rank`; the depth must split evenly so every rank runs the same kernel shape."""
    if depth % world:
        raise ValueError(f"depth {depth} is not divisible by {world} ranks")
    nz = depth // world
    return rank * nz, nz


def exchange_ghosts(buf: torch.Tensor, nz: int, rank: int, world: int, group=None) -> List:
    """Fill the 2+2 ghost planes of `buf` ([2][nz+4][...]) from the ring neighbours' boundary planes.

    Returns the outstanding work handles (empty when world == 1, where the wrap is a local copy).
    Send order (top planes -> lower neighbour, bottom planes -> upper neighbour) and receive order (upper
    neighbour first) are chosen so that world == 2, where both neighbours are the same peer, pairs up.
    """
    if world == 1:
        buf[:, 0:2].copy_(buf[:, nz:nz + 2])
        buf[:, nz + 2:nz + 4].copy_(buf[:, 2:4])
        return []
    lo, hi = (rank - 1) % world, (rank + 1) % world
    ops = []
    for f in range(2):
        ops.append(dist.P2POp(dist.isend, buf[f, 2:4], lo, group))
    for f in range(2):
        ops.append(dist.P2POp(dist.isend, buf[f, nz:nz + 2], hi, group))
    for f in range(2):
        ops.append(dist.P2POp(dist.irecv, buf[f, nz + 2:nz + 4], hi, group))
    for f in range(2):
        ops.append(dist.P2POp(dist.irecv, buf[f, 0:2], lo, group))
    return dist.batch_isend_irecv(ops)


def slab_loss_spec(global_shape, z0: int, nz: int, nsteps: int, sel, stride: int) -> "engine.DataLossSpec":
    """Slab-local view of the global fused data loss `mse(states[sel][:, :, ::s, ::s, ::s], truth)`: the sampling
    lattice `::stride` of the global grid must coincide with the slab's own (slab origin and depth multiples of the
    stride), and the mean runs over the GLOBAL number of sampled points, so that the ranks' partial sums simply add
    up (one scalar all-reduce).  Pure host logic (tests/test_halo_gloo.py)."""
    D, H, W = (int(n) for n in global_shape)
    stride = int(stride)
    if stride < 1:
        raise ValueError("stride must be >= 1")
    if z0 % stride or nz % stride:
        raise ValueError(f"slab [{z0}, {z0 + nz}) does not respect the loss stride {stride}")
    sel = tuple(bool(e) for e in sel)
    if len(sel) != nsteps + 1:
        raise ValueError("selection mask needs nsteps + 1 entries")
    if sel[nsteps]:
        raise NotImplementedError("the last state has no adjoint step; the scripts' `[0:-1:...]` never selects it")
    low = [(n + stride - 1) // stride for n in (D, H, W)]
    n_total = sum(sel) * 2 * low[0] * low[1] * low[2]
    return engine.DataLossSpec(sel=sel, stride=stride, n_total=n_total)


class SlabRollout:
    """Forward rollout of a 3-D Pi-block cell on one slab of a slab-decomposed periodic grid."""

    def __init__(self, cell, global_shape, device, rank: int, world: int, group=None, transport: str = "auto",
                 use_graph: bool = True):
        D, H, W = (int(s) for s in global_shape)
        self.rank, self.world, self.group = rank, world, group
        self.z0, self.nz = slab_bounds(D, rank, world)
        if self.nz < 4:
            raise ValueError("slabs need at least 4 planes per rank")
        self.device = torch.device(device)
        self.cell = cell
        self.plan = engine.get_plan(cell._spec(), (self.nz, H, W), self.device, slab_ghost=True)
        if not self.plan.uses_tma:
            raise NotImplementedError("slab mode drives the TMA kernel (W % 128 == 0, H % 16 == 0, fp32, k = 1)")
        self.flat = engine.pack_params(cell._packed_tensors(), torch.float32)
        self.plan.params_load(self.flat)
        self.transport = transport
        self.symm = None
        shape = self.plan.buffer_shape
        if transport in ("auto", "symm", "fused") and world > 1:
            # peer-mapped buffers over NVLink; every rank must take the same branch, so agree on the outcome
            ok, both = 1, None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                both = symm_mem.empty((2, *shape), dtype=torch.float32, device=self.device)
                self.symm = symm_mem.rendezvous(both, group=group if group is not None else dist.group.WORLD)
                self._words = symm_mem.empty((8,), dtype=torch.int32, device=self.device)
                self._words_hdl = symm_mem.rendezvous(self._words, group=group if group is not None else dist.group.WORLD)
            except Exception as e:  # noqa: BLE001
                if transport in ("symm", "fused"):
                    raise
                ok, self.symm, self._symm_error = 0, None, repr(e)[:200]
            flag = torch.tensor([ok], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.symm = None
        if world == 1 and transport == "fused":
            # single rank, periodic ring of one: both neighbours are this rank itself, so the "peer" buffers are the
            # local ones and the fused kernel mirrors its boundary planes into its own ghost planes.  Lets one GPU
            # run (and time) exactly the kernel the multi-GPU path uses.
            both = torch.zeros((2, *shape), dtype=torch.float32, device=self.device)
            self._words = torch.zeros((8,), dtype=torch.int32, device=self.device)
            self.symm = "self"
        self.epoch = 1
        if self.symm == "self":
            self.transport = "fused"
            self.bufs = [both[0], both[1]]
            self.peer_lo = self.peer_hi = both
            self._peer_lo_words = self._peer_hi_words = self._words
        elif self.symm is not None:
            self.transport = "symm" if transport == "symm" else "fused"
            both.zero_()
            self._words.zero_()
            self.bufs = [both[0], both[1]]
            lo, hi = (rank - 1) % world, (rank + 1) % world
            self.peer_lo = self.symm.get_buffer(lo, (2, *shape), torch.float32)
            self.peer_hi = self.symm.get_buffer(hi, (2, *shape), torch.float32)
            self._peer_lo_words = self._words_hdl.get_buffer(lo, (8,), torch.int32)
            self._peer_hi_words = self._words_hdl.get_buffer(hi, (8,), torch.int32)
        else:
            self.transport = "nccl" if world > 1 else "local"
            self.bufs = [torch.zeros(shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.cur = 0
        # Small slabs (cfg4-class): the persistent rollout kernel exchanges 2K ghost planes every K time steps; it works
        # in four "wide" buffers [2][nz + 4K][H][W] (peer-mapped like the state buffers).  PERCNN_SLAB_TB_K=1 keeps
        # the per-step hand-shake of round 2's first persistent kernel.
        self._wide = None
        if self.transport == "fused" and self.plan.slab_persistent:
            # measured: 16 planes of 128^2 per rank on 8 GPUs want K = 6 (8.4 us/step; 9.1 at K = 4, 8.5 at K = 8, 19 with
            # a hand-shake per step -- profiles/r02_cfg4_8gpu_time_blocking.txt); on a ring of one, 32 planes want
            # K = 2-3 (profiles/r02_slab_small_time_blocking.txt): the redundant planes of a long block cost more the
            # deeper the slab
            # (over real NVLink the hand-shake is ~2x the ring-of-one's, so deeper slabs also want K = 4: 32 planes per rank on
            # 4 GPUs 12.0 us/step at K = 2, 64 planes on 2 GPUs 13.4 at K = 2 vs 13.0 at K = 4)
            k_default = 6 if self.nz <= 16 else 4
            k = max(1, min(int(os.environ.get("PERCNN_SLAB_TB_K", str(k_default))), self.nz // 2))
            if k >= 2:
                wshape = (4, 2, self.nz + 4 * k, H, W)
                if self.symm == "self":
                    wide = torch.zeros(wshape, dtype=torch.float32, device=self.device)
                    wlo = whi = wide
                else:
                    import torch.distributed._symmetric_memory as symm_mem
                    wide = symm_mem.empty(wshape, dtype=torch.float32, device=self.device)
                    whdl = symm_mem.rendezvous(wide, group=group if group is not None else dist.group.WORLD)
                    wide.zero_()
                    wlo = whdl.get_buffer((rank - 1) % world, wshape, torch.float32)
                    whi = whdl.get_buffer((rank + 1) % world, wshape, torch.float32)
                    self._wide_hdl = whdl
                self._wide = (wide, wlo, whi, k)
        self.comm_stream = torch.cuda.Stream(self.device)
        self.ev_boundary = torch.cuda.Event()
        self.ev_ghosts = torch.cuda.Event()
        self.use_graph = use_graph and (self.transport == "symm" or world == 1)
        self._graphs = {}
        self._extra_launches = 0

    # -- state ----------------------------------------------------------------------------------
    def set_state_from_host(self, host: torch.Tensor) -> None:
        """Like set_state, from a (pinned) HOST tensor [2][nz][H][W]: each field's planes are one contiguous block of the
        slab buffer, so the upload is two asynchronous copies straight into place (no staging tensor, no second pass)."""
        b = self.bufs[self.cur]
        for f in range(2):
            b[f, 2:self.nz + 2].copy_(host[f], non_blocking=True)
        self._publish_state()

    def interior_to_host(self, host: torch.Tensor) -> None:
        """The current slab into a (pinned) host tensor [2][nz][H][W]: two asynchronous copies; synchronise before reading."""
        b = self.bufs[self.cur]
        for f in range(2):
            host[f].copy_(b[f, 2:self.nz + 2], non_blocking=True)

    def set_state(self, interior: torch.Tensor) -> None:
        """interior: [2][nz][H][W] slab of the global field (planes z0 .. z0+nz)."""
        b = self.bufs[self.cur]
        b[:, 2:self.nz + 2].copy_(interior)
        self._publish_state()

    def _publish_state(self) -> None:
        self._exchange_blocking(self.cur)
        if self.transport == "fused":
            # ghosts of the current buffer are valid up to the current epoch on every rank
            self._words[0:2].fill_(self.epoch)
            self._words[2:5].zero_()
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier(self.group)

    def interior(self) -> torch.Tensor:
        return self.bufs[self.cur][:, 2:self.nz + 2]

    # -- sharded initial-state generator ------------------------------------------
    def set_state_from_upscaler(self, upscaler, init_state_low: torch.Tensor) -> None:
        """h0 = upscaler(init_state_low) (GS3D:186), every rank producing only its own planes, straight into the slab
        buffer (no full-resolution tensor exists anywhere); `init_state_low` is the whole low-resolution input,
        replicated (1/8 of the cells).  Keeps what `upscaler_backward` needs."""
        from . import upscaler as up
        H, W = self.plan.spatial[1:]
        low = init_state_low.detach().to(self.device, torch.float32).contiguous()
        geo = upscaler.geometry(tuple(low.shape[2:]), torch.float32, self.device, out_z0=self.z0, out_nz=self.nz,
                                out_field_stride=(self.nz + 4) * H * W)
        if tuple(geo.out_shape) != (self.nz * self.world, H, W):
            raise ValueError(f"upscaler output {geo.out_shape} does not match the global grid {(self.nz * self.world, H, W)}")
        flat = up._pack(upscaler.up_parameters(), torch.float32)
        b = self.bufs[self.cur]
        _, mid = up.upscaler_fwd(geo, flat, low, out=b[:, 2:])
        self._up = (upscaler, geo, flat, low, mid)
        self.set_state(b[:, 2:self.nz + 2])

    def upscaler_backward(self, g_h0: torch.Tensor) -> torch.Tensor:
        """Parameter gradient of the upscaler for dL/dh0 = `g_h0` (this rank's [2][nz][H][W] planes, as `backward`
        returns them): ghost planes of the gradient are exchanged once, every rank reduces the sums of its own planes,
        one all-reduce adds them.  Returns the flat gradient (packing order), identical on every rank."""
        from . import upscaler as up
        upscaler, geo, flat, low, mid = self._up
        gb = self.bufs[self.cur ^ 1]
        gb[:, 2:self.nz + 2].copy_(g_h0)
        self._exchange_blocking(self.cur ^ 1)
        gp = up.upscaler_bwd(geo, flat, low, mid, gb[:, 2:].data_ptr())
        if self.world > 1:
            dist.all_reduce(gp, op=dist.ReduceOp.SUM, group=self.group)
        return gp

    @property
    def launch_count(self) -> int:
        return self.plan.launch_count + self._extra_launches

    def describe(self):
        H, W = self.plan.spatial[1:]
        overlap = {
            "fused": "one kernel per step: boundary planes are stored straight into the neighbours' ghost planes over NVLink "
                     "from inside the z-march, flags raised/awaited in-kernel, march direction alternating per step; "
                     "whole rollout issued by one C call (no second stream, no NCCL on the data path)",
            "symm": "boundary kernels first, peer copies + stream signals on a second stream under the interior kernel",
            "nccl": "boundary kernels first, grouped ncclSend/ncclRecv on a second stream under the interior kernel",
            "local": "single rank: ghost planes are a local wrap copy",
        }[self.transport]
        d = {"transport": self.transport, "planes_per_rank": self.nz, "ghost_bytes_per_side_per_step": 2 * 2 * H * W * 4,
             "cuda_graph": bool(self.use_graph), "overlap": overlap}
        if self._wide is not None:
            k = self._wide[3]
            d["time_blocking"] = {"steps_per_exchange": k, "ghost_planes_per_side": 2 * k,
                                  "note": "persistent small-slab kernel: 2K ghost planes exchanged every K steps, K sub-steps on "
                                          "shrinking plane ranges between hand-shakes (one cooperative launch per rollout)"}
        return d

    # -- exchange -------------------------------------------------------------------------------
    def _exchange_blocking(self, b: int) -> None:
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(self.group)
        if self.symm is not None:
            nz = self.nz
            me = self.bufs[b]
            self.peer_lo[b][:, nz + 2:nz + 4].copy_(me[:, 2:4])
            self.peer_hi[b][:, 0:2].copy_(me[:, nz:nz + 2])
        else:
            for w in exchange_ghosts(self.bufs[b], self.nz, self.rank, self.world, self.group):
                w.wait()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(self.group)

    def _symm_push(self, b: int) -> None:
        """Copy my boundary planes of buffer b into the neighbours' ghost planes, then raise their signals."""
        nz = self.nz
        me = self.bufs[b]
        self.peer_lo[b][:, nz + 2:nz + 4].copy_(me[:, 2:4])       # my top planes = lower neighbour's upper ghosts
        self.peer_hi[b][:, 0:2].copy_(me[:, nz:nz + 2])           # my bottom planes = upper neighbour's lower ghosts
        lo, hi = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        self.symm.put_signal(lo, channel=0)
        self.symm.put_signal(hi, channel=1)
        self._extra_launches += 4

    def _symm_wait(self) -> None:
        lo, hi = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        self.symm.wait_signal(hi, channel=0)
        self.symm.wait_signal(lo, channel=1)
        self._extra_launches += 2

    # -- stepping -------------------------------------------------------------------------------
    def _step(self, first: bool) -> None:
        cur, nxt = self.bufs[self.cur], self.bufs[self.cur ^ 1]
        nz = self.nz
        main = torch.cuda.current_stream(self.device)
        if self.world > 1 and not first:        # ghosts of `cur` were pushed during the previous step
            if self.symm is not None:
                self._symm_wait()
            else:
                main.wait_event(self.ev_ghosts)
        self.plan.step_fwd_range(cur, nxt, 0, 2)
        self.plan.step_fwd_range(cur, nxt, nz - 2, nz)
        if self.world == 1:
            self.plan.step_fwd_range(cur, nxt, 2, nz - 2)
            exchange_ghosts(nxt, nz, 0, 1)
        else:
            self.ev_boundary.record(main)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(self.ev_boundary)
                if self.symm is not None:
                    self._symm_push(self.cur ^ 1)
                else:
                    for w in exchange_ghosts(nxt, nz, self.rank, self.world, self.group):
                        w.wait()                 # stream-level wait: comm_stream now depends on the NCCL stream
                    self.ev_ghosts.record(self.comm_stream)
            self.plan.step_fwd_range(cur, nxt, 2, nz - 2)
        self.cur ^= 1

    def _ring(self):
        """percnn_slab_ring_t over the two peer-mapped ping-pong buffers."""
        from ._lib import SlabRing
        r = SlabRing()
        for i in range(2):
            r.buf[i] = self.bufs[i].data_ptr()
            r.peer_lo_buf[i] = self.peer_lo[i].data_ptr()
            r.peer_hi_buf[i] = self.peer_hi[i].data_ptr()
        r.my_flags = self._words.data_ptr()
        r.peer_lo_flags = self._peer_lo_words.data_ptr()
        r.peer_hi_flags = self._peer_hi_words.data_ptr()
        r.scratch = self._words.data_ptr() + 8
        return r

    def _wide_struct(self):
        from ._lib import SlabWide
        wide, wlo, whi, k = self._wide
        w = SlabWide()
        for i in range(4):
            w.buf[i] = wide[i].data_ptr()
            w.peer_lo_buf[i] = wlo[i].data_ptr()
            w.peer_hi_buf[i] = whi[i].data_ptr()
        w.k = k
        return w

    def refresh_params(self) -> None:
        """Re-read the cell's parameters (call after an optimiser step)."""
        self.flat = engine.pack_params(self.cell._packed_tensors(), torch.float32)
        self.plan.params_load(self.flat)

    def rollout_tape(self, nsteps: int) -> torch.Tensor:
        """Forward rollout that keeps every state: returns [nsteps+1, 2, nz+4, H, W] (ghosted slabs, slot 0 = the
        current state).  The tape lives in peer-mapped memory so that each step's kernel can store its boundary
        planes straight into the neighbours' tape slot (fused transport only)."""
        if self.transport != "fused":
            raise NotImplementedError("rollout_tape needs the fused peer-memory transport")
        self.refresh_params()
        shape = (nsteps + 1, *self.plan.buffer_shape)
        if getattr(self, "_tape_shape", None) != shape:
            if self.symm == "self":
                self._tape = torch.empty(shape, dtype=torch.float32, device=self.device)
                self._tape_lo = self._tape_hi = self._tape
            else:
                import torch.distributed._symmetric_memory as symm_mem
                self._tape = symm_mem.empty(shape, dtype=torch.float32, device=self.device)
                hdl = symm_mem.rendezvous(self._tape, group=self.group if self.group is not None else dist.group.WORLD)
                lo, hi = (self.rank - 1) % self.world, (self.rank + 1) % self.world
                self._tape_lo = hdl.get_buffer(lo, shape, torch.float32)
                self._tape_hi = hdl.get_buffer(hi, shape, torch.float32)
            self._tape_shape = shape
        tape = self._tape
        tape[0].copy_(self.bufs[self.cur])
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(self.group)      # every rank's slot 0 (incl. ghosts) is in place before anyone mirrors into the tape
        self.plan.slab_rollout_tape(tape, self._tape_lo, self._tape_hi, self._ring(), nsteps, self.epoch)
        self.epoch += nsteps
        self.bufs[self.cur].copy_(tape[nsteps])
        return tape

    def _loss_spec(self, nsteps: int, sel, stride: int) -> "engine.DataLossSpec":
        D, (H, W) = self.nz * self.world, self.plan.spatial[1:]
        return slab_loss_spec((D, H, W), self.z0, self.nz, nsteps, sel, stride)

    def data_loss(self, tape: torch.Tensor, target_sub: torch.Tensor, sel, stride: int) -> torch.Tensor:
        """Fused data loss over the whole (global) grid: `mse_loss(states[sel][:, :, ::s, ::s, ::s], truth_sub)` with
        `target_sub` this rank's slab of the low-res truth, [nsel, 2, nz/s, ceil(H/s), ceil(W/s)].  Each rank reduces
        its sampled points (percnn_data_loss_fwd), one scalar all-reduce adds them up."""
        nsteps = tape.shape[0] - 1
        spec = self._loss_spec(nsteps, sel, stride)
        part = self.plan.data_loss_fwd(tape, nsteps, spec, target_sub)
        if self.world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return part

    def backward(self, tape: torch.Tensor, g_tape: Optional[torch.Tensor] = None, *, loss=None):
        """Back-propagate through the taped rollout.  g_tape[t] = dL/d(tape[t]) in the same ghosted layout (what
        autograd returns for a loss computed on tape[:, :, 2:-2]); ghost entries are ignored.  `loss` =
        (target_sub, sel, stride, gscale) adds the fused data loss of `data_loss()` as a gradient source: its
        gradient is injected inside the adjoint kernel of each selected step, so cfg5-sized training needs no dense
        gradient tape at all (g_tape=None).
        Returns (dL/dh0 interior [2, nz, H, W], flat parameter gradient summed over all ranks)."""
        if self.transport != "fused":
            raise NotImplementedError("backward needs the fused peer-memory transport")
        nsteps = tape.shape[0] - 1
        nz = self.nz
        plan = self.plan
        spec = target_sub = gscale = None
        if loss is not None:
            target_sub, sel, stride, gscale = loss
            spec = self._loss_spec(nsteps, sel, stride)
            if gscale is not None:
                gscale = torch.as_tensor(gscale, dtype=torch.float32, device=self.device).reshape(1)
            want = (spec.nsel, *plan.lowres_shape(spec.stride))
            if tuple(target_sub.shape) != want or target_sub.dtype != torch.float32 or not target_sub.is_contiguous():
                raise ValueError(f"loss target: need contiguous float32 {want}, got {tuple(target_sub.shape)}")
        elif g_tape is None:
            raise ValueError("backward needs g_tape and/or loss")
        plan.params_load(self.flat)
        plan.param_grads_begin()
        # G_nsteps = dL/dh_nsteps with exchanged ghosts, in the peer-mapped ping-pong buffers
        b = 0
        self.bufs[b].zero_()
        if g_tape is not None:
            self.bufs[b][:, 2:nz + 2].copy_(g_tape[nsteps][:, 2:nz + 2])
        self._exchange_blocking(b)
        self._words[0:2].fill_(self.epoch)
        self._words[2:5].zero_()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(self.group)
        if g_tape is not None and (not g_tape.is_contiguous() or tuple(g_tape.shape) != tuple(tape.shape)):
            raise ValueError("g_tape must be a contiguous tensor of the tape's shape")
        plan.slab_rollout_bwd(tape, g_tape, spec, target_sub, gscale, self._ring(), nsteps, self.epoch)
        self.epoch += nsteps
        b = nsteps & 1
        g_h0 = self.bufs[b][:, 2:nz + 2].clone()
        if self.world > 1:
            sums = plan.reduction_sums()
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)     # one tiny all-reduce per backward pass
        grads = plan.param_grads_finish(self.flat)
        self.cur = 0   # the state buffers were used as gradient scratch: the caller must set_state() again
        return g_h0, grads

    def error_word(self) -> int:
        """Non-zero if a fused step gave up waiting for a neighbour (device-side spin deadline).  The kernel also
        TRAPS in that case, so normally the caller sees a CUDA error at its next synchronisation instead."""
        return int(self._words[3].item()) if self.transport == "fused" else 0

    def run(self, nsteps: int) -> None:
        """Advance the slab by nsteps time steps.  Invariant on entry and exit: the ghosts of the current
        buffer are valid and every exchange signal has been consumed (set_state() establishes it)."""
        if self.transport == "fused":
            if self._wide is not None and nsteps >= 2:
                self.plan.slab_rollout_fwd_blocked(self._ring(), self._wide_struct(), self.cur, nsteps, self.epoch)
            else:
                self.plan.slab_rollout_fwd(self._ring(), self.cur, nsteps, self.epoch)
            self.epoch += nsteps
            self.cur ^= nsteps & 1
            return
        for i in range(nsteps):
            self._step(first=(i == 0))
        if self.world > 1 and nsteps > 0:
            main = torch.cuda.current_stream(self.device)
            if self.symm is not None:
                main.wait_stream(self.comm_stream)
                self._symm_wait()
            else:
                main.wait_event(self.ev_ghosts)
