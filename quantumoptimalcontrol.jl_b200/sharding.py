"""Multi-GPU partitioning of the GRAPE evaluation: one process per GPU, torch.distributed (NCCL over NVLink on
GPUs, gloo in the CPU tests) for the plumbing.  Two natural axes:

  * pulse batch (multistart / parameter batch): pulses are block-partitioned over ranks, every rank evaluates its
    block with its own handle; NO data-path collective (results are gathered only if the caller asks).
  * one long pulse: contiguous time segments, slices [p*Nt/P, (p+1)*Nt/P) on rank p.
        phase 1 (local)   U_k, dU_k/du_j for the local slices and the rank propagator S_p = U_last ... U_first
        exchange          all-gather of the S_p  (P x 16 d^2 bytes)
        phase 2 (local, boundary algebra redundant on every rank)
                          x_start(p) = S_{p-1} ... S_0 x0,  x_N,  J,  lambda_N = dJ(x_N),
                          lambda_end(p) = S_{p+1}' ... S_{P-1}' lambda_N,
                          local forward / backward sweeps and the local gradient columns
        exchange          all-gather of the nc x Nt/P gradient segments
    This replaces the reference's serial loops src/gradient_computations.jl:27-29, :52-58, :65-74.
    With a running state penalty (src/penalty_fcns.jl:1-11) the costate recurrence is affine, lambda_start = S_p' lambda_end + c_p:
        phase 2a          local forward sweep from x_start(p); affine term c_p and the local sum_k L(x_k)
        exchange          all-gather of (c_p, sum_k L) -- the second exchange of SURVEY.md section 7
        phase 2b          lambda_end(p) walked down from the last rank through (S_q', c_q), local backward sweep

The compute is delegated to a *segment engine* (phase1 / forward / backward).  The product engine is
CudaSegmentEngine (C ABI, device pointers, no host round trip); the CPU tests plug in an oracle-backed engine to
exercise exactly this host logic with world_size 2 under gloo.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .grape import GrapeCache, QOCError, _BuiltinCost, COST_INFIDELITY, COST_ABS_TRACE, COST_ZCAL

__all__ = ["block_partition", "time_partition", "CudaSegmentEngine", "TimeShardedEvaluator", "evaluate_batch_sharded",
           "InProcessSharded"]


def block_partition(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def time_partition(nt: int, world: int, rank: int):
    """Slices [lo, hi) of a pulse of nt slices owned by `rank` ([p*Nt/P, (p+1)*Nt/P))."""
    return block_partition(nt, world, rank)


# ----------------------------------------------------------------------------------------------------------------------
# engines
# ----------------------------------------------------------------------------------------------------------------------
class CudaSegmentEngine:
    """Local time segment on one B200 through the C ABI (qoc_shard_*_device).  Tensors stay in HBM."""

    def __init__(self, A0, A, nt_local, m, device_index=0, order=0, penalty=None):
        self.d, self.m, self.nc, self.nt = A0.shape[0], m, len(A), nt_local
        self.device = torch.device("cuda", device_index)
        self.cache = GrapeCache(A0, np.zeros((self.d, m), dtype=np.complex128), (self.nc, nt_local), batch=1,
                                device=device_index, dUkdp_order=order, store_costates=False)
        pen = penalty[0] if isinstance(penalty, tuple) else penalty   # (L, dL_dx) of setup_state_penalty, or L alone
        self.cache._ensure(A0, A, np.zeros((self.d, m), dtype=np.complex128), pen)
        self.penalty = pen
        self.lib = _lib.load()

    def _check(self, rc):
        if rc != 0:
            raise QOCError(rc, self.lib.qoc_last_error(self.cache.handle).decode())

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr())

    def phase1(self, u_local):
        """u_local: (nc, nt_local) float64 (numpy or tensor).  -> S_p as a (d, d) complex128 tensor on the GPU."""
        u = torch.as_tensor(np.ascontiguousarray(np.asarray(u_local, dtype=np.float64).T)).to(self.device)  # [k][j]
        S_cm = torch.empty((self.d, self.d), dtype=torch.complex128, device=self.device)  # column-major memory
        self._check(self.lib.qoc_shard_phase1_device(self.cache.handle, self._ptr(u), self._ptr(S_cm),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self._u = u
        return S_cm.t()  # logical (row, col) view of the column-major buffer

    def set_builtin_cost(self, cost, x0):
        """Built-in cost + initial state for phase2 (the boundary algebra then runs on the device)."""
        self.cache._ensure_x0_cost(cost, x0)

    def phase2(self, S_all, nranks, rank):
        """S_all: (nranks, d, d) complex128 tensor of column-major rank propagators, as all-gathered from phase1.
        -> (J tensor (1,), local gradient (nt_local, nc) tensor), both on the GPU."""
        J = torch.empty(1, dtype=torch.float64, device=self.device)
        g = torch.empty((self.nt, self.nc), dtype=torch.float64, device=self.device)
        self._check(self.lib.qoc_shard_phase2_device(self.cache.handle, self._ptr(S_all), nranks, rank, self._ptr(J), self._ptr(g),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return J, g

    def phase1_cm(self, u_dev):
        """u_dev: (nt_local, nc) float64 tensor already on the GPU.  -> S_p as a column-major (d, d) buffer tensor."""
        S_cm = torch.empty((self.d, self.d), dtype=torch.complex128, device=self.device)
        self._check(self.lib.qoc_shard_phase1_device(self.cache.handle, self._ptr(u_dev), self._ptr(S_cm),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return S_cm

    def forward(self, x_start):
        xs = x_start.t().contiguous()  # (m, d) row-major == d x m column-major
        xe = torch.empty_like(xs)
        self._check(self.lib.qoc_shard_forward_device(self.cache.handle, self._ptr(xs), self._ptr(xe),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return xe.t()

    def affine(self):
        """Running penalty: (c_p as a (d, m) tensor, local sum_k L(x_k) as a (1,) tensor) after forward()."""
        c = torch.empty((self.m, self.d), dtype=torch.complex128, device=self.device)
        Jp = torch.empty(1, dtype=torch.float64, device=self.device)
        self._check(self.lib.qoc_shard_affine_device(self.cache.handle, self._ptr(c), self._ptr(Jp),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return c.t(), Jp

    def backward(self, lam_end):
        le = lam_end.t().contiguous()
        ls = torch.empty_like(le)
        g = torch.empty((self.nt, self.nc), dtype=torch.float64, device=self.device)
        self._check(self.lib.qoc_shard_backward_device(self.cache.handle, self._ptr(le), self._ptr(g), self._ptr(ls),
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return g.t(), ls.t()


# ----------------------------------------------------------------------------------------------------------------------
# one long pulse, time-segment sharded
# ----------------------------------------------------------------------------------------------------------------------
def _builtin_cost_torch(cost, x):
    """J and dJ_dx of the library's built-in costs on a torch tensor (boundary algebra; identical on every rank)."""
    T = torch.as_tensor(np.ascontiguousarray(cost.T), device=x.device)
    om = torch.sum(torch.conj(T) * x)
    if cost.kind == COST_INFIDELITY:  # src/penalty_fcns.jl:15-24
        n2 = float(cost.n) ** 2
        return float(1 - (om.real ** 2 + om.imag ** 2) / n2), (-2 * om / n2) * T
    if cost.kind == COST_ABS_TRACE:   # test/test_gradient_computation.jl:24-25
        a = torch.abs(om)
        return float(1 - a), -(om / a) * T
    if cost.kind == COST_ZCAL:        # src/penalty_fcns.jl:27-42 (scalar tail on four numbers: host formulas)
        xh = x.cpu().numpy()
        return float(cost.host_J(xh)), torch.as_tensor(np.asarray(cost.host_grad(xh), dtype=np.complex128), device=x.device)
    raise ValueError("unknown built-in cost")


class TimeShardedEvaluator:
    """Fidelity + gradient of ONE pulse whose Nt slices are split over the ranks of `group`."""

    def __init__(self, engine, x0, cost, nt_total, group=None):
        self.engine, self.cost, self.nt_total, self.group = engine, cost, nt_total, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = getattr(engine, "device", torch.device("cpu"))
        self.x0 = torch.as_tensor(np.asarray(x0, dtype=np.complex128).reshape(np.asarray(x0).shape[0], -1)).to(self.device)
        self.lo, self.hi = time_partition(nt_total, self.world, self.rank)

    def _all_gather(self, t):
        if self.world == 1:
            return [t]
        t = t.contiguous()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return out

    def _evaluate_device(self, u_full):
        """Built-in cost + CUDA engine: phase 1, ONE all-gather of the rank propagators, phase 2 (boundary algebra, boundary
        scan and sweeps on the device), ONE all-gather of the gradient segments, a single device -> host copy at the end."""
        eng = self.engine
        nc = u_full.shape[0]
        if not getattr(self, "_dev_ready", False):
            eng.set_builtin_cost(self.cost, self.x0.cpu().numpy())
            self._nmax = max(time_partition(self.nt_total, self.world, r)[1] - time_partition(self.nt_total, self.world, r)[0]
                             for r in range(self.world))
            self._S_all = torch.empty((self.world, eng.d, eng.d), dtype=torch.complex128, device=self.device)
            self._g_all = torch.zeros((self.world, self._nmax, nc), dtype=torch.float64, device=self.device)
            self._g_pad = torch.zeros((self._nmax, nc), dtype=torch.float64, device=self.device)
            # pinned staging: u goes up and (gradient segments, J) come down with asynchronous copies and ONE synchronise
            self._u_pin = torch.zeros((self.hi - self.lo, nc), dtype=torch.float64).pin_memory()
            self._u_dev = torch.zeros((self.hi - self.lo, nc), dtype=torch.float64, device=self.device)
            self._g_host = torch.zeros((self.world, self._nmax, nc), dtype=torch.float64).pin_memory()
            self._J_host = torch.zeros(1, dtype=torch.float64).pin_memory()
            self._dev_ready = True
        self._u_pin.copy_(torch.from_numpy(u_full[:, self.lo:self.hi].T))
        self._u_dev.copy_(self._u_pin, non_blocking=True)
        u_dev = self._u_dev
        S = eng.phase1_cm(u_dev)
        if self.world > 1:
            dist.all_gather_into_tensor(self._S_all, S, group=self.group)
        else:
            self._S_all[0].copy_(S)
        J, g_loc = eng.phase2(self._S_all, self.world, self.rank)
        self._g_pad[: self.hi - self.lo].copy_(g_loc)
        if self.world > 1:
            dist.all_gather_into_tensor(self._g_all, self._g_pad, group=self.group)
        else:
            self._g_all[0].copy_(self._g_pad)
        self._g_host.copy_(self._g_all, non_blocking=True)
        self._J_host.copy_(J[:1], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        gh = self._g_host.numpy()
        g = np.zeros((nc, self.nt_total))
        for r in range(self.world):
            lo, hi = time_partition(self.nt_total, self.world, r)
            g[:, lo:hi] = gh[r, : hi - lo].T
        return float(self._J_host[0]), g

    def evaluate(self, u_full):
        """u_full: (nc, Nt) on every rank (only the local columns are used).  -> (J, dJdu (nc, Nt) numpy)."""
        u_full = np.asarray(u_full, dtype=np.float64)
        pen = getattr(self.engine, "penalty", None)
        if pen is None and isinstance(self.cost, _BuiltinCost) and hasattr(self.engine, "phase2"):
            return self._evaluate_device(u_full)
        u_loc = u_full[:, self.lo:self.hi]
        # phase 1 + exchange of the boundary propagators
        S = self._all_gather(self.engine.phase1(u_loc))
        # boundary algebra (redundant on every rank; P small products of d x d by d x m)
        x = self.x0
        x_start, x_starts = None, []
        for p in range(self.world):
            if p == self.rank:
                x_start = x
            x_starts.append(x)
            x = S[p] @ x
        x_N = x
        if isinstance(self.cost, _BuiltinCost):
            J, lam = _builtin_cost_torch(self.cost, x_N)
        else:  # arbitrary host closures (J, dJ_dx) like the reference's
            Jf, dJf = self.cost
            xh = x_N.cpu().numpy()
            J = float(Jf(xh))
            lam = torch.as_tensor(np.asarray(dJf(xh), dtype=np.complex128)).to(self.device)
        if pen is None:
            lam_end = None
            for p in range(self.world - 1, -1, -1):
                if p == self.rank:
                    lam_end = lam
                lam = S[p].conj().t() @ lam
            # phase 2: local sweeps
            self.engine.forward(x_start)
        else:
            # phase 2a: forward sweep, then the affine term of the local costate recurrence and the local sum of L
            self.engine.forward(x_start)
            c_loc, Jp_loc = self.engine.affine()
            d, m = c_loc.shape
            pack = torch.cat([torch.view_as_real(c_loc.contiguous()).reshape(-1), Jp_loc.reshape(1).to(torch.float64)])
            packs = self._all_gather(pack)   # the second exchange: 2 d m + 1 doubles per rank
            cs = [torch.view_as_complex(q[:-1].reshape(d, m, 2)) for q in packs]
            rows = torch.as_tensor(np.asarray(pen.rows), device=self.device, dtype=torch.long)
            cols = torch.as_tensor(np.asarray(pen.cols), device=self.device, dtype=torch.long)
            # a boundary state x_start(q), q >= 1, is the last state of rank q - 1 and the first of rank q: L and dL_dx once
            J += float(sum(q[-1] for q in packs))
            for q in range(1, self.world):
                J -= float(pen.mu * torch.sum(torch.abs(x_starts[q][rows][:, cols]) ** 2))
            # lam = the costate entering a rank from the right BEFORE dL_dx of its last state (the engine adds that, as the
            # reference does for lambda_N: src/gradient_computations.jl:47-49)
            lam_end = None
            for p in range(self.world - 1, -1, -1):
                if p == self.rank:
                    lam_end = lam
                    break
                lam = S[p].conj().t() @ lam + cs[p]
                dl = torch.zeros_like(lam)
                dl[rows[:, None], cols[None, :]] = 2 * pen.mu * x_starts[p][rows][:, cols]
                lam = lam - dl
        g_loc, _ = self.engine.backward(lam_end)
        # gradient segments: ranks may own different numbers of slices -> pad to the longest
        nmax = max(time_partition(self.nt_total, self.world, r)[1] - time_partition(self.nt_total, self.world, r)[0]
                   for r in range(self.world))
        nc = u_full.shape[0]
        pad = torch.zeros((nc, nmax), dtype=torch.float64, device=self.device)
        pad[:, : self.hi - self.lo] = g_loc
        parts = self._all_gather(pad)
        g = np.zeros((nc, self.nt_total))
        for r in range(self.world):
            lo, hi = time_partition(self.nt_total, self.world, r)
            g[:, lo:hi] = parts[r][:, : hi - lo].cpu().numpy()
        return J, g


# ----------------------------------------------------------------------------------------------------------------------
# independent pulses, batch sharded (no data-path collective)
# ----------------------------------------------------------------------------------------------------------------------
def evaluate_batch_sharded(eval_local, u_batch, group=None, gather=True):
    """u_batch: (B, nc, Nt) identical on every rank.  eval_local(u_block) -> (J (b,), dJdu (b, nc, Nt)) evaluates this
    rank's block.  With gather=True the per-rank results are all-gathered so every rank returns the full (J, dJdu)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = u_batch.shape[0]
    lo, hi = block_partition(B, world, rank)
    J_loc, g_loc = eval_local(u_batch[lo:hi]) if hi > lo else (np.zeros(0), np.zeros((0,) + u_batch.shape[1:]))
    if not gather or world == 1:
        return np.asarray(J_loc), np.asarray(g_loc), (lo, hi)
    nmax = max(block_partition(B, world, r)[1] - block_partition(B, world, r)[0] for r in range(world))
    padJ = torch.zeros(nmax, dtype=torch.float64)
    padg = torch.zeros((nmax,) + tuple(u_batch.shape[1:]), dtype=torch.float64)
    padJ[: hi - lo] = torch.as_tensor(np.asarray(J_loc, dtype=np.float64))
    padg[: hi - lo] = torch.as_tensor(np.asarray(g_loc, dtype=np.float64))
    backend = dist.get_backend(group)
    if backend == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        padJ, padg = padJ.to(dev), padg.to(dev)
    outJ = [torch.empty_like(padJ) for _ in range(world)]
    outg = [torch.empty_like(padg) for _ in range(world)]
    dist.all_gather(outJ, padJ, group=group)
    dist.all_gather(outg, padg, group=group)
    J = np.zeros(B)
    g = np.zeros(u_batch.shape)
    for r in range(world):
        l, h = block_partition(B, world, r)
        J[l:h] = outJ[r][: h - l].cpu().numpy()
        g[l:h] = outg[r][: h - l].cpu().numpy()
    return J, g, (lo, hi)


# ----------------------------------------------------------------------------------------------------------------------
# one process, several GPUs: the library drives the devices itself (qoc_create_sharded / qoc_sharded_eval)
# ----------------------------------------------------------------------------------------------------------------------
class InProcessSharded:
    """The multi-GPU evaluation a host without torchrun (a Julia process, a C program) uses: ONE process, the C library
    owns a stream per device, exchanges the rank propagators through NVLink peer stores and writes every rank's results
    into the caller's host arrays.  kind: "batch" (pulses block-partitioned, no exchange) or "time" (one pulse, contiguous
    time segments).  devices: CUDA ordinals, one per rank (an ordinal may repeat: virtual ranks on one GPU)."""

    def __init__(self, A0, A, x0, cost, u_size, devices, kind="time", batch=1, dUkdp_order=0, penalty=None):
        from .grape import _c128, _dptr
        if not isinstance(cost, _BuiltinCost):
            raise QOCError(_lib.ERR_INVALID, "the in-library sharded evaluation needs a built-in cost")
        lib = _lib.load()
        self.lib = lib
        A0c = _c128(A0)
        Ac = np.stack([_c128(a) for a in A], axis=0)
        x0c = _c128(np.asarray(x0).reshape(np.asarray(x0).shape[0], -1))
        Aflat = np.ascontiguousarray(np.stack([a.T for a in Ac], axis=0))
        pr = _lib.Problem()
        pr.d, pr.m, pr.nc, pr.nt, pr.batch = A0c.shape[0], x0c.shape[1], len(A), int(u_size[1]), int(batch)
        pr.order, pr.cost, pr.n, pr.device = int(dUkdp_order), cost.kind, cost.n, 0
        self.nc, self.nt, self.batch, self.kind = pr.nc, pr.nt, pr.batch, kind
        pen = penalty[0] if isinstance(penalty, tuple) else penalty   # setup_state_penalty(...): J then includes sum_k L(x_k)
        if pen is not None and len(pen.rows) and len(pen.cols):
            rows = np.ascontiguousarray(pen.rows, dtype=np.int32)
            cols = np.ascontiguousarray(pen.cols, dtype=np.int32)
            pr.n_pen_rows, pr.n_pen_cols = len(rows), len(cols)
            pr.pen_rows = rows.ctypes.data_as(C.POINTER(C.c_int32))
            pr.pen_cols = cols.ctypes.data_as(C.POINTER(C.c_int32))
            pr.mu = pen.mu
        devs = (C.c_int * len(devices))(*[int(x) for x in devices])
        h = C.c_void_p()
        rc = lib.qoc_create_sharded(C.byref(pr), _dptr(A0c), _dptr(Aflat), _dptr(x0c), _dptr(cost.T), len(devices), devs,
                                    1 if kind == "time" else 0, C.byref(h))
        if rc != _lib.OK:
            raise QOCError(rc, lib.qoc_sharded_last_error(None).decode())
        self._h = h

    def evaluate(self, u):
        """u: (nc, Nt) or (batch, nc, Nt) -> (J, dJdu) with the shapes of qoc_b200.evaluate."""
        from .grape import _dptr
        u = np.asarray(u, dtype=np.float64)
        if self.batch == 1:
            uu = np.ascontiguousarray(u.T)
        else:
            uu = np.ascontiguousarray(np.transpose(u, (0, 2, 1)))
        J = np.zeros(self.batch)
        g = np.zeros(uu.shape)
        rc = self.lib.qoc_sharded_eval(self._h, _dptr(uu), _dptr(J), _dptr(g))
        if rc != _lib.OK:
            raise QOCError(rc, self.lib.qoc_sharded_last_error(self._h).decode())
        if self.batch == 1:
            return float(J[0]), g.T.copy()
        return J, np.transpose(g, (0, 2, 1)).copy()

    def last_ms(self):
        return self.lib.qoc_sharded_last_ms(self._h)

    def close(self):
        if self._h is not None:
            self.lib.qoc_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
