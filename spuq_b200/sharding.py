"""Multi-GPU form of the SG operator application (BASELINE.json north_star): the M expansion terms
are sharded over the ranks, every rank applies its terms to a full replica of w, and the partial
results are summed with a reduce-scatter along the Lambda (column) axis, so that rank r ends up
owning the r-th contiguous block of output columns.

The reference has no multi-process code at all; the split is the sum over m
of application/egsz/multi_operator2.py:59-82, which is independent per m.
"""
import collections
import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist

from .multi_operator import MeanSolveSlabOperator, MultiOperator
from .multi_vector import _axpy, _dot, _scal
from .pcg import PCG_MESSAGES
from . import _abi

__all__ = ["shard_ranges", "padded_columns", "ShardedApply", "row_ranges", "RowShardedApply", "RowShardedPCG",
           "PartitionedMeanSolver", "StripRowsMeanSolver"]


def shard_ranges(tables, M, world):
    """Contiguous m-ranges per rank, balanced by term count (the L terms of the mean block go to
    rank 0).  Deterministic, so that every rank computes the same split."""
    w = np.array([(tables.nbr_up[m] >= 0).sum() + (tables.nbr_dn[m] >= 0).sum()
                  + (tables.beta_0[m] != 0).sum() for m in range(M)], dtype=np.float64)
    total = w.sum() + tables.L
    target = total / world
    ranges, lo, acc = [], 0, float(tables.L)
    for g in range(world):
        hi = lo
        if g == world - 1:
            hi = M
        else:
            while hi < M and acc + w[hi] / 2 <= target * (g + 1):
                acc += w[hi]
                hi += 1
        ranges.append((lo, hi))
        lo = hi
    return ranges


def padded_columns(L, world):
    """Columns per rank after padding L up to a multiple of the world size."""
    return (L + world - 1) // world


class ShardedApply:
    """y_own = (A w)[columns of this rank] with A = sum over ranks of the rank's term shard.

    ``blocks`` / ``rvs`` as for MultiOperator.from_blocks; ``tables`` the LambdaTables of the full
    index set.  ``apply(W, Yfull, Yown)``: W is the L x N slab (full replica), Yfull an Lp x N work
    slab (Lp = world * padded_columns), Yown receives this rank's Lp/world columns."""

    def __init__(self, blocks, rvs, tables, group=None, path=0):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        M = tables.maxm
        self.ranges = shard_ranges(tables, M, self.world)
        self.op = MultiOperator.from_blocks(blocks, rvs, m_range=self.ranges[self.rank],
                                            include_mean=(self.rank == 0), path=path)
        self.L = tables.L
        self.cols = padded_columns(self.L, self.world)

    def plan_for(self, w):
        return self.op.plan_for(w)

    def apply(self, plan, W, Yfull, Yown):
        assert Yfull.shape[0] == self.cols * self.world and Yown.shape[0] == self.cols
        plan.apply(W, Yfull[:self.L])
        if self.world == 1:
            Yown.copy_(Yfull)
            return Yown
        if dist.get_backend(self.group) == "nccl":
            dist.reduce_scatter_tensor(Yown, Yfull, op=dist.ReduceOp.SUM, group=self.group)
        else:
            # gloo (CPU tests) has no reduce-scatter on tensors: all-reduce and keep the own block
            tmp = Yfull.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            Yown.copy_(tmp[self.rank * self.cols:(self.rank + 1) * self.cols])
        return Yown


# ---- row sharding ---------------------------------------------------------------------------------
def row_ranges(ny, world, align=8):
    """Contiguous grid-row ranges per rank, multiples of `align` rows (the kernel's tile height)
    except for the last one."""
    per = -(-ny // world)
    per = -(-per // align) * align
    return [(min(r * per, ny), min((r + 1) * per, ny)) for r in range(world)]


class PeerUnavailable(RuntimeError):
    """Raised on ALL ranks (the decision is taken on gathered data) when some rank cannot export or map a slab."""


class PeerHaloExchange:
    """Halo rows of row-sharded slabs written straight into the neighbours' memory (one NVSwitch box).

    Every rank maps its neighbours' slabs into its own device context through CUDA IPC (handles exported and
    opened by the library, with the accessing device current) and, per exchange,
    enqueues on its own stream ONE small kernel that stores its first / last owned grid row of all columns
    into the neighbours' halo rows (coalesced stores over NVLink, no staging buffer) and then publishes a
    sequence number in the neighbours' flag words, followed by a one-thread wait for the two numbers the
    neighbours publish in its own flags.  A second pair of flags is the READY-TO-RECEIVE handshake: the push
    kernel first tells both neighbours that everything enqueued before it on this rank's stream is complete
    (the kernel that read the previous halo rows, and any update that swept over the slab since), and stores
    nothing until it has the same word from them -- a writer can neither overtake the reader of the previous
    exchange nor be overwritten by the owner's own later sweep over the slab.  No NCCL call, no host
    synchronisation: two tiny launches instead of six staging copies and a send/receive kernel.  (A first
    version used two cudaMemcpy2DAsync per exchange: ~1000 rows of 8 KB at a 4 MB pitch cost the copy engine
    1 ms.)

    flags (int64, per rank): [0] top halo written by the rank above, [1] bottom halo written by the rank
    below, [2] the rank above is ready to receive exchange k, [3] same for the rank below."""

    MAX_SLABS = 8

    def __init__(self, rank, world, device, group=None):
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.flags = torch.zeros(4, dtype=torch.int64, device=device)
        self.ticket = torch.zeros(2, dtype=torch.int32, device=device)
        self.step = 0
        self._slabs = collections.OrderedDict()
        self._mapped = {}                      # exported handle -> base address in this device's context
        handles = [None] * world
        dist.all_gather_object(handles, (device.index, self._try_export(self.flags)), group=group)
        if any(h is None for _, h in handles):
            raise PeerUnavailable("a rank could not export its flag words through CUDA IPC")
        self.peer_flags = {}
        lib = _abi.lib()
        err = None
        try:
            for r in (rank - 1, rank + 1):
                if 0 <= r < world:
                    dev_r, h = handles[r]
                    _abi.check(lib.spuq_sg_enable_peer_access(device.index, dev_r), "enable_peer_access")
                    self.peer_flags[r] = self._map(h)
        except _abi.SpuqError as e:
            err = e
        self._agree(err is None, "a rank could not map its neighbours' flag words (%r)" % (err,))

    def _agree(self, ok, what):
        """Collective: all ranks continue only if all of them succeeded."""
        oks = [None] * self.world
        dist.all_gather_object(oks, bool(ok), group=self.group)
        if not all(oks):
            raise PeerUnavailable(what)

    def _try_export(self, t):
        try:
            return self._export(t)
        except _abi.SpuqError:
            return None

    @staticmethod
    def _export(t):
        """(cudaIpcMemHandle bytes of the allocation that holds t, byte offset of t inside it)."""
        buf = C.create_string_buffer(64)
        off = C.c_int64()
        _abi.check(_abi.lib().spuq_sg_ipc_export(_abi.ptr(t), buf, C.byref(off)), "ipc_export")
        return (buf.raw, int(off.value))

    def _map(self, exported):
        """Device address (in this device's context) of a tensor another rank exported."""
        handle, offset = exported
        base = self._mapped.get(handle)
        if base is None:
            p = C.c_void_p()
            _abi.check(_abi.lib().spuq_sg_ipc_open(self.device.index, handle, C.byref(p)), "ipc_open")
            base = self._mapped[handle] = int(p.value)
        return base + offset

    def _peers_of(self, W, nyl, nx):
        """Neighbours' addresses of the slab that corresponds to W (registered collectively on first use)."""
        key = (W.data_ptr(), tuple(W.shape))
        if key in self._slabs:
            self._slabs.move_to_end(key)
        else:
            info = [None] * self.world
            dist.all_gather_object(info, (self._try_export(W), W.shape[0], nyl), group=self.group)
            if any(e is None for e, _, _ in info):
                raise PeerUnavailable("a rank could not export its slab through CUDA IPC")
            peers, err = {}, None
            try:
                for r in (self.rank - 1, self.rank + 1):
                    if 0 <= r < self.world:
                        exported, L_r, nyl_r = info[r]
                        assert L_r == W.shape[0]
                        peers[r] = (self._map(exported), nyl_r)
            except _abi.SpuqError as e:
                err = e
            self._agree(err is None, "a rank could not map a neighbour's slab (%r)" % (err,))
            # the entry keeps the tensor ALIVE: as long as an address is in the table it cannot be handed to
            # another tensor, here or (same registration, same order of calls) on the neighbours, so a hit never
            # names memory that has moved.  Bounded: the oldest registration goes when a ninth arrives.
            self._slabs[key] = (peers, W)
            while len(self._slabs) > self.MAX_SLABS:
                self._slabs.popitem(last=False)
        return self._slabs[key][0]

    def exchange(self, W, nyl, nx, stream):
        """Enqueue the exchange of the halo rows of W (L x nyl*nx) on `stream`; returns after enqueueing."""
        lib = _abi.lib()
        peers = self._peers_of(W, nyl, nx)
        self.step += 1
        k = self.step
        up, dn = self.rank - 1, self.rank + 1
        up_slab, up_nyl = (C.c_void_p(peers[up][0]), peers[up][1]) if up in peers else (None, 0)
        dn_slab, dn_nyl = (C.c_void_p(peers[dn][0]), peers[dn][1]) if dn in peers else (None, 0)
        up_fl = C.c_void_p(self.peer_flags[up]) if up in peers else None
        dn_fl = C.c_void_p(self.peer_flags[dn]) if dn in peers else None
        _abi.check(lib.spuq_sg_halo_push(_abi.ptr(W), W.shape[0], nyl, nx, up_slab, up_nyl, dn_slab, dn_nyl,
                                         _abi.ptr(self.flags), up_fl, dn_fl, k, _abi.ptr(self.ticket), stream), "halo_push")
        _abi.check(lib.spuq_sg_halo_wait(_abi.ptr(self.flags), int(up in peers), int(dn in peers), k, stream), "halo_wait")

class RowShardedApply:
    """The SG operator with the SPATIAL rows sharded over the ranks (axis 3).

    Rank r owns the grid rows [y0, y1) for all |Lambda| columns; its local slab carries one halo grid
    row above and below (local grid (y1 - y0 + 2) x nx).  One apply = exchange of the two boundary
    rows with the neighbouring ranks (2 x nx x L doubles per neighbour, point to point) followed by
    the single-GPU kernel on the local grid; output halo rows are not meaningful.  No reduction of
    Y is needed, so the traffic per rank is 1/world of the single-GPU traffic plus the halos.

    ``blocks``: local blocks from synthetic.fd_blocks_rows(n, M, y0, y1) (or any assembly of the same
    rows); ``W`` slabs are L x ((y1-y0+2)*nx)."""

    def __init__(self, blocks, rvs, nx, y0, y1, group=None, path=0):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nx, self.y0, self.y1 = nx, y0, y1
        self.nyl = y1 - y0 + 2
        self.op = MultiOperator.from_blocks(blocks, rvs, path=path)
        self._bufs = None
        self._side = None
        self._peer = None
        self.halo = "auto"               # "peer": CUDA-IPC copies; "nccl": send/recv; "auto": peer on NCCL groups
        self.last_launches = 0

    def plan_for(self, w):
        return self.op.plan_for(w)

    def exchange_halos(self, W):
        """Fill the halo rows of the local slab W (L x nyl*nx) from the neighbours' boundary rows."""
        if self.world == 1:
            return
        L = W.shape[0]
        W3 = W.view(L, self.nyl, self.nx)
        if self._bufs is None or self._bufs[0].shape[0] != L:
            mk = lambda: torch.empty((L, self.nx), dtype=W.dtype, device=W.device)      # noqa: E731
            self._bufs = (mk(), mk(), mk(), mk())
        s_up, s_dn, r_up, r_dn = self._bufs
        ops = []
        if self.rank > 0:
            s_up.copy_(W3[:, 1, :])                              # my first owned row -> bottom halo of rank-1
            ops += [dist.P2POp(dist.isend, s_up, self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, r_up, self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            s_dn.copy_(W3[:, self.nyl - 2, :])                   # my last owned row -> top halo of rank+1
            ops += [dist.P2POp(dist.isend, s_dn, self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, r_dn, self.rank + 1, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if self.rank > 0:
            W3[:, 0, :].copy_(r_up)
        if self.rank < self.world - 1:
            W3[:, self.nyl - 1, :].copy_(r_dn)

    def _use_peer(self, W):
        if self.world == 1 or W.device.type != "cuda" or self.halo == "nccl":
            return False
        if self.halo == "peer":
            return True
        return dist.get_backend(self.group) == "nccl"

    def apply(self, plan, W, Y, overlap=False):
        """Y = A W on the local slab.

        Default on the GPUs of one box: the halo rows are written straight into the neighbours' slabs
        (PeerHaloExchange: two DMA copies per neighbour and a flag), then ONE kernel launch.  With
        ``halo = "nccl"`` the rows travel by NCCL send/recv through staging buffers; ``overlap`` then runs
        that exchange on a side stream while the kernel already works on the tile rows that read no halo
        row (spuq_sg_apply_part 1) and finishes the tile rows at the two ends afterwards (part 2)."""
        if self._use_peer(W):
            # set-up and the registration of a new slab are collective and fail on ALL ranks together
            # (PeerUnavailable: no CUDA IPC between the processes, no peer access): fall back to NCCL
            try:
                if self._peer is None:
                    self._peer = PeerHaloExchange(self.rank, self.world, W.device, self.group)
                st = _abi.stream_ptr()
                self._peer.exchange(W, self.nyl, self.nx, st)
            except PeerUnavailable as e:
                if self.halo == "peer":
                    raise
                self._peer_error, self.halo = repr(e), "nccl"
            else:
                plan.apply(W, Y)
                self.last_launches = plan.last_launches + 2        # + push, wait (one small kernel each)
                return Y
        if self.world == 1 or not overlap or W.device.type != "cuda":
            self.exchange_halos(W)
            plan.apply(W, Y)
            self.last_launches = plan.last_launches
            return Y
        cur = torch.cuda.current_stream(W.device)
        if self._side is None:
            self._side = torch.cuda.Stream(W.device)
            self._ev_ready, self._ev_halo = torch.cuda.Event(), torch.cuda.Event()
        self._ev_ready.record(cur)                      # W is complete on the current stream
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ev_ready)
            self.exchange_halos(W)                      # enqueued FIRST: the NCCL kernel gets its SMs before
            self._ev_halo.record(self._side)            # the persistent apply kernel fills the device
        plan.apply_part(W, Y, 1)
        n1 = plan.last_launches
        cur.wait_event(self._ev_halo)
        plan.apply_part(W, Y, 2)
        self.last_launches = n1 + plan.last_launches
        return Y


def _const_tridiag_solve(b, o, rhs):
    """Solve tridiag(o, b[k], o) x = rhs for all k at once (Thomas, vectorised over the last axis):
    b: [nk], rhs: [n, nk] -> [n, nk]."""
    n = rhs.shape[0]
    piv = np.empty_like(rhs)
    x = rhs.astype(np.float64).copy()
    piv[0] = b
    for y in range(1, n):
        piv[y] = b - o * o / piv[y - 1]
        x[y] -= o / piv[y - 1] * x[y - 1]
    x[n - 1] /= piv[n - 1]
    for y in range(n - 2, -1, -1):
        x[y] = (x[y] - o * x[y + 1]) / piv[y]
    return x


class PartitionedMeanSolver:
    """A_0^{-1} on ROW-sharded slabs without moving the slab: the mean block is diagonalised in x (rows
    are whole on every rank) and the tridiagonal systems along y, which cross the ranks, are solved by
    the partition method.  Rank p solves its own diagonal block for the right-hand side (library:
    spuq_sg_precond_mean_fd_forward on the local nx x n_p grid); the influence of the two neighbouring
    interface values enters through the block's first / last columns of its inverse (the spikes), which
    depend on the mode and the block size only and are set-up data.  The 2P interface values of every
    (mode, column) satisfy a 2P x 2P system whose matrix depends on the mode only: it is inverted at
    set-up, every rank all-gathers the 2 x nx x |Lambda| boundary values and evaluates the two rows it
    needs.  Per call: two streaming corrections and one small all-gather instead of two transpositions
    of the whole slab."""

    def __init__(self, nx, ny, ranges, rank, stencil, group=None):
        from .multi_operator import MeanSolveSlabOperator
        d, ox, oy = stencil
        self.nx, self.ny, self.ranges, self.rank, self.group = int(nx), int(ny), list(ranges), int(rank), group
        self.P = len(self.ranges)
        sizes = [r1 - r0 for r0, r1 in self.ranges]
        if ny < 2 or min(sizes) < 2:
            raise ValueError("PartitionedMeanSolver needs at least two grid rows on every rank")
        self.n_loc = sizes[rank]
        self.local = MeanSolveSlabOperator(None, 1, stencil=(nx, self.n_loc, d, ox, oy))
        order = np.empty(nx, dtype=np.int32)
        _abi.check(_abi.lib().spuq_sg_precond_mean_fd_mode_order(self.local.handle, order.ctypes.data_as(C.c_void_p)),
                   "mode_order")
        k = order.astype(np.float64)
        lam = 0.5 * d + 2.0 * ox * np.cos((k + 1.0) * math.pi / (nx + 1.0))       # eigenvalues of the x part
        b = 0.5 * d + lam                                                         # diagonal of the y systems
        P = self.P
        first, last = [], []                                                      # spikes at the block ends, all ranks
        mine = None
        for q, nq in enumerate(sizes):
            e0 = np.zeros((nq, nx)); e0[0] = 1.0
            e1 = np.zeros((nq, nx)); e1[nq - 1] = 1.0
            sl, sr = _const_tridiag_solve(b, oy, e0), _const_tridiag_solve(b, oy, e1)
            first.append((sl[0], sr[0])); last.append((sl[nq - 1], sr[nq - 1]))
            if q == rank:
                mine = (sl, sr)
        # reduced system in the unknowns (a_0, z_0, a_1, z_1, ...) = (first, last value of every block)
        Mred = np.zeros((nx, 2 * P, 2 * P))
        for q in range(P):
            for row, (sl_v, sr_v) in ((2 * q, first[q]), (2 * q + 1, last[q])):
                Mred[:, row, row] = 1.0
                if q > 0:
                    Mred[:, row, 2 * (q - 1) + 1] += oy * sl_v
                if q < P - 1:
                    Mred[:, row, 2 * (q + 1)] += oy * sr_v
        Minv = np.linalg.inv(Mred)
        dev = _abi.device()
        zero = np.zeros((nx, 2 * P))
        cz = Minv[:, 2 * (rank - 1) + 1, :] if rank > 0 else zero                 # z_{p-1} = cz . g
        ca = Minv[:, 2 * (rank + 1), :] if rank < P - 1 else zero                 # a_{p+1} = ca . g
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)          # noqa: E731
        self.cz, self.ca = up(cz.T), up(ca.T)                                     # [2P, nx]
        self.sl, self.sr = up(oy * mine[0]), up(oy * mine[1])                     # [n_loc, nx], coupling folded in
        self._work = None

    def solve(self, Rc, Sc):
        """Rc, Sc: contiguous L x (n_loc * nx) slabs of the owned rows; Sc = (A_0^{-1} R) restricted to them."""
        lib = _abi.lib()
        L = Rc.shape[0]
        N = self.n_loc * self.nx
        assert Rc.is_contiguous() and Sc.is_contiguous() and Rc.shape == (L, N) and Sc.shape == (L, N)
        need = int(lib.spuq_sg_precond_work_bytes(self.local.handle, N, L))
        if self._work is None or self._work.numel() * 8 < need:
            self._work = torch.empty(need // 8 + 1, dtype=torch.float64, device=Rc.device)
        T = torch.empty_like(Rc)
        _abi.check(lib.spuq_sg_precond_mean_fd_forward(self.local.handle, N, L, _abi.ptr(Rc), _abi.ptr(T),
                                                       _abi.ptr(self._work), _abi.stream_ptr()), "mean_fd_forward")
        T3 = T.view(L, self.n_loc, self.nx)
        mine = torch.stack([T3[:, 0, :], T3[:, self.n_loc - 1, :]]).contiguous()   # [2, L, nx]
        if self.P > 1:
            parts = [torch.empty_like(mine) for _ in range(self.P)]
            dist.all_gather(parts, mine, group=self.group)
            g = torch.cat(parts)                                                  # [2P, L, nx], unknown order
        else:
            g = mine
        zl = torch.einsum("jk,jlk->lk", self.cz, g).contiguous()                   # [L, nx]
        zr = torch.einsum("jk,jlk->lk", self.ca, g).contiguous()
        _abi.check(lib.spuq_sg_spike_correct(L, self.n_loc, self.nx, _abi.ptr(T), _abi.ptr(zl), _abi.ptr(zr),
                                             _abi.ptr(self.sl), _abi.ptr(self.sr), _abi.stream_ptr()), "spike_correct")
        _abi.check(lib.spuq_sg_precond_mean_fd_inverse(self.local.handle, N, L, _abi.ptr(T), _abi.ptr(Sc),
                                                       _abi.stream_ptr()), "mean_fd_inverse")
        return Sc


def _all_to_all(send, recv, rank, group):
    """send[q] -> rank q, recv[q] <- rank q (lists of contiguous tensors, possibly empty)."""
    world = len(send)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)
        return
    ops = []
    for q in range(world):
        if q == rank:
            recv[q].copy_(send[q])
            continue
        if send[q].numel():
            ops.append(dist.P2POp(dist.isend, send[q], q, group))
        if recv[q].numel():
            ops.append(dist.P2POp(dist.irecv, recv[q], q, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class StripRowsMeanSolver:
    """A_0^{-1} on ROW-sharded slabs with the strip solver of libspuq_sg (csrc/precond_strip.cuh): every
    rank solves the strips of its own band in shared memory; the last row of every band but the last is a
    separator row, and the separator system of the whole grid (a few per cent of the rows) is solved
    with its COLUMNS sharded over the ranks: an all-to-all moves the right-hand sides of the separator rows
    (every rank receives all separators of |Lambda| / world columns), every rank solves its share, and a
    second all-to-all returns to each rank the separator values next to its own strips.  Per call: one row
    of L x nx doubles to the previous rank and two all-to-alls of a few separator rows; no transposition of
    the slab, no full-length transform of the bulk, no host synchronisation.  ``schur="gather"`` solves the
    whole separator system redundantly on every rank instead (one all-gather)."""

    def __init__(self, nx, ranges, rank, stencil, group=None, schur="sharded"):
        d, ox, oy = stencil
        self.nx, self.ranges, self.rank, self.group = int(nx), list(ranges), int(rank), group
        self.world = len(self.ranges)
        for (a0, a1), (b0, b1) in zip(self.ranges[:-1], self.ranges[1:]):
            assert a1 == b0, "row ranges must be contiguous"
        bounds = np.array([r0 for r0, _ in self.ranges] + [self.ranges[-1][1]], dtype=np.int32)
        self.n_loc = int(bounds[rank + 1] - bounds[rank])
        lib = _abi.lib()
        handle = C.c_void_p()
        _abi.check(lib.spuq_sg_precond_mean_fd_rows(C.byref(handle), _abi.device_index(), self.nx, float(d), float(ox),
                                                    float(oy), self.world, self.rank, _abi.ptr(bounds)), "precond_mean_fd_rows")
        self.handle = handle
        vals = [C.c_int32() for _ in range(5)]
        _abi.check(lib.spuq_sg_precond_strip_info(handle, *[C.byref(v) for v in vals]), "precond_strip_info")
        self.P, self.q_loc, self.q_glob, self.first_sep, self.gf_ld = [int(v.value) for v in vals]
        if self.world > 1:
            info = [None] * self.world
            dist.all_gather_object(info, (self.first_sep, self.q_loc), group=group)
            self.sep_blocks = info                                  # (first separator, count) of every rank
        else:
            self.sep_blocks = [(self.first_sep, self.q_loc)]
        self.q_max = max(c for _, c in self.sep_blocks)
        self.schur = schur if self.world > 1 else "gather"
        # separator rows a rank needs for its final pass: the one above its first strip and its own
        self.need = [(max(f0 - 1, 0), f0 + cnt) for f0, cnt in self.sep_blocks]
        self._bufs = None

    def _buffers(self, L, like):
        if self._bufs is None or self._bufs[0] != L:
            mk = lambda *shape: torch.empty(shape, dtype=like.dtype, device=like.device)      # noqa: E731
            self._bufs = (L, mk(L, self.gf_ld, self.nx), mk(L, self.gf_ld, self.nx),
                          mk(max(self.q_glob, 1), L, self.nx), mk(max(self.q_glob, 1), L, self.nx),
                          mk(self.world, max(self.q_max, 1), L, self.nx), mk(L, self.nx), mk(L, self.nx))
        return self._bufs[1:]

    def solve(self, Rc, Sc, dot_out=None):
        """Sc = (A_0^{-1} R) restricted to the owned rows.  Rc, Sc: L x (n_loc * nx) slabs of the owned rows, or
        L x n_loc x nx VIEWS of them inside halo-padded slabs (rows and grid columns contiguous, any column
        stride): the kernels take the stride, nothing is copied.  ``dot_out`` (device scalar, one element):
        receives <R, S> over the band, computed by the final strip pass itself."""
        lib = _abi.lib()
        L = Rc.shape[0]
        ld = 0
        for T in (Rc, Sc):
            if T.dim() == 3:
                assert T.shape == (L, self.n_loc, self.nx) and T.stride(2) == 1 and T.stride(1) == self.nx
                ld = ld or int(T.stride(0))
                assert ld == int(T.stride(0)), "R and S must share their column stride"
            else:
                assert T.is_contiguous() and T.shape == (L, self.n_loc * self.nx)
        if ld:
            assert Rc.dim() == 3 and Sc.dim() == 3, "R and S must both be views when one is"
        Gf, Gl, Z, Z2, gath, snd, rcv = self._buffers(L, Rc)
        st = _abi.stream_ptr()
        _abi.check(lib.spuq_sg_precond_strip_forward(self.handle, L, _abi.ptr(Rc), ld, _abi.ptr(Gf), _abi.ptr(Gl), st), "strip_forward")
        if self.world > 1:
            ops = []
            if self.rank > 0:
                snd.copy_(Gf[:, 0, :])                              # first row of my first strip -> previous rank
                ops.append(dist.P2POp(dist.isend, snd, self.rank - 1, self.group))
            if self.rank < self.world - 1:
                ops.append(dist.P2POp(dist.irecv, rcv, self.rank + 1, self.group))
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if self.rank < self.world - 1:
                Gf[:, self.gf_ld - 1, :].copy_(rcv)
        U_ptr = _abi.ptr(Z)
        if self.q_glob > 0:
            _abi.check(lib.spuq_sg_precond_strip_assemble(self.handle, L, _abi.ptr(Rc), ld, _abi.ptr(Gf), _abi.ptr(Gl), _abi.ptr(Z), st),
                       "strip_assemble")
        if self.q_glob > 0 and self.schur == "sharded":
            U_loc, a_me = self._schur_sharded(Z, L)
            # strip_final indexes separators globally: shift the base so that row a_me of U is U_loc[0]
            U_ptr = C.c_void_p(U_loc.data_ptr() - a_me * L * self.nx * 8)
            self._u_keep = U_loc
        elif self.q_glob > 0:
            if self.world > 1:
                mine = gath[self.rank]
                if self.q_loc:
                    mine[:self.q_loc].copy_(Z[self.first_sep:self.first_sep + self.q_loc])
                if dist.get_backend(self.group) == "nccl":
                    dist.all_gather_into_tensor(gath, mine.clone(), group=self.group)
                else:
                    parts = [torch.empty_like(mine) for _ in range(self.world)]
                    dist.all_gather(parts, mine.clone(), group=self.group)
                    for r_, part in enumerate(parts):
                        gath[r_].copy_(part)
                for r_, (f0, cnt) in enumerate(self.sep_blocks):
                    if cnt and r_ != self.rank:
                        Z[f0:f0 + cnt].copy_(gath[r_, :cnt])
            _abi.check(lib.spuq_sg_precond_strip_schur(self.handle, L, _abi.ptr(Z), _abi.ptr(Z2), st), "strip_schur")
        _abi.check(lib.spuq_sg_precond_strip_final(self.handle, L, _abi.ptr(Rc), U_ptr, _abi.ptr(Sc), ld, _abi.ptr(dot_out), st),
                   "strip_final")
        return Sc

    def _schur_sharded(self, Z, L):
        """Separator system with the columns sharded over the ranks.  Z [q_glob][L][nx] holds this rank's rows;
        returns (U_loc [rows][L][nx], global index of its first row) with the rows this rank's strips touch."""
        lib = _abi.lib()
        world, nx, rank = self.world, self.nx, self.rank
        Lc = -(-L // world)
        Lp = Lc * world
        f0, cnt = self.sep_blocks[rank]
        mk = lambda *shape: torch.empty(shape, dtype=Z.dtype, device=Z.device)      # noqa: E731
        mine = Z[f0:f0 + cnt]
        if Lp != L:
            pad = torch.zeros((cnt, Lp, nx), dtype=Z.dtype, device=Z.device)
            pad[:, :L, :] = mine
            mine = pad
        send = list(mine.view(cnt, world, Lc, nx).permute(1, 0, 2, 3).contiguous().unbind(0))      # [world] x [cnt][Lc][nx]
        Zc = mk(max(self.q_glob, 1), Lc, nx)
        recv = [Zc[g0:g0 + c] for g0, c in self.sep_blocks]                       # contiguous row blocks, rank order
        _all_to_all(send, recv, rank, self.group)
        Z2 = mk(max(self.q_glob, 1), Lc, nx)
        _abi.check(lib.spuq_sg_precond_strip_schur(self.handle, Lc, _abi.ptr(Zc), _abi.ptr(Z2), _abi.stream_ptr()), "strip_schur")
        a_me, b_me = self.need[rank]
        back = [Zc[a:b].contiguous() for a, b in self.need]
        got = mk(world, max(b_me - a_me, 1), Lc, nx)
        _all_to_all(back, [got[q][:b_me - a_me] for q in range(world)], rank, self.group)
        U_loc = got[:, :b_me - a_me].permute(1, 0, 2, 3).reshape(b_me - a_me, Lp, nx)
        U_loc = U_loc[:, :L, :].contiguous() if Lp != L else U_loc.contiguous()
        return U_loc, a_me

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _abi.lib().spuq_sg_precond_destroy(self.handle)
        except Exception:
            pass


class RowShardedPCG:
    """``pcg`` (application/egsz/pcg.py:15-53) on row-sharded slabs.

    Every vector lives as the local slab of ``RowShardedApply`` (L x (y1-y0+2)*nx, halo rows zero
    except in the search direction right after its halo exchange).  Per pass:
      * ``z = A v``: halo exchange + the single-GPU kernel, then the output halo rows are zeroed, so
        that every inner product can run over the whole local slab;
      * the two inner products are local sums + one all-reduce of a scalar each;
      * the mean-based preconditioner solves with A_0 column by column and needs whole columns: the
        residual is transposed to a COLUMN sharding (rank q gets columns Lambda_q of all rows: every
        rank sends contiguous blocks, all-to-all over NVLink), solved with the single-GPU exact solver
        and transposed back.  With ``mean_stencil=None`` the identity is used.
    The reference's four failure conditions are checked on the all-reduced scalars, so all ranks
    take the same branch."""

    def __init__(self, sharded_apply, nx, ny, ranges, mean_stencil=None, mean_solver="strip"):
        self.A = sharded_apply
        self.group = sharded_apply.group
        self.world, self.rank = sharded_apply.world, sharded_apply.rank
        self.nx, self.ny = int(nx), int(ny)
        self.ranges = list(ranges)                       # (y0, y1) of every rank
        self.y0, self.y1 = self.ranges[self.rank]
        self.nyl = self.y1 - self.y0 + 2
        self.mean_stencil = mean_stencil                 # (d, ox, oy) of the constant-coefficient mean block
        # "partition": tridiagonal partition method, the slab stays row-sharded (default);
        # "transpose": all-to-all to column shards and back around the single-GPU solver
        # "strip" (default): the strip solver on every band + a gathered separator system (StripRowsMeanSolver)
        if mean_solver not in ("strip", "partition", "transpose"):
            raise ValueError("mean_solver must be 'strip', 'partition' or 'transpose'")
        self.mean_solver = mean_solver
        if min(r1 - r0 for r0, r1 in self.ranges) < 2:
            self.mean_solver = "transpose"
        self._solver = None
        self._partition = None
        self.last_history = []

    # ---- helpers -------------------------------------------------------------------------------
    def _zero_halos(self, X):
        X3 = X.view(X.shape[0], self.nyl, self.nx)
        X3[:, 0, :].zero_()
        X3[:, self.nyl - 1, :].zero_()

    def _allreduce(self, value, dev):
        if self.world == 1:
            return float(value)
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t[0].item())

    def _col_range(self, q, L):
        per = padded_columns(L, self.world)
        return min(q * per, L), min((q + 1) * per, L)

    def _exchange(self, send, recv):
        """send[q] -> rank q, recv[q] <- rank q (lists of contiguous tensors, possibly empty)."""
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all(recv, send, group=self.group)
            return
        ops = []
        for q in range(self.world):
            if q == self.rank:
                recv[q].copy_(send[q])
                continue
            if send[q].numel():
                ops.append(dist.P2POp(dist.isend, send[q], q, self.group))
            if recv[q].numel():
                ops.append(dist.P2POp(dist.irecv, recv[q], q, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def precondition(self, R, S, dot_out=None):
        """S = A_0^{-1} R on the owned rows (halo rows of S zero).  ``dot_out`` (device scalar): receives the
        local <R, S> when the solver can produce it on the way (returns True then, else False / S)."""
        if self.mean_stencil is None:
            S.copy_(R)
            self._zero_halos(S)
            return False if dot_out is not None else S
        L = R.shape[0]
        dev = R.device
        if self.mean_solver in ("partition", "strip"):
            if self._partition is None:
                if self.mean_solver == "strip":
                    try:
                        self._partition = StripRowsMeanSolver(self.nx, self.ranges, self.rank, self.mean_stencil, self.group)
                    except _abi.SpuqError:
                        self.mean_solver = "partition"         # grid the strip solver does not take
                if self._partition is None:
                    self._partition = PartitionedMeanSolver(self.nx, self.ny, self.ranges, self.rank, self.mean_stencil, self.group)
            R3 = R.view(L, self.nyl, self.nx)
            S3 = S.view(L, self.nyl, self.nx)
            S3[:, 0, :].zero_()
            S3[:, self.nyl - 1, :].zero_()
            if isinstance(self._partition, StripRowsMeanSolver):
                # the strip kernels work on the band where it lies inside the halo-padded slabs
                self._partition.solve(R3[:, 1:-1, :], S3[:, 1:-1, :], dot_out=dot_out)
                return True if dot_out is not None else S
            Rc = R3[:, 1:-1, :].contiguous().view(L, -1)
            Sc = torch.empty_like(Rc)
            self._partition.solve(Rc, Sc)
            S3[:, 1:-1, :].copy_(Sc.view(L, self.nyl - 2, self.nx))
            return False if dot_out is not None else S
        c0, c1 = self._col_range(self.rank, L)
        if self._solver is None:
            d, ox, oy = self.mean_stencil
            self._solver = MeanSolveSlabOperator(None, max(c1 - c0, 1), stencil=(self.nx, self.ny, d, ox, oy))
        R3 = R.view(L, self.nyl, self.nx)
        S3 = S.view(L, self.nyl, self.nx)
        if self.world == 1:
            Rc = R3[:, 1:-1, :].contiguous().view(L, self.ny * self.nx)     # (reshape alone keeps the column stride)
            Sc = torch.empty_like(Rc)
            self._solver.apply_slab(Rc, Sc)
            S3.zero_()
            S3[:, 1:-1, :].copy_(Sc.view(L, self.ny, self.nx))
            return False if dot_out is not None else S
        own = self.y1 - self.y0
        send = [R3[q0:q1, 1:-1, :].contiguous() for q0, q1 in (self._col_range(q, L) for q in range(self.world))]
        recv = [torch.empty((c1 - c0, r1 - r0, self.nx), dtype=R.dtype, device=dev) for r0, r1 in self.ranges]
        self._exchange(send, recv)
        Rc = torch.empty((c1 - c0, self.ny, self.nx), dtype=R.dtype, device=dev)
        for (r0, r1), blk in zip(self.ranges, recv):
            Rc[:, r0:r1, :].copy_(blk)
        Sc = torch.empty_like(Rc)
        if c1 > c0:
            self._solver.apply_slab(Rc.view(c1 - c0, -1), Sc.view(c1 - c0, -1))
        back = [Sc[:, r0:r1, :].contiguous() for r0, r1 in self.ranges]
        got = [torch.empty((q1 - q0, own, self.nx), dtype=R.dtype, device=dev)
               for q0, q1 in (self._col_range(q, L) for q in range(self.world))]
        self._exchange(back, got)
        S3.zero_()
        for q, blk in enumerate(got):
            q0, q1 = self._col_range(q, L)
            S3[q0:q1, 1:-1, :].copy_(blk)
        return False if dot_out is not None else S

    # ---- the solver ----------------------------------------------------------------------------
    def solve(self, plan, F, W0, eps=1e-4, maxiter=100, poll_every=8):
        """F, W0: local slabs (halo rows ignored).  Returns (W, zeta, numit) like ``pcg``; raises the
        reference's exceptions.

        The loop is enqueued by the host but decided on the device: zeta, alpha and the status word live in
        a device state buffer, the inner products are all-reduced in place on it (NCCL on the stream), the
        reference's branch conditions (pcg.py:33-45) are evaluated by the library's one-thread kernels, and
        every kernel of a pass is a no-op once the status has left RUNNING (spuq_sg_set_gate).  The host
        reads the status word once per ``poll_every`` passes."""
        lib = _abi.lib()
        dev = F.device
        n = F.numel()
        st = torch.zeros(_abi.PCG_STATE_DOUBLES, dtype=torch.float64, device=dev)
        st[4] = float(_abi.PCG_RUNNING)
        st[6] = float(eps) ** 2
        hist = torch.zeros(max(int(maxiter), 1), dtype=torch.float64, device=dev)
        scratch = torch.zeros(int(lib.spuq_sg_reduce_scratch_bytes()) // 8 + 1, dtype=torch.float64, device=dev)
        sp = lambda k: C.c_void_p(st.data_ptr() + 8 * k)          # noqa: E731  address of state word k
        stream = _abi.stream_ptr

        def dot_to(x, y, slot, stage):
            """state[slot] = <x, y> summed over the ranks (staged in state[stage])."""
            _abi.check(lib.spuq_sg_dot(n, _abi.ptr(x), _abi.ptr(y), sp(stage), _abi.ptr(scratch), stream()), "dot")
            if self.world > 1:
                dist.all_reduce(st[stage:stage + 1], op=dist.ReduceOp.SUM, group=self.group)
            _abi.check(lib.spuq_sg_pcg_accept(_abi.ptr(st), stage, slot, stream()), "pcg_accept")

        W = W0.clone()
        self._zero_halos(W)
        rho = F.clone()
        self._zero_halos(rho)
        z = torch.empty_like(W)
        self.A.apply(plan, W, z)                          # rho = f - A w0      (pcg.py:26)
        self._zero_halos(z)
        self._zero_halos(W)
        _axpy(-1.0, z, rho)
        s = torch.empty_like(W)
        self.precondition(rho, s)                         # s = P rho           (:27)
        v = s.clone()                                     # v = s               (:28)
        dot_to(rho, s, 0, 8)                              # zeta = <rho, s>     (:30)
        lib.spuq_sg_set_gate(sp(4))
        status, done = _abi.PCG_RUNNING, 1
        try:
            i = 1
            while i < maxiter and status == _abi.PCG_RUNNING:
                stop = min(maxiter, i + max(int(poll_every), 1))
                for i in range(i, stop):                  # (:31)
                    _abi.check(lib.spuq_sg_pcg_check_zeta_hist(_abi.ptr(st), i, _abi.ptr(hist), stream()), "check_zeta")   # :33-38
                    self.A.apply(plan, v, z)              # z = A v             (:40)
                    self._zero_halos(z)
                    dot_to(z, v, 1, 9)                    # alpha = <z, v>      (:41)
                    _abi.check(lib.spuq_sg_pcg_check_alpha(_abi.ptr(st), stream()), "check_alpha")                      # :42-45
                    _abi.check(lib.spuq_sg_axpy2(n, sp(0), sp(1), _abi.ptr(v), _abi.ptr(W), _abi.ptr(z), _abi.ptr(rho),
                                                 stream()), "axpy2")                                                   # :47-48
                    if self.precondition(rho, s, dot_out=st[10:11]):      # (:49) with the local <rho, s> (:50)
                        if self.world > 1:
                            dist.all_reduce(st[10:11], op=dist.ReduceOp.SUM, group=self.group)
                        _abi.check(lib.spuq_sg_pcg_accept(_abi.ptr(st), 10, 2, stream()), "pcg_accept")
                    else:
                        dot_to(rho, s, 2, 10)             # zeta_new = <rho, s> (:50)
                    _abi.check(lib.spuq_sg_xpay(n, sp(2), sp(0), _abi.ptr(s), _abi.ptr(v), stream()), "xpay")            # :51
                    _abi.check(lib.spuq_sg_pcg_rotate(_abi.ptr(st), stream()), "rotate")
                i = stop
                h = st.cpu()                              # the one host read-back per poll_every passes
                status, done = int(h[4]), int(h[5])
        finally:
            lib.spuq_sg_set_gate(None)
        h = st.cpu()
        zeta, alpha = float(h[0]), float(h[1])
        self.last_history = hist[:done].cpu().tolist() if status != _abi.PCG_RUNNING else hist[:max(maxiter - 1, 0)].cpu().tolist()
        if status == _abi.PCG_CONVERGED:
            self._zero_halos(W)
            return W, zeta, done
        if status == _abi.PCG_PRECOND_NOT_PD:
            raise Exception(PCG_MESSAGES[status] % zeta)
        if status in (_abi.PCG_SINGULAR, _abi.PCG_NOT_PD):
            raise Exception(PCG_MESSAGES[status] % alpha)
        raise Exception(PCG_MESSAGES[_abi.PCG_NOT_CONVERGED])
