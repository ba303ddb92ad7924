"""Preconditioned conjugate gradients with all solver state on the device.

Drop-in for ``pcg(A, f, P, w0, eps=1e-4, maxiter=100) -> (w, zeta, numit)`` of
spuq/application/egsz/pcg.py:15-53: same arguments, same return triple (``numit`` is the 1-based
index of the loop pass in which ``zeta <= eps**2`` was detected), same four failure messages raised
as plain ``Exception`` (:33-36, :42-45, :53), same type check (Operator, Vector, Operator, Vector, :14).

The loop itself runs in libspuq_sg (``spuq_sg_pcg``): operator application, fused dot products and
vector updates read and write zeta / alpha in device memory, the reference's branch conditions are
evaluated by one-thread kernels, and the host only polls a status word every ``poll_every`` passes.
The extra ``<rho, rho>`` the reference computes every pass for its log line (:32) is not computed; the
line itself ("pcg iter: i -> zeta=...") is written after the solve from the device-side zeta history when the
module's logger is enabled for INFO.
"""
import ctypes as C
import logging

import numpy as np
import torch

from . import _abi
from .linalg import (DiagonalMatrixOperator, FlatVector, MatrixOperator, MatrixSolveOperator,
                     MultiplicationOperator, Operator, ScipyOperator, Vector, _CsrLeaf)
from .multi_operator import (CsrSlabOperator, MeanSolveSlabOperator, MultiOperator,
                             PreconditioningOperator, SGPlan, SparseLUSlabOperator)
from .multi_vector import MultiVector, SlabMultiVector
from .multiindex import Multiindex

__all__ = ["pcg", "PCG_MESSAGES"]

logger = logging.getLogger(__name__)

PCG_MESSAGES = {
    _abi.PCG_PRECOND_NOT_PD: "Preconditioner for PCG is not positive definite (%s)",
    _abi.PCG_SINGULAR: "Matrix for PCG is singular (%s)",
    _abi.PCG_NOT_PD: "Matrix for PCG is not positive definite (%s)",
    _abi.PCG_NOT_CONVERGED: "PCG did not converge",
}


class _Precond:
    """RAII wrapper of a ``spuq_sg_precond`` handle."""

    def __init__(self, handle, keep=None):
        self.handle, self.keep = handle, keep

    @classmethod
    def identity(cls, scale):
        h = C.c_void_p()
        _abi.check(_abi.lib().spuq_sg_precond_identity(C.byref(h), float(scale)), "precond_identity")
        return cls(h)

    @classmethod
    def from_plan(cls, plan, keep=None):
        h = C.c_void_p()
        _abi.check(_abi.lib().spuq_sg_precond_from_plan(C.byref(h), plan.handle), "precond_from_plan")
        return cls(h, keep=(plan, keep))

    def close(self):
        if self.handle is not None and not isinstance(self.keep, (MeanSolveSlabOperator, SparseLUSlabOperator)):
            _abi.lib().spuq_sg_precond_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _precond_for(P, slab):
    """Translate a preconditioner Operator into a device handle for slab vectors like ``slab``."""
    L = len(slab)
    if isinstance(P, MultiplicationOperator):
        return _Precond.identity(P._a)
    if isinstance(P, PreconditioningOperator):
        solver = P.slab_operator(slab)
        if isinstance(solver, (MeanSolveSlabOperator, SparseLUSlabOperator)):
            solver.set_columns(L)
            return _Precond(solver.handle, keep=solver)
        if solver.plan.L != L:
            solver.plan.close()
            solver.plan = SGPlan(solver.blocks, L)
        return _Precond.from_plan(solver.plan, keep=solver)
    if isinstance(P, (_CsrLeaf,)):
        op = CsrSlabOperator(P.as_csr(), L)
        return _Precond.from_plan(op.plan, keep=op)
    raise TypeError("pcg: preconditioner of type %s has no device form" % type(P).__name__)


def _solve_on_device(plan, precond, F, W0, eps, maxiter, poll_every, history):
    lib = _abi.lib()
    W = W0.clone()
    work_bytes = int(lib.spuq_sg_pcg_work_bytes(plan.handle, precond.handle))
    work = torch.empty(work_bytes // 8 + 1, dtype=torch.float64, device=W.device)
    zeta, numit, status = C.c_double(), C.c_int32(), C.c_int32()
    hist = np.zeros(max(int(maxiter), 1), dtype=np.float64)
    _abi.check(lib.spuq_sg_pcg(plan.handle, precond.handle, _abi.ptr(F), _abi.ptr(W), float(eps), int(maxiter),
                               int(poll_every), _abi.ptr(work), C.byref(zeta), C.byref(numit), C.byref(status),
                               _abi.ptr(hist), _abi.stream_ptr()), "pcg")
    if history is not None:
        history.extend(hist[:numit.value].tolist())
    if logger.isEnabledFor(logging.INFO):
        for i, z in enumerate(hist[:numit.value].tolist(), 1):          # pcg.py:32 (without the rho^2 it adds)
            logger.info("pcg iter: %s -> zeta=%s", i, z)
    return W, zeta.value, numit.value, status.value


def _raise_for(status, zeta):
    if status == _abi.PCG_CONVERGED:
        return
    msg = PCG_MESSAGES[status]
    raise Exception(msg % zeta if "%s" in msg else msg)


def pcg(A, f, P, w0, eps=1e-4, maxiter=100, poll_every=8, history=None):
    if not (isinstance(A, Operator) and isinstance(P, Operator)
            and isinstance(f, Vector) and isinstance(w0, Vector)):
        raise TypeError("pcg(A: Operator, f: Vector, P: Operator, w0: Vector, eps, maxiter)")
    if maxiter < 2:
        raise Exception("PCG did not converge")         # the reference loop body never runs

    if isinstance(A, MultiOperator):
        if not (isinstance(f, MultiVector) and isinstance(w0, MultiVector)):
            raise TypeError("pcg: a MultiOperator needs MultiVector arguments")
        fs = f if isinstance(f, SlabMultiVector) else SlabMultiVector.from_multivector(f)
        ws = w0 if isinstance(w0, SlabMultiVector) else SlabMultiVector.from_multivector(w0)
        assert fs.active_indices() == ws.active_indices()
        plan = A.plan_for(ws)
        precond = _precond_for(P, ws)
        W, zeta, numit, status = _solve_on_device(plan, precond, fs.slab, ws.slab, eps, maxiter,
                                                  poll_every, history)
        precond.close()
        _raise_for(status, zeta)
        out = ws.like(W)
        if not isinstance(w0, SlabMultiVector):
            out = out.to_multivector(type(w0))
        return (out, zeta, numit)

    if isinstance(A, _CsrLeaf) and isinstance(f, FlatVector) and isinstance(w0, FlatVector):
        # the reference's own test configuration (tests/test_pcg.py): one spatial block, flat
        # vectors -- run it as a one-column slab through the same device loop
        op = CsrSlabOperator(A.as_csr(), 1)
        n = op.plan.n_rows
        dev = _abi.device()
        one = SlabMultiVector([Multiindex()], n=n, basis=f.basis)
        up = lambda v: torch.from_numpy(np.ascontiguousarray(v.coeffs, dtype=np.float64).reshape(1, -1)).to(dev)  # noqa: E731
        precond = _precond_for(P, one)
        W, zeta, numit, status = _solve_on_device(op.plan, precond, up(f), up(w0), eps, maxiter,
                                                  poll_every, history)
        precond.close()
        _raise_for(status, zeta)
        return (type(w0)(W.cpu().numpy().reshape(-1), w0.basis), zeta, numit)

    raise TypeError("pcg: operator of type %s has no device form (supported: MultiOperator, "
                    "MatrixOperator, DiagonalMatrixOperator, ScipyOperator)" % type(A).__name__)
