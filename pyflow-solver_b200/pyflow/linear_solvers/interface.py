"""Linear-solve seam (drop-in for the reference's
``src/pyflow/linear_solvers/interface.py``: ``LinearSolveResult`` :9-15,
``solve`` :23-53).

The reference hands any matrix / LinearOperator to ``scipy.sparse.linalg.cg``.
The B200 path accelerates the one operator the hot path uses -- the matrix-free
pinned Neumann ``-laplacian_coherent`` of pressure_solver.py:97-105 -- so
``solve`` accepts its descriptor ``NegLaplacianPinned`` and runs scipy's CG
recurrence (x0 = 0, strict ``||r|| < tol*||b||`` test at the top of each
iteration, info = maxiter when exhausted) on the device; ``method='bicgstab'``
(the signature's default, interface.py:23,44-45; no call site in the reference
asks for it) runs scipy's BiCGSTAB recurrence, exit codes included, in
``k_bicgstab``.  ``M`` may be the ring-identity Jacobi descriptor, any diagonal
(``DiagonalJacobi``), or a foreign operator object that proves to be diagonal
(``preconditioners.probe_diagonal``); general diagonals run in ``k_cg_diag`` /
``k_bicgstab`` from an array of inverse entries.  Anything else raises: there is
no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal

import numpy as np

from .._b200.device import slab_for_shape
from .preconditioners import RingJacobi, classify


@dataclass(slots=True)
class LinearSolveResult:
    x: np.ndarray
    converged: bool
    iterations: int
    residual_norm: float
    method: str


@dataclass(frozen=True)
class NegLaplacianPinned:
    """x -> -(laplacian_coherent(x)) with row 0 replaced by the identity."""
    nx: int
    ny: int
    dx: float
    dy: float

    @property
    def shape(self):
        n = self.nx * self.ny
        return (n, n)

    def matvec(self, x):
        from ..numerics.core_ops_coherent import laplacian_coherent
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        out = (-laplacian_coherent(x.reshape(self.ny, self.nx), self.dx, self.dy)).reshape(-1)
        out[0] = x[0]
        return out

    __matmul__ = matvec


def solve(A, b, method: Literal["cg", "bicgstab"] = "bicgstab", tol: float = 1e-8, maxiter: int = 500,
          **kwargs) -> LinearSolveResult:
    if not isinstance(A, NegLaplacianPinned):
        raise NotImplementedError("the device solve handles the matrix-free NegLaplacianPinned operator only")
    M = kwargs.pop("M", None)
    if kwargs:
        raise TypeError(f"unsupported solver arguments: {sorted(kwargs)}")
    if M is not None:
        M = classify(M, A.nx, A.ny)         # RingJacobi (analytic) or a general diagonal; non-diagonal operators raise
    b = np.asarray(b, dtype=np.float64)
    slab = slab_for_shape(A.ny, A.nx)
    # like the reference, anything that is not 'cg' takes the bicgstab branch (interface.py:42-45)
    if M is None or isinstance(M, RingJacobi):
        run = slab.pcg if method == "cg" else slab.bicgstab
        x, info, _iters, rnorm = run(b.reshape(A.ny, A.nx), A.dx, A.dy, tol, maxiter, jacobi=M is not None,
                                     inv_diag_interior=(M.inv_diag_interior if M is not None else 1.0))
    else:
        x, info, _iters, rnorm = slab.krylov_diag(method, b.reshape(A.ny, A.nx), M.inv_diag.reshape(A.ny, A.nx),
                                                  A.dx, A.dy, tol, maxiter)
    bnorm = float(np.linalg.norm(b)) or 1.0
    converged = (info == 0) and (rnorm <= tol * bnorm)
    iterations = info if info > 0 else 0
    return LinearSolveResult(x=x.reshape(b.shape), converged=converged, iterations=iterations,
                             residual_norm=float(rnorm), method=method)
