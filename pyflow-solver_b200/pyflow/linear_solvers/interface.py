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
``k_bicgstab`` from an array of inverse entries.

``A`` may also be the ASSEMBLED matrix of the reference --
``assemble_negative_laplacian(nx, ny, dx, dy)`` (pressure_solver.py:15-40), the one
matrix the reference itself builds -- as a scipy sparse matrix: ``recognise_ring5``
reads its CSR arrays, accepts it when every boundary-ring row is an identity row
and every interior row is the same 5-point stencil, and the solve runs on the
device with that stencil evaluated in CSR column order (``pf_op_krylov_stencil5``),
i.e. with the arithmetic of scipy's ``csr_matvec``.  Anything else raises: there
is no CPU fallback.
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


@dataclass(frozen=True)
class Ring5Matrix:
    """What ``recognise_ring5`` found: identity rows on the boundary ring of an ``ny x nx`` grid, the 5-point row
    ``coef = (south, west, centre, east, north)`` everywhere inside."""
    nx: int
    ny: int
    coef: tuple

    @property
    def shape(self):
        n = self.nx * self.ny
        return (n, n)


def recognise_ring5(A) -> Ring5Matrix | None:
    """``Ring5Matrix`` if the sparse matrix ``A`` has the structure of the reference's
    ``assemble_negative_laplacian`` (pressure_solver.py:15-40), else ``None``.  Exact checks on the CSR arrays:
    row k of a boundary cell holds the single entry A[k, k] = 1; every interior row holds five entries at columns
    k-nx, k-1, k, k+1, k+nx with the same five values in every row."""
    tocsr = getattr(A, "tocsr", None)
    shape = getattr(A, "shape", None)
    if tocsr is None or shape is None or len(shape) != 2 or shape[0] != shape[1] or shape[0] < 9:
        return None
    M = tocsr()
    M.sort_indices()
    n = int(shape[0])
    indptr, indices, data = np.asarray(M.indptr), np.asarray(M.indices), np.asarray(M.data, dtype=np.float64)
    counts = np.diff(indptr)
    five = np.flatnonzero(counts == 5)
    if five.size == 0 or np.any((counts != 1) & (counts != 5)):
        return None
    k0 = int(five[0])
    cols0 = indices[indptr[k0]:indptr[k0] + 5]
    nx = int(cols0[4] - cols0[2])
    if nx < 3 or n % nx or k0 != nx + 1:
        return None
    ny = n // nx
    if ny < 3:
        return None
    jj, ii = np.divmod(np.arange(n), nx)
    ring = (ii == 0) | (ii == nx - 1) | (jj == 0) | (jj == ny - 1)
    if not np.array_equal(counts == 1, ring):
        return None
    rk = np.flatnonzero(ring)
    if not (np.array_equal(indices[indptr[rk]], rk) and np.all(data[indptr[rk]] == 1.0)):
        return None
    ik = np.flatnonzero(~ring)
    base = indptr[ik][:, None] + np.arange(5)[None, :]
    want_cols = ik[:, None] + np.array([-nx, -1, 0, 1, nx])[None, :]
    if not np.array_equal(indices[base], want_cols):
        return None
    vals = data[base]
    if not np.all(vals == vals[0][None, :]) or not np.all(np.isfinite(vals[0])):
        return None
    return Ring5Matrix(nx, ny, tuple(float(v) for v in vals[0]))


def solve(A, b, method: Literal["cg", "bicgstab"] = "bicgstab", tol: float = 1e-8, maxiter: int = 500,
          **kwargs) -> LinearSolveResult:
    if not isinstance(A, (NegLaplacianPinned, Ring5Matrix)):
        found = recognise_ring5(A)
        if found is None:
            raise NotImplementedError(
                "the device solve handles the matrix-free NegLaplacianPinned operator and the assembled matrix of "
                "assemble_negative_laplacian (identity ring rows + one 5-point interior row); this operator is "
                "neither, and there is no CPU fallback")
        A = found
    M = kwargs.pop("M", None)
    if kwargs:
        raise TypeError(f"unsupported solver arguments: {sorted(kwargs)}")
    if isinstance(A, Ring5Matrix):
        return _solve_ring5(A, b, method, tol, maxiter, M)
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
    return _result(x, b, info, rnorm, tol, method)


def _result(x, b, info, rnorm, tol, method):
    bnorm = float(np.linalg.norm(b)) or 1.0
    converged = (info == 0) and (rnorm <= tol * bnorm)
    iterations = info if info > 0 else 0
    return LinearSolveResult(x=x.reshape(b.shape), converged=converged, iterations=iterations,
                             residual_norm=float(rnorm), method=method)


def _solve_ring5(A: Ring5Matrix, b, method, tol, maxiter, M):
    """The assembled matrix of the reference on the device.  Preconditioner: none, or the ring-identity diagonal
    (what ``jacobi_preconditioner(A)`` of that matrix is)."""
    jacobi, inv_inner = False, 1.0
    if M is not None:
        M = classify(M, A.nx, A.ny)
        if not isinstance(M, RingJacobi):
            raise NotImplementedError("on an assembled matrix only the ring-identity (Jacobi) diagonal has a device form")
        jacobi, inv_inner = True, M.inv_diag_interior
    b = np.asarray(b, dtype=np.float64)
    slab = slab_for_shape(A.ny, A.nx)
    x, info, _iters, rnorm = slab.krylov_stencil5(method, b.reshape(A.ny, A.nx), A.coef, tol, maxiter, jacobi=jacobi,
                                                  inv_diag_interior=inv_inner)
    return _result(x, b, info, rnorm, tol, method)
