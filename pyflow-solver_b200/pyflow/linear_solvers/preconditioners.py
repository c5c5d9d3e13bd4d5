"""Jacobi preconditioner (drop-in for the reference's
``src/pyflow/linear_solvers/preconditioners.py:jacobi_preconditioner`` :8-25).

The reference builds ``M`` from the diagonal of an ASSEMBLED matrix,
``invD = where(D != 0, 1/D, 0)``, and on the hot path that matrix is always
``assemble_negative_laplacian(nx, ny, dx, dy)`` (pressure_solver.py:139-143):
identity rows on the boundary ring, ``2/dx^2 + 2/dy^2`` inside.  The device
applies exactly that diagonal analytically inside the CG kernels, so here a
preconditioner is a DESCRIPTOR, not an operator:

* ``jacobi_preconditioner(NegLaplacianPinned(...))`` -> ``RingJacobi`` (the
  matrix-free route: nothing is assembled);
* ``jacobi_preconditioner(<scipy sparse matrix or 2-D array>)`` -- the
  reference's own recipe -- reads the diagonal and returns a
  ``DiagonalJacobi``; at solve time it is accepted if it IS the ring-identity
  diagonal of the grid being solved (checked value by value), i.e. for the
  matrix ``assemble_negative_laplacian`` returns;
* a foreign operator object (``scipy.sparse.linalg.LinearOperator``, e.g. what the
  reference's own ``jacobi_preconditioner`` returns, or anything with ``matvec``
  and ``shape``) is PROBED: ``d = M(1)``, then ``M(v) == d * v`` bit for bit on two
  fixed pseudo-random vectors.  A diagonal operator becomes a ``DiagonalJacobi``;
* the linear-solve seam (``interface.solve``) runs ANY diagonal on the device
  (``k_cg_diag`` / ``k_bicgstab`` read the inverse entries from an array);
  ``solve_pressure_poisson`` -- the fused time-step path -- applies the
  ring-identity diagonal only and refuses the others;
* operators that are not diagonal raise ``NotImplementedError``: they exist only
  as host callables and there is no host path to fall back to.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass(frozen=True)
class RingJacobi:
    nx: int
    ny: int
    dx: float
    dy: float

    @property
    def inv_diag_interior(self) -> float:
        return 1.0 / (2.0 / self.dx ** 2 + 2.0 / self.dy ** 2)

    @property
    def shape(self):
        n = self.nx * self.ny
        return (n, n)


@dataclass(frozen=True, eq=False)
class DiagonalJacobi:
    """Inverse diagonal read off an assembled matrix (reference preconditioners.py:17-19)."""
    inv_diag: np.ndarray = field(repr=False)

    @property
    def shape(self):
        n = int(self.inv_diag.shape[0])
        return (n, n)

    def as_ring(self, nx: int, ny: int) -> RingJacobi | None:
        """The equivalent ``RingJacobi`` if this is the ring-identity diagonal of an ``ny x nx`` grid:
        1.0 on the whole boundary ring and one common value inside.  ``None`` otherwise."""
        if self.inv_diag.shape != (nx * ny,):
            return None
        d = self.inv_diag.reshape(ny, nx)
        ring = np.ones((ny, nx), dtype=bool)
        ring[1:-1, 1:-1] = False
        if not np.all(d[ring] == 1.0):
            return None
        inner = d[~ring]
        if inner.size == 0:
            return RingJacobi(nx, ny, 1.0, 1.0)
        if not np.all(inner == inner[0]) or not inner[0] > 0.0:
            return None
        return _RingWithValue(nx, ny, 1.0, 1.0, float(inner[0]))


@dataclass(frozen=True)
class _RingWithValue(RingJacobi):
    """Ring-identity diagonal whose interior value was read from a matrix rather than derived from dx, dy."""
    value: float = 0.25

    @property
    def inv_diag_interior(self) -> float:
        return self.value


def jacobi_preconditioner(A):
    """Descriptor of the Jacobi preconditioner of ``A`` (see the module docstring)."""
    from .interface import NegLaplacianPinned
    if isinstance(A, NegLaplacianPinned):
        return RingJacobi(A.nx, A.ny, A.dx, A.dy)
    diagonal = getattr(A, "diagonal", None)
    if callable(diagonal) and len(getattr(A, "shape", ())) == 2 and A.shape[0] == A.shape[1]:
        D = np.asarray(diagonal(), dtype=np.float64).reshape(-1)
        with np.errstate(divide="ignore"):
            return DiagonalJacobi(np.where(D != 0, 1.0 / D, 0.0))
    raise NotImplementedError(
        "jacobi_preconditioner needs the matrix-free NegLaplacianPinned operator or an assembled (sparse) matrix; "
        "general LinearOperators have no diagonal to read and there is no CPU fallback")


def probe_diagonal(op, n: int) -> DiagonalJacobi | None:
    """``DiagonalJacobi`` of a foreign operator object if it acts as a diagonal matrix on R^n, else ``None``.
    The entries are ``op(1)``; diagonality is accepted when ``op(v) == op(1) * v`` BIT FOR BIT on two fixed
    pseudo-random vectors (what ``lambda x: invD * x`` -- the reference's preconditioners.py:20-21 -- satisfies;
    an operator with any off-diagonal coupling does not)."""
    apply = getattr(op, "matvec", None) or (op if callable(op) else None)
    shape = getattr(op, "shape", None)
    if apply is None or shape is None or tuple(shape) != (n, n):
        return None
    try:
        d = np.asarray(apply(np.ones(n)), dtype=np.float64).reshape(-1)
        if d.shape != (n,):
            return None
        rng = np.random.default_rng(20250117)
        for _ in range(2):
            v = rng.standard_normal(n) * np.exp(rng.uniform(-3.0, 3.0, n))
            got = np.asarray(apply(v.copy()), dtype=np.float64).reshape(-1)
            if got.shape != (n,) or not np.array_equal(got, d * v, equal_nan=True):
                return None
    except Exception:
        return None
    return DiagonalJacobi(d)


def classify(preconditioner, nx: int, ny: int) -> RingJacobi | DiagonalJacobi:
    """The device form of ``preconditioner`` on an ``ny x nx`` grid: ``RingJacobi`` (analytic, every CG kernel),
    or a general ``DiagonalJacobi`` (array form, the generic Krylov kernels of the linear-solve seam).
    NotImplementedError for operators that are not diagonal."""
    if isinstance(preconditioner, RingJacobi):
        return preconditioner
    if not isinstance(preconditioner, DiagonalJacobi):
        found = probe_diagonal(preconditioner, nx * ny)
        if found is None:
            raise NotImplementedError(
                "this preconditioner is not a diagonal operator on the grid being solved; only diagonal "
                "preconditioners have a device form (there is no CPU fallback)")
        preconditioner = found
    if preconditioner.inv_diag.shape != (nx * ny,):
        raise NotImplementedError(f"diagonal preconditioner of size {preconditioner.inv_diag.shape[0]} on a "
                                  f"{ny} x {nx} grid")
    return preconditioner.as_ring(nx, ny) or preconditioner


def resolve(preconditioner, nx: int, ny: int) -> RingJacobi:
    """The ring-identity Jacobi descriptor of ``preconditioner`` -- what the fused time-step path applies -- or
    NotImplementedError."""
    found = classify(preconditioner, nx, ny)
    if isinstance(found, RingJacobi):
        return found
    raise NotImplementedError(
        "this diagonal preconditioner is not the ring-identity Jacobi diagonal of the grid being solved "
        "(1 on the boundary ring, one value inside); the projection step applies only that one "
        "(linear_solvers.solve runs general diagonals)")
