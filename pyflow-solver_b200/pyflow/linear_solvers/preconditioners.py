"""Jacobi preconditioner (drop-in for the reference's
``src/pyflow/linear_solvers/preconditioners.py:jacobi_preconditioner`` :8-25).

On the device the only preconditioner is the one the reference actually uses
on the hot path: the inverse diagonal of ``assemble_negative_laplacian`` --
1.0 on the boundary ring, ``1/(2/dx^2 + 2/dy^2)`` inside -- applied inside the
CG kernels.  ``RingJacobi`` is its descriptor.
"""
from __future__ import annotations

from dataclasses import dataclass


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


def jacobi_preconditioner(A) -> RingJacobi:
    """``A`` is the device operator descriptor (``NegLaplacianPinned``)."""
    from .interface import NegLaplacianPinned
    if isinstance(A, NegLaplacianPinned):
        return RingJacobi(A.nx, A.ny, A.dx, A.dy)
    raise NotImplementedError(
        "jacobi_preconditioner on the device needs the matrix-free NegLaplacianPinned operator; assembled sparse "
        "matrices are not part of the B200 hot path (no CPU fallback)")
