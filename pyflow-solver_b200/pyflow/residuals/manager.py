"""Residual tracker (reference ``src/pyflow/residuals/manager.py`` :2-16): keeps
the ``Ru`` series (seeded with 1.0) for ``residual_drop_Ru_orders``; plateau
detection is a stub that never fires, as in the reference."""
from __future__ import annotations


class _Series:
    def __init__(self):
        self.values = [1.0]


class ResidualManager:
    def __init__(self):
        self.series = {"Ru": _Series()}

    def __call__(self, *args, **kwargs):
        return None

    def add(self, name, value):
        if name == "Ru":
            self.series["Ru"].values.append(value)

    def plateau_detect(self, name, window=10, threshold=-0.005):
        return False
