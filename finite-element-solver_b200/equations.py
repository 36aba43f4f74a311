"""Equation specs the drivers accept (ref fem/equations.py): plain data."""
from dataclasses import dataclass


@dataclass(frozen=True)
class LinearElastic:
    E: float = 1.0
    nu: float = 0.3
    source: object = None


@dataclass(frozen=True)
class Poisson:
    source: object = None
