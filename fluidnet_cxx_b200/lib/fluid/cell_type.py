"""Manta / FluidNet cell-type flags (reference: pytorch/lib/fluid/cell_type.py:5-14).

Flags travel as float32 tensors holding these integers (pytorch/plume.py:132-135)."""
from enum import IntEnum


class CellType(IntEnum):
    TypeNone = 0
    TypeFluid = 1
    TypeObstacle = 2
    TypeEmpty = 4
    TypeInflow = 8
    TypeOutflow = 16
    TypeOpen = 32
    TypeStick = 128
    TypeReserved = 256
