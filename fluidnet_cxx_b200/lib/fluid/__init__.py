"""`lib.fluid` -- same exported names as the reference (pytorch/lib/fluid/__init__.py:1-14)."""
from .cell_type import CellType
from .ops import (getDx, getCentered, setWallBcs, flagsToOccupancy, velocityDivergence, velocityUpdate,
                  addBuoyancy, addGravity, addViscosity, emptyDomain, correctScalar, advectScalar, advectVelocity,
                  solveLinearSystemJacobi, setConstVals, outputFields, outputFieldsToHost,
                  OUTPUT_PLANES)
from .extras import setWallBcsStick, createCylinder, createBox2D
from .init_conditions import createPlumeBCs, createRayleighTaylorBCs
