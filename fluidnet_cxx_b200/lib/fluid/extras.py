"""Members of the reference surface that are OFF the plume / Rayleigh-Taylor path
(`viscosity: 0`, no `flags_stick`; partly broken upstream).  They are kept
importable because `lib.fluid` exports them (pytorch/lib/fluid/__init__.py:4,9,10); the
geometry helpers are setup-time torch code; setWallBcsStick is an explicit "not on the B200 path"
error rather than a silent fallback (the reference version cannot run either)."""
import torch

from .cell_type import CellType


def setWallBcsStick(U, flags, flags_stick):
    raise NotImplementedError("setWallBcsStick (set_wall_bcs_stick.py) is outside the B200 hot path "
                              "(the reference version raises NameError at :62)")


def createCylinder(batch_dict, centerX, centerY, radius):
    """Marks a disc of Obstacle cells (geometry_utils.py:4-33)."""
    flags = batch_dict['flags']
    H, W = flags.size(3), flags.size(4)
    X = torch.arange(W, device=flags.device).view(1, W).expand(H, W)
    Y = torch.arange(H, device=flags.device).view(H, 1).expand(H, W)
    mask = ((X - centerX).pow(2) + (Y - centerY).pow(2)) <= radius * radius
    flags[:, :, :, mask] = float(CellType.TypeObstacle)
    batch_dict['flags'] = flags


def createBox2D(batch_dict, x0, x1, y0, y1):
    """Marks the box [x0,x1) x [y0,y1) as Obstacle (geometry_utils.py:35-64)."""
    flags = batch_dict['flags']
    flags[:, :, :, y0:y1, x0:x1] = float(CellType.TypeObstacle)
    batch_dict['flags'] = flags
