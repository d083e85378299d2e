"""Mantaflow `.bin` state files (reference: pytorch/lib/load_manta_data.py:4-41, format written by
solver_cpp/test/load_manta_data.h:18-116): the data format on the input side of the step.

    header  5 x int32   (transpose, nx, ny, nz, is3D)
    fp32    Ux[n], Uy[n], p[n]           n = nx*ny*nz, x fastest
    fp32    Uz[n]                        only when is3D == 1
    int32   flags[n]                     Manta cell types
    fp32    density[n]

`loadMantaFile` returns what the reference returns -- (p, U, flags, density, is3D) as 5-D float32
tensors (1, C, nz, ny, nx), flags converted to float -- but reads with numpy.frombuffer instead of
struct.unpack of one Python float per cell.  `saveMantaFile` writes the same format (fixtures)."""
import numpy as np
import torch


def loadMantaFile(fname):
    with open(fname, 'rb') as f:
        raw = f.read()
    head = np.frombuffer(raw, dtype='<i4', count=5)
    nx, ny, nz = int(head[1]), int(head[2]), int(head[3])
    is3D = bool(head[4] == 1)
    n = nx * ny * nz
    need = 20 + 4 * n * (6 if is3D else 5)
    assert n > 0 and len(raw) >= need, f"{fname}: truncated Manta file ({len(raw)} bytes, header needs {need})"
    off = 20

    def take(dtype):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off)
        off += 4 * n
        return a
    Ux, Uy, p = take('<f4'), take('<f4'), take('<f4')
    Uz = take('<f4') if is3D else None
    flags = take('<i4').astype(np.float32)
    density = take('<f4')

    def t5(a):
        return torch.from_numpy(np.array(a, dtype=np.float32)).view(1, 1, nz, ny, nx)
    comps = [t5(Ux), t5(Uy)] + ([t5(Uz)] if is3D else [])
    U = torch.cat(comps, 1).contiguous()
    return t5(p), U, t5(flags), t5(density), is3D


def saveMantaFile(fname, p, U, flags, density, transpose=0):
    """inverse of loadMantaFile"""
    is3D = U.size(1) == 3
    nz, ny, nx = (int(s) for s in p.shape[2:])
    with open(fname, 'wb') as f:
        f.write(np.array([transpose, nx, ny, nz, int(is3D)], dtype='<i4').tobytes())
        for a in (U[0, 0], U[0, 1], p[0, 0]):
            f.write(a.detach().cpu().contiguous().numpy().astype('<f4').tobytes())
        if is3D:
            f.write(U[0, 2].detach().cpu().contiguous().numpy().astype('<f4').tobytes())
        f.write(flags[0, 0].detach().cpu().contiguous().numpy().astype('<i4').tobytes())
        f.write(density[0, 0].detach().cpu().contiguous().numpy().astype('<f4').tobytes())
