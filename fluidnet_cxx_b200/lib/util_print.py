"""`lib.summary` (reference: pytorch/lib/util_print.py:6-84): per-layer parameter table of a model; a
training-time convenience kept importable for drivers that call it."""
import torch.nn as nn


def summary(model, input_size=None):
    rows, total = [], 0
    for name, m in model.named_modules():
        if isinstance(m, (nn.Sequential, nn.ModuleList)) or m is model:
            continue
        n = sum(p.numel() for p in m.parameters(recurse=False))
        if n:
            rows.append((name, m.__class__.__name__, n))
            total += n
    width = max([len(r[0]) for r in rows] + [10])
    print(f"{'layer':{width}s}  {'type':18s} {'params':>10s}")
    for name, cls, n in rows:
        print(f"{name:{width}s}  {cls:18s} {n:10d}")
    print(f"{'total':{width}s}  {'':18s} {total:10d}")
    return total
