"""`lib.plotField` (reference: pytorch/lib/plot_field.py): matplotlib visualisation of a training sample.
Plotting is outside the B200 path; the name is kept importable and fails clearly when called."""


def plotField(*args, **kwargs):
    raise NotImplementedError("lib.plotField (training-time matplotlib plot) is outside the B200 per-timestep path")
