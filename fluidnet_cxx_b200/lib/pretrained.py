"""The shipped ScaleNet (trained_models/ScaleNet_ShortTerm_LongTermLoss of the reference):
its model configuration (`convModel_mconf.pth`) and the MultiScaleNet weights, repackaged as a
flat fp32 archive (data/scalenet_weights.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from .model import FluidNet

WEIGHTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "scalenet_weights.npz")

# convModel_mconf.pth of the shipped model
SCALENET_MCONF = {
    'gravityScale': 0, 'divL1Lambda': 0, 'timeScaleSigma': 1, 'viscosity': 0, 'pL2Lambda': 0, 'inputDim': 2,
    'pL1Lambda': 0, 'normalizeInputThreshold': 1e-05,
    'inputChannels': {'pDiv': False, 'UDiv': False, 'div': True}, 'longTermDivNumSteps': [4, 16],
    'sampleOutsideFluid': False, 'divL2Lambda': 1, 'longTermDivProbability': 0.9, 'maccormackStrength': 0.6,
    'normalizeInput': True, 'divLongTermLambda': 5, 'is3D': False, 'lr': 5e-05, 'buoyancyScale': 0,
    'normalizeInputChan': 'UDiv', 'dt': 0.1, 'model': 'ScaleNet',
}


def load_scalenet(device="cuda", mconf_overrides=None):
    """FluidNet with the shipped ScaleNet weights, in eval mode on `device`."""
    mconf = dict(SCALENET_MCONF)
    mconf.update(mconf_overrides or {})
    net = FluidNet(mconf, dropout=False)
    z = np.load(WEIGHTS)
    sd = {"multiScale." + k: torch.from_numpy(z[k]) for k in z.files}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith("multiScale.") for k in missing)
    return net.to(device).eval(), mconf
