"""Reference-facing package: the names `pytorch/lib/__init__.py:1-7` exports that the
per-timestep path uses.  Import as `fluidnet_cxx_b200.lib`, or as top-level `lib` through
`fluidnet_cxx_b200.compat.install()` so the reference drivers run unchanged."""
from . import fluid
from .dataset_load import FluidNetDataset
from .util_print import summary
from .simulate import simulate, setConstVals
from .multi_scale_net import MultiScaleNet
from .model import FluidNet
from .plot_field import plotField
from .argument_parser import SmartFormatter
from .host_pipeline import HostStepPipeline      # not in the reference: host-resident states, copies overlapped
