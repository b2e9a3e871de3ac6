"""pbml_mantle_convection_b200 -- B200-native surrogate time-stepping path.

Drop-in for the rollout path of agsiddhant/PBML_Mantle_Convection: same module names as the
reference (`pytorch_networks_convae`, `symmetric_layers_torch`, `scaler`, `calculate_profiles`)
so that `from pbml_mantle_convection_b200.pytorch_networks_convae import NewFluidNet, ADNet, TS`
replaces `from pytorch_networks_convae import ...`.
"""
from .pytorch_networks_convae import (ADNet, BoundaryLearnedConvolution2D, FluidLayer, FluidNet, NewFluidNet, TS,  # noqa: F401
                                      Unet, count_parameters)
from .symmetric_layers_torch import SymmetricConv2d  # noqa: F401
from .scaler import scale_var, unscale_var  # noqa: F401
from .calculate_profiles import calc_mlp_profile  # noqa: F401
from .rollout import EnsembleRollout, synthetic_grid, synthetic_T0  # noqa: F401
from . import driver, multigpu  # noqa: F401

__version__ = "0.1.0"
