"""genb200: B200-native (sm_100a) drop-in for Gen.jl's SMC / importance-sampling hot path.

The directory is named after the reference (`gen.jl_b200/`), which is not a valid Python identifier;
`__graft_entry__.load_package()` imports it under the module name `genb200`.
Product code only: nothing in this package imports anything from `oracle/`.
"""
from ._lib import GenB200Error, LIB_PATH  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (AffineNormalProposal, CategoricalTableProposal, DeviceModel, DeviceParticleFilterState,  # noqa: F401
                  NoChange, UnknownChange)
