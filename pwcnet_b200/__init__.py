"""pwcnet_b200: B200-native (sm_100a) compute path behind the Python call surface of
daigo0927/pwcnet.  See DESIGN.md; the C ABI is in include/pwc_b200.h."""
from .model import PWCDCNet, PWCNet, glorot_init, layer_table   # noqa: F401
from .losses import L1loss, L2loss, EPE, multiscale_loss, multirobust_loss   # noqa: F401
from .pipeline import InferenceStream, TrainStream   # noqa: F401
from ._abi import PwcError   # noqa: F401
from .train import Trainer, piecewise_lr   # noqa: F401
from . import ops, ops_bwd, modules   # noqa: F401

__version__ = "0.1.0"
