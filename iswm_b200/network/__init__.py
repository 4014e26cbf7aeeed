"""Drop-in for the reference's `network` package: `network.modeling.deeplabv3plus_resnet50(...)`."""
from . import modeling  # noqa: F401
from .modeling import *  # noqa: F401,F403
