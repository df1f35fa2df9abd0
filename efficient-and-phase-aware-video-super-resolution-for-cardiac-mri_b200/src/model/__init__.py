from . import losses, metrics, nets  # noqa: F401
from .losses import *  # noqa: F401,F403
from .metrics import *  # noqa: F401,F403
from .nets import *  # noqa: F401,F403
