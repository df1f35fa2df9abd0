from . import loggers, monitor  # noqa: F401
from .loggers import *  # noqa: F401,F403
from .monitor import Monitor  # noqa: F401
