from . import predictors, trainers  # noqa: F401
from .predictors import *  # noqa: F401,F403
from .trainers import *  # noqa: F401,F403
