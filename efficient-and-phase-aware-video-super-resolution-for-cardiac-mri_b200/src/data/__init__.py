from . import datasets, dataloader, transforms  # noqa: F401
from .datasets import *  # noqa: F401,F403
from .dataloader import Dataloader  # noqa: F401
