from .optimizer_utils import *  # noqa: F401,F403
from .apts_utils import *  # noqa: F401,F403
