from isoext_b200.utils import *  # noqa: F401,F403
from isoext_b200.utils import __all__  # noqa: F401
