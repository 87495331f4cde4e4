from isoext_b200.sdf import *  # noqa: F401,F403
from isoext_b200.sdf import __all__  # noqa: F401
