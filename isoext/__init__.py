"""Drop-in alias: ``import isoext`` resolves to the Blackwell-native package ``isoext_b200`` so
that code (and the reference's own test-suite) written against GuangyanCai/isoext runs unchanged."""
from isoext_b200 import *  # noqa: F401,F403
from isoext_b200 import __all__  # noqa: F401
