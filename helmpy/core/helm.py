"""``from helmpy.core.helm import helm`` (reference helmpy/core/helm.py:896) -> the B200 build."""
from helmpy_b200.core.helm import *  # noqa: F401,F403
from helmpy_b200.core.helm import helm, helm_batch, validate_arguments, non_islanding_branches  # noqa: F401
