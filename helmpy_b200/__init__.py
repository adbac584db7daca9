"""helmpy_b200 — B200-native drop-in for HELMpy's HELM + Pade hot path.

``import helmpy_b200 as helmpy`` gives the reference's four entry points
(reference helmpy/core/__init__.py:1-4) plus ``helm_batch``.
"""
from helmpy_b200.core import *  # noqa: F401,F403
from helmpy_b200.core import (helm, helm_batch, non_islanding_branches,  # noqa: F401
                              create_case_data_object_from_xlsx, nr, nr_ds,
                              create_case_data_object_from_matpower, matpower_to_xlsx, nr_batch)
