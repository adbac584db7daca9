"""``from helmpy.core.classes import ...`` (reference helmpy/core/classes.py:116, 180) -> the B200 build."""
from helmpy_b200.core.classes import *  # noqa: F401,F403
from helmpy_b200.core.classes import CaseData, create_case_data_object_from_xlsx  # noqa: F401
