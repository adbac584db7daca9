"""Mirror of reference helmpy/core/__init__.py:1-4: the four entry points, as functions,
on ``helmpy.core`` (importing the submodules first and then rebinding the names, so that
``helmpy.core.helm`` is the function and ``sys.modules['helmpy.core.helm']`` the module,
exactly as in the reference), plus the batch calls of the B200 build."""
from helmpy.core.helm import helm  # noqa: F401
from helmpy.core.classes import create_case_data_object_from_xlsx  # noqa: F401
from helmpy.core.nr import nr  # noqa: F401
from helmpy.core.nr_ds import nr_ds  # noqa: F401
from helmpy.core.helm import helm_batch, non_islanding_branches  # noqa: F401
from helmpy.core.nr import nr_batch  # noqa: F401
from helmpy_b200.core.matpower import create_case_data_object_from_matpower, matpower_to_xlsx  # noqa: F401
