"""``helmpy`` — the reference's package name, bound to the B200 implementation.

An unmodified user script (``import helmpy``; ``helmpy.helm(case, ...)``,
``helmpy.create_case_data_object_from_xlsx(path)``, ``helmpy.nr(...)``,
``helmpy.nr_ds(...)``; reference helmpy/__init__.py:1) runs on the GPU path when
this repository precedes the reference on ``sys.path``.  Everything is
re-exported from ``helmpy_b200``; there is no second implementation here.
"""
from helmpy.core import *  # noqa: F401,F403  (reference helmpy/__init__.py:1)
from helmpy.core import (helm, nr, nr_ds, create_case_data_object_from_xlsx,  # noqa: F401
                         helm_batch, nr_batch, non_islanding_branches,
                         create_case_data_object_from_matpower, matpower_to_xlsx)

IMPLEMENTATION = 'helmpy_b200'
