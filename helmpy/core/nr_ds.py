"""``from helmpy.core.nr_ds import nr_ds`` (reference helmpy/core/nr_ds.py:547) -> the B200 build."""
from helmpy_b200.core.nr import nr_ds  # noqa: F401
