"""``from helmpy.core.nr import nr`` (reference helmpy/core/nr.py:506) -> the B200 build."""
from helmpy_b200.core.nr import nr, nr_batch  # noqa: F401
