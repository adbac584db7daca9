"""Result reporting (reference helm.py:662-850) — SURVEY §8(f) row 3, not built yet."""


def report(*a, **k):
    raise NotImplementedError("detailed_run_print / save_results reporting is not built yet")
