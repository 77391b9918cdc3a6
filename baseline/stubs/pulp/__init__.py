"""Import-only stub: PuLP/CBC cannot be installed offline, so the reference's optimal_allocate
(MILP) stays unusable; even/dynamic allocation do not touch it."""


def __getattr__(name):
    raise RuntimeError("pulp is unavailable in this environment (offline)")
