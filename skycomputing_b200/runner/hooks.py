"""Hook protocol of the training loop.

Capability parity with scaelum/runner/hooks.py:5-58 - the fourteen callback names a ``Runner`` may
fire and the ``every_n_*`` / ``end_of_epoch`` cadence predicates - but built differently: the base
class carries no hand-written stubs.  The callback surface is a table (``STAGES`` x ``SCOPES`` x
``MODES``) from which the methods are generated, so that

* the six generic callbacks (``before_run`` ... ``after_iter``) are no-ops,
* the eight mode-specific callbacks (``before_train_epoch`` ... ``after_val_iter``) forward to the
  generic callback of the same stage and scope unless a subclass overrides them,
* ``Hook.fire(runner, name)`` is the single entry the Runner uses; it validates the name against
  the table, which turns a typo in a runner / hook into an error instead of a silent no-op,
* ``Hook.overrides()`` reports which callbacks a concrete hook really implements (the Runner
  skips hooks that do not implement the callback being fired - one attribute test per hook
  instead of a Python call on the hot loop).
"""
from __future__ import annotations

from typing import FrozenSet, Tuple

STAGES: Tuple[str, ...] = ("before", "after")
SCOPES: Tuple[str, ...] = ("run", "epoch", "iter")
MODES: Tuple[str, ...] = ("train", "val")

GENERIC_CALLBACKS: Tuple[str, ...] = tuple(f"{st}_{sc}" for sc in SCOPES for st in STAGES)
MODE_CALLBACKS: Tuple[str, ...] = tuple(
    f"{st}_{mode}_{sc}" for sc in SCOPES[1:] for mode in MODES for st in STAGES)
ALL_CALLBACKS: FrozenSet[str] = frozenset(GENERIC_CALLBACKS + MODE_CALLBACKS)


def _noop(name: str):
    def callback(self, runner):
        return None

    callback.__name__ = callback.__qualname__ = name
    callback.__doc__ = f"generic callback ``{name}``: nothing to do in the base class"
    callback._sky_default = True
    return callback


def _forward(name: str, target: str):
    def callback(self, runner):
        return getattr(self, target)(runner)

    callback.__name__ = callback.__qualname__ = name
    callback.__doc__ = f"``{name}`` falls through to ``{target}`` unless overridden"
    callback._sky_default = True
    return callback


class Hook:
    """Base class of everything registered in ``HOOKS``; see the module docstring."""

    @classmethod
    def overrides(cls) -> FrozenSet[str]:
        """Names of the callbacks this class (or a base other than ``Hook``) implements.  A
        mode-specific callback counts as implemented when its generic target is."""
        own = {n for n in ALL_CALLBACKS if not getattr(getattr(cls, n), "_sky_default", False)}
        for n in MODE_CALLBACKS:
            st, _mode, sc = n.split("_")
            if f"{st}_{sc}" in own:
                own.add(n)
        return frozenset(own)

    def fire(self, runner, name: str):
        if name not in ALL_CALLBACKS:
            raise AttributeError(f"'{name}' is not a hook callback (known: {sorted(ALL_CALLBACKS)})")
        return getattr(self, name)(runner)

    # ---- cadence predicates (0-based counters on the runner, like the reference) -------------
    @staticmethod
    def _every(counter: int, n) -> bool:
        return bool(n) and n > 0 and (counter + 1) % n == 0

    def every_n_epochs(self, runner, n) -> bool:
        return self._every(runner.epoch, n)

    def every_n_inner_iters(self, runner, n) -> bool:
        return self._every(runner.inner_iter, n)

    def every_n_iters(self, runner, n) -> bool:
        return self._every(runner.iter, n)

    def end_of_epoch(self, runner) -> bool:
        loader = getattr(runner, "data_loader", None)
        return loader is not None and runner.inner_iter + 1 == len(loader)


for _name in GENERIC_CALLBACKS:
    setattr(Hook, _name, _noop(_name))
for _name in MODE_CALLBACKS:
    _st, _mode, _sc = _name.split("_")
    setattr(Hook, _name, _forward(_name, f"{_st}_{_sc}"))
del _name, _st, _mode, _sc
