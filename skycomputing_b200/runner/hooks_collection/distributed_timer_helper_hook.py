"""Deletes the file timer before / after a run (parity with
scaelum/runner/hooks_collection/distributed_timer_helper_hook.py:10-16)."""
from ...registry import HOOKS
from ..hooks import Hook


@HOOKS.register_module
class DistributedTimerHelperHook(Hook):
    def before_run(self, runner):
        if runner.is_rank0:
            runner._timer.clean_prev_file()

    def after_run(self, runner):
        if runner.is_rank0:
            runner._timer.clean_prev_file()
