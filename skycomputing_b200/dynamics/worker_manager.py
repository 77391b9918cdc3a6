"""Ordered worker pool (capability parity with scaelum/dynamics/worker_manager.py:7-79).

``first_rank`` selects the numbering scheme: 1 reproduces the reference (rank 0 reserved for the
central server), 0 is the SPMD scheme used by the launcher (rank 0 is both first pipeline stage
and central server).
"""
from __future__ import annotations

from typing import List, Optional

from .worker import Worker


class WorkerManager:
    def __init__(self, first_rank: int = 1):
        self._worker_pool: List[Worker] = []
        self._first_rank = first_rank

    @property
    def size(self) -> int:
        return len(self._worker_pool)

    @property
    def worker_pool(self) -> List[Worker]:
        return self._worker_pool

    @property
    def first_rank(self) -> int:
        return self._first_rank

    def get_by_id(self, id_str: str, allow_not_found: bool = False) -> Optional[Worker]:
        for worker in self._worker_pool:
            if worker.id == id_str:
                return worker
        if allow_not_found:
            return None
        raise LookupError("Worker with id {} is not found in the worker pool".format(id_str))

    def get_by_rank(self, rank: int) -> Worker:
        for worker in self._worker_pool:
            if worker.rank == rank:
                return worker
        raise LookupError("Worker with rank {} is not found in the worker pool".format(rank))

    def load_worker_pool_from_config(self, config: list) -> None:
        for i, worker_config in enumerate(config):
            cfg = dict(worker_config)
            self._worker_pool.append(Worker(rank=i + self._first_rank, device=cfg.pop("device", i),
                                            **cfg))

    def assign_model_to_worker(self, rank: int, model_config: list) -> None:
        self.get_by_rank(rank).model_config = model_config

    def add_worker(self, worker_id: Optional[str], worker_config: dict) -> None:
        rank = len(self._worker_pool) + self._first_rank
        self._worker_pool.append(Worker(rank=rank, worker_id=worker_id, **dict(worker_config)))

    def _allocate_rank(self) -> None:
        for i, worker in enumerate(self._worker_pool):
            worker.rank = i + self._first_rank

    def remove_worker_by_id(self, id_str: str) -> None:
        worker = self.get_by_id(id_str)
        assert not worker.is_running, "Worker {} is still running".format(id_str)
        self._worker_pool.remove(worker)
        self._allocate_rank()

    def reset_rank_by_order(self) -> None:
        self._worker_pool.sort(key=lambda w: w.order)
        self._allocate_rank()

    def serialize(self) -> list:
        return [w.serialize() for w in self._worker_pool]

    @staticmethod
    def deserialize(data: list, first_rank: int = 1) -> "WorkerManager":
        wm = WorkerManager(first_rank=first_rank)
        for worker_data in data:
            wm.worker_pool.append(Worker.deserialize(worker_data))
        return wm
