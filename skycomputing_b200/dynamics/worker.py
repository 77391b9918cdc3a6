"""Worker record (capability parity with scaelum/dynamics/worker.py:8-97).

New fields for the SPMD design: ``device`` (physical GPU index that runs this worker's span) -
the reference moves the whole Worker object to a new RPC rank when the allocator re-orders the
pipeline, which is only right for simulated heterogeneity; here the pipeline
position is ``order``/``rank`` and the physical GPU is carried explicitly.
"""
from __future__ import annotations

import uuid
from typing import Optional


class Worker:
    def __init__(self, rank: int, name: str, server_config: Optional[dict] = None,
                 worker_id: Optional[str] = None, order: Optional[int] = None,
                 model_config: Optional[list] = None, extra_config: Optional[dict] = None,
                 is_running: bool = False, device: Optional[int] = None,
                 layer_range: Optional[tuple] = None, chunks: Optional[list] = None) -> None:
        self._rank = rank
        self._name = name
        self._is_running = is_running
        self._order = order
        self._worker_id = str(uuid.uuid4()) if worker_id is None else worker_id
        self._server_config = server_config or {}
        self._model_config = model_config
        self._extra_config = extra_config or {}
        self._device = device
        self._layer_range = layer_range
        # looped (virtual-stage) pipelines: this worker runs several NON-adjacent layer spans
        # [(begin, end), ...]; virtual stage k of the pipeline is chunk k // D of worker k % D
        self._chunks = chunks

    rank = property(lambda self: self._rank)
    id = property(lambda self: self._worker_id)
    name = property(lambda self: self._name)
    model_config = property(lambda self: self._model_config)
    server_config = property(lambda self: self._server_config)
    extra_config = property(lambda self: self._extra_config)
    is_running = property(lambda self: self._is_running)
    order = property(lambda self: self._order)
    device = property(lambda self: self._device)
    layer_range = property(lambda self: self._layer_range)
    chunks = property(lambda self: self._chunks)

    @property
    def env_config(self) -> dict:
        return dict(self._server_config)

    @rank.setter
    def rank(self, rank: int) -> None:
        self._rank = rank

    @order.setter
    def order(self, order: int) -> None:
        self._order = order

    @is_running.setter
    def is_running(self, status: bool) -> None:
        self._is_running = status

    @model_config.setter
    def model_config(self, config: list) -> None:
        self._model_config = config

    @device.setter
    def device(self, device: Optional[int]) -> None:
        self._device = device

    @layer_range.setter
    def layer_range(self, rng: Optional[tuple]) -> None:
        self._layer_range = rng

    @chunks.setter
    def chunks(self, chunks: Optional[list]) -> None:
        self._chunks = chunks

    def serialize(self) -> dict:
        return dict(self.__dict__)

    @staticmethod
    def deserialize(data: dict) -> "Worker":
        return Worker(**{k.lstrip("_"): v for k, v in data.items()})
