"""Python-file configs (capability parity with scaelum/config/config.py:10-78).

A config is a ``.py`` file whose public module-level names become the keys of an attribute-dict.
``base = "other.py"`` gives single-level inheritance (child overrides base, shallow).  The file is
executed with ``runpy`` (no ``sys.modules`` pollution); private names, modules, classes and
callables are dropped, so helper functions defined in a config do not leak into it.
"""
from __future__ import annotations

import inspect
import os.path as osp
import runpy
import sys
from typing import Any, Dict


class Config(dict):
    """Dictionary with attribute access."""

    def __missing__(self, name):
        raise KeyError(name)

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def update(self, config=(), **kwargs) -> "Config":  # type: ignore[override]
        for k, v in dict(config, **kwargs).items():
            self[k] = v
        return self

    @staticmethod
    def from_dict(data: Dict[str, Any]) -> "Config":
        return Config().update(data)


def _py2dict(py_path: str) -> Dict[str, Any]:
    assert py_path.endswith(".py"), f"config must be a python file, got {py_path}"
    py_path = osp.abspath(py_path)
    if not osp.isfile(py_path):
        raise FileNotFoundError(py_path)
    parent_dir = osp.dirname(py_path)
    inserted = parent_dir not in sys.path
    if inserted:
        sys.path.insert(0, parent_dir)
    try:
        namespace = runpy.run_path(py_path, run_name="__skyconfig__")
    finally:
        if inserted and parent_dir in sys.path:
            sys.path.remove(parent_dir)
    return {
        k: v
        for k, v in namespace.items()
        if not k.startswith("_")
        and not inspect.ismodule(v)
        and not inspect.isclass(v)
        and not inspect.isfunction(v)
    }


def load_config(file_path: str) -> Config:
    config = Config(_py2dict(file_path))
    base = config.pop("base", None)
    if base:
        base_path = osp.join(osp.dirname(osp.abspath(file_path)), base)
        config = Config(_py2dict(base_path)).update(config)
    return config
