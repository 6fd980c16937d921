"""Plugin surface of the B200 path: the six named registries the reference's scripts look detectors up in.

Contract mirrored from R/visualDet3D/networks/utils/registry.py:21-50 (behaviour, not code):
  * objects are keyed by their ``__name__``; only classes and plain functions are accepted (TypeError otherwise);
  * registering a taken name raises KeyError unless ``force=True``;
  * ``registry[name]`` raises KeyError for unknown names, ``registry.get(name)`` returns None;
  * ``@registry.register_module`` is a bare decorator returning the object unchanged.
Detectors are then built exactly like the reference does: ``DETECTOR_DICT[cfg.detector.name](cfg.detector)``
(R/scripts/train.py:87, R/scripts/eval.py:37).
"""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, Iterator, Optional


class Registry:
    __slots__ = ("_name", "_items")

    def __init__(self, name: str):
        self._name = name
        self._items: Dict[str, Any] = {}

    # -- introspection -----------------------------------------------------------------------------------
    @property
    def name(self) -> str:
        return self._name

    @property
    def module_dict(self) -> Dict[str, Any]:
        return self._items

    def __repr__(self) -> str:
        return f"Registry(name={self._name}, items={list(self._items)})"

    def __iter__(self) -> Iterator[str]:
        return iter(self._items)

    def __len__(self) -> int:
        return len(self._items)

    def __contains__(self, key: str) -> bool:
        return key in self._items

    # -- lookup --------------------------------------------------------------------------------------------
    def __getitem__(self, key: str) -> Any:
        return self._items[key]          # KeyError for unknown names, like the reference

    def get(self, key: str) -> Optional[Any]:
        return self._items.get(key)

    # -- registration ----------------------------------------------------------------------------------------
    def _register_module(self, obj: Any, force: bool = False) -> None:
        if not (inspect.isclass(obj) or inspect.isfunction(obj)):
            raise TypeError(f"module must be a class or function, but got {type(obj)}")
        key = obj.__name__
        if key in self._items and not force:
            raise KeyError(f"{key} is already registered in {self._name}")
        self._items[key] = obj

    def register_module(self, obj: Callable = None):
        self._register_module(obj)
        return obj


DATASET_DICT, BACKBONE_DICT, DETECTOR_DICT = Registry("datasets"), Registry("backbones"), Registry("detectors")
PIPELINE_DICT, AUGMENTATION_DICT, SAMPLER_DICT = Registry("pipelines"), Registry("augmentation"), Registry("sampler")


def install_into_reference(force: bool = True):
    """Put every B200 detector / pipeline into the REFERENCE's registries (when `visualDet3D` is importable) so the
    reference's own scripts/eval.py and scripts/train.py pick them up by `cfg.detector.name` with no edit."""
    from visualDet3D.networks.utils import registry as ref   # ImportError if the reference is not on sys.path
    for cls in DETECTOR_DICT.module_dict.values():
        ref.DETECTOR_DICT._register_module(cls, force=force)
    for fn in PIPELINE_DICT.module_dict.values():
        ref.PIPELINE_DICT._register_module(fn, force=force)
    return ref
