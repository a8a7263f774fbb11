"""Process-local warm-start cache (size 1) for expensive objects such as the encoder.

Same contract as distllm/registry.py:44-136: ``registry.register(fn)`` then
``registry.get(fn, **kwargs)`` returns the cached object while the (callable, arguments) pair is
unchanged, and tears the previous object down (optional shutdown callback) when it changes.  It is
what keeps weights, the native handle and its workspace alive across input files in a worker.
"""

from __future__ import annotations

import functools
from typing import Any
from typing import Callable


class _Slot:
    __slots__ = ('shutdown_callback', 'obj', 'arg_hash')

    def __init__(self, shutdown_callback: Callable[[Any], Any] | None) -> None:
        self.shutdown_callback = shutdown_callback
        self.obj: Any = None
        self.arg_hash = 0

    def shutdown(self) -> None:
        if self.obj is None:
            return
        if self.shutdown_callback is not None:
            self.shutdown_callback(self.obj)
        self.obj = None
        self.arg_hash = 0


class RegistrySingleton:
    """At most one live object across all registered factories."""

    _instance: 'RegistrySingleton | None' = None

    def __new__(cls) -> 'RegistrySingleton':
        if cls._instance is None:
            inst = super().__new__(cls)
            inst._slots = {}
            inst._active = None
            cls._instance = inst
        return cls._instance

    def __contains__(self, factory: Callable[..., Any]) -> bool:
        return factory in self._slots

    def register(
        self,
        factory: Callable[..., Any],
        shutdown_callback: Callable[[Any], Any] | None = None,
    ) -> None:
        self._slots.setdefault(factory, _Slot(shutdown_callback))

    def clear(self) -> None:
        for slot in self._slots.values():
            slot.shutdown()
        self._slots = {}
        self._active = None

    def get(self, factory: Callable[..., Any], *args: Any, **kwargs: Any) -> Any:
        if factory not in self._slots:
            raise ValueError(f'Object {getattr(factory, "__name__", factory)} not registered.')
        # arguments must be hashable, exactly like functools.lru_cache keys
        key = hash(functools._make_key((factory, *args), kwargs, typed=False))
        slot = self._slots[factory]
        if self._active is factory and slot.arg_hash == key and slot.obj is not None:
            return slot.obj
        if self._active is not None and self._active in self._slots:
            self._slots[self._active].shutdown()
        obj = factory(*args, **kwargs)
        slot.obj, slot.arg_hash = obj, key
        self._active = factory
        return obj


registry = RegistrySingleton()


def register(shutdown_callback: Callable[[Any], Any] | None = None) -> Callable[[Callable[..., Any]], Callable[..., Any]]:
    """Decorator form: calls to the decorated factory go through the registry."""

    def decorate(factory: Callable[..., Any]) -> Callable[..., Any]:
        registry.register(factory, shutdown_callback)

        @functools.wraps(factory)
        def cached(*args: Any, **kwargs: Any) -> Any:
            return registry.get(factory, *args, **kwargs)

        return cached

    return decorate
