"""Shared ``name -> (ConfigCls, Cls)`` lookup used by the five plugin families."""

from __future__ import annotations

from typing import Any
from typing import Mapping


def build_from_strategies(
    family: str,
    strategies: Mapping[str, tuple[type, type]],
    kwargs: Mapping[str, Any],
) -> Any:
    """Instantiate ``Cls(ConfigCls(**kwargs))`` for ``kwargs['name']``.

    Unknown names raise ``ValueError`` listing what is available, like every ``get_*`` factory
    in the reference (e.g. distllm/embed/poolers/__init__.py:48-57).
    """
    name = kwargs.get('name', '')
    entry = strategies.get(name)
    if not entry:
        raise ValueError(f'Unknown {family} name: {name}. Available: {set(strategies.keys())}')
    config_cls, cls = entry
    return cls(config_cls(**kwargs))
