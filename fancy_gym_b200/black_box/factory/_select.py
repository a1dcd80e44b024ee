"""Type-string dispatch shared by the four factories (fancy_gym/black_box/factory/*.py).

Each factory of the reference is an if/elif chain over a lower-cased type string that ends in a
ValueError naming the supported types; some strings are reserved and raise NotImplementedError.
Here that is one table-driven selector.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, Optional


class TypeSelector:
    def __init__(self, what: str, builders: Dict[str, Callable[..., Any]], reserved: Iterable[str] = (),
                 advertised: Optional[Iterable[str]] = None, requires: Optional[Dict[str, Callable[..., None]]] = None):
        self.what = what
        self.builders = dict(builders)
        self.reserved = frozenset(reserved)
        self.advertised = list(advertised) if advertised is not None else list(builders)
        self.requires = requires or {}

    def build(self, type_name: str, *args, **kwargs):
        key = str(type_name).lower()
        if key in self.reserved:
            raise NotImplementedError()
        make = self.builders.get(key)
        if make is None:
            raise ValueError(f"Specified {self.what} type {key} not supported, please choose one of {self.advertised}.")
        check = self.requires.get(key)
        if check is not None:
            check(*args, **kwargs)
        return make(*args, **kwargs)
