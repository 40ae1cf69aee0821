"""Small host helpers the boundary needs: TP head arithmetic, TP rank info, backend registry.

Mirrors ``div_even`` (reference utils/misc.py:20-26), ``get_tp_info`` (distributed/info.py) and
``Registry`` (utils/registry.py:6-38) -- same names and error behaviour, so the backend reads the
same inside and outside a mini-sglang process.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Generic, Iterable, List, TypeVar

T = TypeVar("T")


def div_even(a: int, b: int, allow_replicate: bool = False) -> int:
    """``a / b`` for head sharding; with ``allow_replicate`` a KV head is replicated when tp > Hkv."""
    if allow_replicate and b > a:
        if b % a != 0:
            raise AssertionError(f"b = {b} must be divisible by a = {a} for KV head replication")
        return 1
    if a % b != 0:
        raise AssertionError(f"a = {a} must be divisible by b = {b}")
    return a // b


@dataclass(frozen=True)
class TPInfo:
    rank: int
    size: int

    def is_primary(self) -> bool:
        return self.rank == 0


_TP = TPInfo(0, 1)


def set_tp_info(rank: int, size: int) -> None:
    global _TP
    if not (0 <= rank < size):
        raise ValueError(f"bad tp rank/size {rank}/{size}")
    _TP = TPInfo(rank, size)


def get_tp_info() -> TPInfo:
    """The reference's TP info when running inside mini-sglang, else the local one."""
    try:  # pragma: no cover - only inside a real mini-sglang process
        from minisgl.distributed import try_get_tp_info as _ref
    except ImportError:  # reference not installed: the stand-alone TP info set by set_tp_info()
        return _TP
    info = _ref()  # errors of the reference propagate
    if info is None:  # reference importable but no engine running in this process (tests, bench.py)
        return _TP
    return TPInfo(info.rank, info.size)


class Registry(Generic[T]):
    """name -> creator; duplicate registration raises KeyError, unknown lookup raises KeyError."""

    def __init__(self, kind: str) -> None:
        self._kind = kind
        self._items: Dict[str, T] = {}

    def register(self, name: str) -> Callable[[T], None]:
        if name in self._items:
            raise KeyError(f"{self._kind} '{name}' is already registered.")

        def deco(item: T) -> None:
            self._items[name] = item

        return deco

    def __getitem__(self, name: str) -> T:
        try:
            return self._items[name]
        except KeyError:
            raise KeyError(f"Unsupported {self._kind}: {name}") from None

    def supported_names(self) -> List[str]:
        return list(self._items)

    def assert_supported(self, names: "str | Iterable[str]") -> None:
        for n in [names] if isinstance(names, str) else names:
            if n not in self._items:
                from argparse import ArgumentTypeError

                raise ArgumentTypeError(
                    f"Unsupported {self._kind}: {n}. Supported items: {self.supported_names()}"
                )
