"""Index / SVDConfig -- the data-model types of the reference's TT path.

Mirrors `pytens/types.py:19-66`: an index is identified by (name, size) only
(equality and hash ignore `value_choices`), and `SVDConfig` carries the options
of `TensorNetwork.svd`.
"""

from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Sequence, Union

IntOrStr = Union[str, int]
IndexName = IntOrStr
NodeName = IntOrStr


@dataclass(frozen=True, eq=False)
class Index:
    """A named tensor index (pytens/types.py:19-57)."""

    name: IntOrStr
    size: int
    value_choices: Sequence[float] = ()

    def with_new_size(self, new_size: int) -> "Index":
        return Index(self.name, int(new_size))

    def with_new_name(self, name: IntOrStr) -> "Index":
        return Index(name, self.size)

    def with_new_rng(self, rng: Sequence[float]) -> "Index":
        return Index(self.name, self.size, rng)

    def __eq__(self, other: object) -> bool:
        return isinstance(other, Index) and self.name == other.name and self.size == other.size

    def __lt__(self, other: "Index") -> bool:
        return str(self.name) < str(other.name)

    def __hash__(self) -> int:
        return hash((self.name, self.size))

    def to_dict(self) -> dict:
        return dataclasses.asdict(self)

    @classmethod
    def from_dict(cls, data_dict: dict) -> "Index":
        return cls(**data_dict)


@dataclass
class SVDConfig:
    """Options of TensorNetwork.svd (pytens/types.py:60-66)."""

    delta: float = 1e-5
    with_orthonormal: bool = True
    compute_data: bool = True
