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


class NodeInfo:
    """Children / parent side of a dimension-tree node (pytens/types.py:69-81)."""

    def __init__(self, nodes, indices, vals):
        self.nodes = nodes
        self.indices = indices
        self.vals = vals
        self.rank = 0


class DimTreeNode:
    """Node of the dimension tree of a tree network (pytens/types.py:84-120): `indices` are the free
    indices below the node, `down_info.nodes` its children, `up_info.nodes` its parent (0 or 1)."""

    def __init__(self, node, indices, free_indices, up_info: NodeInfo, down_info: NodeInfo):
        self.node = node
        self.indices = indices
        self.free_indices = free_indices
        self.up_info = up_info
        self.down_info = down_info
        self.perm = list(range(len(free_indices) + len(down_info.nodes) + len(up_info.nodes)))

    def __lt__(self, other: "DimTreeNode") -> bool:
        return sorted(self.indices) < sorted(other.indices)

    def preorder(self):
        out = [self]
        for child in self.down_info.nodes:
            out.extend(child.preorder())
        return out
