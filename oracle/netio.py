"""(De)serialisation of tensor networks for the golden fixtures -- TEST INFRASTRUCTURE ONLY.

`oracle/make_golden_trees.py` stores networks built with the reference's classes; the tests rebuild
them with the classes of `tensor_networks_b200.algs` (same container layout: `tn.network` is an
nx.Graph whose nodes carry "tensor" = Tensor(value, indices)).  A network is stored in an npz as
`<prefix>meta` (JSON: node names in insertion order, their index (name, size) lists, edges) plus
`<prefix>v<j>` (value of node j).
"""

from __future__ import annotations

import json

import numpy as np


def _enc(name):
    if isinstance(name, (int, np.integer)):
        return ["i", int(name)]
    return ["s", str(name)]


def _dec(pair):
    return int(pair[1]) if pair[0] == "i" else str(pair[1])


def structure(tn):
    """Hashable description of a network: nodes in order with their indices, and the edge set."""
    nodes = []
    for n, data in tn.network.nodes(data=True):
        nodes.append((_dec(_enc(n)), tuple((_dec(_enc(i.name)), int(i.size)) for i in data["tensor"].indices)))
    edges = sorted(tuple(sorted((repr(_dec(_enc(a))), repr(_dec(_enc(b)))))) for a, b in tn.network.edges())
    return nodes, edges


def pack(tn, prefix: str, with_values: bool = True) -> dict:
    meta = {"nodes": [], "edges": [[_enc(a), _enc(b)] for a, b in tn.network.edges()],
            # neighbour order per node (nx adjacency = insertion order): the tree walks of the reference
            # (orthonormalize, round, dimension_tree) visit neighbours in this order
            "adj": [[_enc(n), [_enc(m) for m in tn.network.neighbors(n)]] for n in tn.network.nodes]}
    out = {}
    for j, (n, data) in enumerate(tn.network.nodes(data=True)):
        t = data["tensor"]
        meta["nodes"].append({"name": _enc(n), "indices": [[_enc(i.name), int(i.size)] for i in t.indices]})
        if with_values:
            v = t.value
            if hasattr(v, "detach"):
                v = v.detach().cpu().numpy()
            out[f"{prefix}v{j}"] = np.array(v)
    out[f"{prefix}meta"] = np.array(json.dumps(meta))
    return out


def unpack(z, prefix: str, TensorNetwork, Tensor, Index):
    meta = json.loads(str(z[f"{prefix}meta"]))
    tn = TensorNetwork()
    for j, nd in enumerate(meta["nodes"]):
        inds = [Index(_dec(nm), int(sz)) for nm, sz in nd["indices"]]
        val = np.array(z[f"{prefix}v{j}"]) if f"{prefix}v{j}" in z else np.array([])
        tn.add_node(_dec(nd["name"]), Tensor(val, inds))
    for a, b in _edge_sequence(meta):
        tn.add_edge(a, b)
    return tn


def _edge_sequence(meta):
    """An insertion order of the edges that reproduces the stored per-node neighbour orders."""
    if "adj" not in meta:
        return [(_dec(a), _dec(b)) for a, b in meta["edges"]]
    todo = {repr(_dec(n)): [_dec(m) for m in nbrs] for n, nbrs in meta["adj"]}
    names = {repr(_dec(n)): _dec(n) for n, _ in meta["adj"]}
    seq = []
    remaining = sum(len(v) for v in todo.values()) // 2
    while remaining:
        progressed = False
        for key, nbrs in todo.items():
            if not nbrs:
                continue
            other = nbrs[0]
            if todo[repr(other)] and repr(todo[repr(other)][0]) == key:
                seq.append((names[key], other))
                nbrs.pop(0)
                todo[repr(other)].pop(0)
                remaining -= 1
                progressed = True
        if not progressed:
            raise ValueError("inconsistent adjacency orders in fixture")
    return seq


def meta_structure(z, prefix: str):
    """The structure() of a stored network without rebuilding it."""
    meta = json.loads(str(z[f"{prefix}meta"]))
    nodes = [(_dec(nd["name"]), tuple((_dec(nm), int(sz)) for nm, sz in nd["indices"])) for nd in meta["nodes"]]
    edges = sorted(tuple(sorted((repr(_dec(a)), repr(_dec(b))))) for a, b in meta["edges"])
    return nodes, edges


def dense_in_order(tn, names):
    """Contract `tn` and return the dense array with its axes ordered like `names` (index names)."""
    t = tn.contract()
    val = t.value
    if hasattr(val, "detach"):
        val = val.detach().cpu().numpy()
    have = [i.name for i in t.indices]
    return np.transpose(np.asarray(val), [have.index(n) for n in names])
