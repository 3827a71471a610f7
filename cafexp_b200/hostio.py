"""Host-side data preparation for the B200 likelihood engine: text inputs -> flat arrays.

Mirrors what the reference does on the host before the hot path starts (file:line relative to the
reference tree):

* Newick parsing, interior-node naming and the reverse-level-order traversal
  (src/clade.cpp:121-133, 255-280, 282-405)
* gene-family table reader and the "exists at root" filter (src/io.cpp:134-215,
  src/gene_family.cpp:60-89, src/cafexp.cpp:189-199)
* max_family_size / max_root_family_size rule (src/user_data.cpp:45-46)
* error-model file reader and get_probs lookup (src/io.cpp:225-270, src/error_model.cpp:31-57)
* root-distribution file reader (src/user_data.cpp:103-115)

Nothing here touches the GPU; the arrays produced are exactly what include/cafe_b200.h takes.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_TOKEN = re.compile(r"\(|\)|[^\s\(\)\:\;\,]+|\:[+-]?[0-9]*\.?[0-9]+(?:[eE][+-]?[0-9]+)?|\,|\;")


class Clade:
    """One node of the species tree (src/clade.h:23-104)."""

    __slots__ = ("parent", "name", "branch_length", "lambda_index", "children")

    def __init__(self, parent: Optional["Clade"] = None):
        self.parent = parent
        self.name = ""
        self.branch_length = 0.0
        self.lambda_index = 0
        self.children: List["Clade"] = []

    @property
    def is_leaf(self) -> bool:
        return not self.children

    @property
    def is_root(self) -> bool:
        return self.parent is None

    def leaf_names(self) -> List[str]:
        if self.is_leaf:
            return [self.name]
        out: List[str] = []
        for c in self.children:
            out.extend(c.leaf_names())
        return out

    def _rename_interior(self) -> None:
        # interior name = sorted leaf names concatenated, propagated to the root (src/clade.cpp:121-133)
        node: Optional[Clade] = self
        while node is not None:
            node.name = "".join(sorted(node.leaf_names()))
            node = node.parent

    def reverse_level_order(self) -> List["Clade"]:
        """Breadth-first from the root, then reversed (src/clade.cpp:255-280)."""
        order: List[Clade] = []
        queue = [self]
        head = 0
        while head < len(queue):
            cur = queue[head]
            head += 1
            order.append(cur)
            queue.extend(cur.children)
        order.reverse()
        return order


def parse_newick(text: str, parse_to_lambdas: bool = False) -> Clade:
    """Token-driven Newick reader with the reference's conventions (src/clade.cpp:282-405)."""
    root = Clade()
    cur = root
    for tok in _TOKEN.findall(text):
        if tok == "(":
            child = Clade(cur)
            cur.children.append(child)
            cur = child
        elif tok == ",":
            if cur is root:
                # outer parentheses omitted: grow a new root above the current node
                new_root = Clade()
                cur.parent = new_root
                new_root.children.append(cur)
                root = new_root
            sib = Clade(cur.parent)
            cur.parent.children.append(sib)
            cur = sib
        elif tok == ")":
            cur = cur.parent
        elif tok == ";":
            break
        elif tok[0] == ":":
            if parse_to_lambdas:
                cur.lambda_index = int(tok[1:], 0)
            else:
                cur.branch_length = float(tok[1:])
        else:
            cur.name = tok
            if cur.parent is not None:
                cur.parent._rename_interior()
    if parse_to_lambdas:
        if root.lambda_index == 0:
            root.lambda_index = 1
        for c in root.reverse_level_order():
            if c.lambda_index < 1:
                raise ValueError(f"Invalid lambda index set for {c.name}")
    else:
        for c in root.reverse_level_order():
            if not c.is_root and c.branch_length <= 0:
                raise ValueError(f"Invalid branch length set for {c.name}")
    return root


@dataclass
class FlatTree:
    """Species tree as the arrays the C ABI takes; node i's children all have index < i, root last."""

    parent: np.ndarray
    child_offset: np.ndarray
    child_list: np.ndarray
    leaf_col: np.ndarray
    branch: np.ndarray
    lambda_index: np.ndarray
    names: List[str]
    leaf_names: List[str] = field(default_factory=list)

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])

    @property
    def n_leaves(self) -> int:
        return len(self.leaf_names)

    @property
    def n_lambdas(self) -> int:
        return int(self.lambda_index.max()) + 1

    @property
    def internal_names(self) -> List[str]:
        return [self.names[i] for i in range(self.n_nodes) if self.leaf_col[i] < 0]

    def longest_branch(self) -> float:
        return float(self.branch[self.parent >= 0].max())


def flatten_tree(root: Clade, lambda_tree: Optional[Clade] = None) -> FlatTree:
    order = root.reverse_level_order()
    index = {id(c): i for i, c in enumerate(order)}
    n = len(order)
    parent = np.full(n, -1, np.int32)
    leaf_col = np.full(n, -1, np.int32)
    branch = np.zeros(n, np.float64)
    lam_idx = np.zeros(n, np.int32)
    child_offset = np.zeros(n + 1, np.int32)
    child_list: List[int] = []
    leaf_names: List[str] = []
    lam_of_name: Dict[str, int] = {}
    if lambda_tree is not None:
        # node NAME -> lambda index - 1 (src/clade.cpp:154-164); names must match (src/clade.cpp:207-222)
        lam_of_name = {c.name: c.lambda_index - 1 for c in lambda_tree.reverse_level_order()}
        if set(lam_of_name) != {c.name for c in order}:
            raise ValueError("The lambda tree structure does not match that of the tree")
    for i, c in enumerate(order):
        if c.parent is not None:
            parent[i] = index[id(c.parent)]
        branch[i] = c.branch_length
        if c.is_leaf:
            leaf_col[i] = len(leaf_names)
            leaf_names.append(c.name)
        for ch in c.children:
            child_list.append(index[id(ch)])
        child_offset[i + 1] = len(child_list)
        if lam_of_name:
            lam_idx[i] = lam_of_name[c.name]
    return FlatTree(parent, child_offset, np.asarray(child_list, np.int32), leaf_col, branch, lam_idx,
                    [c.name for c in order], leaf_names)


def read_tree(path: str, lambda_tree: bool = False) -> Clade:
    with open(path) as fh:
        line = fh.readline()
    tree = parse_newick(line, lambda_tree)
    if tree.is_leaf:
        raise ValueError(f"{path} does not seem to be a valid tree")
    return tree


def read_gene_families(path: str, tree: FlatTree) -> Tuple[List[str], np.ndarray]:
    """CAFE tab format ("Desc<TAB>Family ID<TAB>species...") -> ids and int32 counts [F, n_leaves] in
    the tree's leaf-column order.  Species match is case-insensitive (src/gene_family.h:10-25)."""
    col_of = {name.lower(): i for i, name in enumerate(tree.leaf_names)}
    ids: List[str] = []
    rows: List[List[int]] = []
    with open(path) as fh:
        header = fh.readline().rstrip("\n").rstrip("\r").split("\t")
        species = header[2:]
        cols = [col_of.get(s.lower(), -1) for s in species]
        missing = set(col_of) - {s.lower() for s in species}
        if missing:
            raise ValueError(f"species missing from family table: {sorted(missing)}")
        for line in fh:
            tok = line.rstrip("\n").rstrip("\r").split("\t")
            if len(tok) < 3:
                continue
            ids.append(tok[1])
            row = [0] * tree.n_leaves
            for s, c in enumerate(cols):
                if c >= 0:
                    row[c] = int(tok[2 + s])
            rows.append(row)
    if not rows:
        raise ValueError("No families found")
    return ids, np.asarray(rows, np.int32)


def write_gene_families(path: str, tree: FlatTree, ids: Sequence[str], counts: np.ndarray) -> None:
    with open(path, "w") as fh:
        fh.write("Desc\tFamily ID\t" + "\t".join(tree.leaf_names) + "\n")
        for i, row in zip(ids, counts):
            fh.write("(null)\t" + str(i) + "\t" + "\t".join(str(int(v)) for v in row) + "\n")


def exists_at_root(tree: FlatTree, counts: np.ndarray) -> np.ndarray:
    """Parsimony presence filter (src/gene_family.cpp:60-89): every child of the root must have at
    least one descendant leaf with a non-zero count."""
    n = tree.n_nodes
    present = np.zeros((counts.shape[0], n), bool)
    for v in range(n):
        if tree.leaf_col[v] >= 0:
            present[:, v] = counts[:, tree.leaf_col[v]] > 0
        else:
            ch = tree.child_list[tree.child_offset[v]:tree.child_offset[v + 1]]
            present[:, v] = present[:, ch].any(axis=1)
    root_children = tree.child_list[tree.child_offset[n - 1]:tree.child_offset[n]]
    return present[:, root_children].all(axis=1)


def family_size_limits(counts: np.ndarray) -> Tuple[int, int]:
    """(max_family_size, max_root_family_size) from the largest observed count (src/user_data.cpp:45-46)."""
    largest = int(counts.max())
    mrf = max(30, int(np.rint(largest * 1.25)))
    mf = largest + max(50, largest // 5)
    return mf, mrf


@dataclass
class ErrorModel:
    """Leaf error model (src/error_model.h:29-65)."""

    max_family_size: int
    deviations: List[int]
    dists: List[List[float]]

    def get_probs(self, size: int) -> List[float]:
        # src/error_model.cpp:52-57
        if size >= len(self.dists) and size <= self.max_family_size:
            return self.dists[-1]
        return self.dists[size]

    def dense(self, rows: int) -> np.ndarray:
        """[rows][n_deviations] table indexed by OBSERVED count; rows the reference could not serve are NaN."""
        out = np.full((rows, len(self.deviations)), np.nan)
        for s in range(rows):
            if s < len(self.dists) or s <= self.max_family_size:
                out[s] = self.get_probs(s)
        return out

    def epsilons(self) -> List[float]:
        return sorted({d[-1] for d in self.dists})


def read_error_model(path: str) -> ErrorModel:
    max_cnt = 0
    deviations = [-1, 0, 1]
    dists: List[List[float]] = []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith("max"):
                max_cnt = int(line.split(":")[1].strip())
            elif line.startswith("cnt"):
                deviations = [int(t) for t in line.split(" ")[1:] if t]
            else:
                tok = [t for t in line.split(" ") if t]
                if not tok:
                    continue
                size = int(tok[0])
                probs = [float(t) for t in tok[1:]]
                # src/error_model.cpp:31-50
                if (size == 0 or not dists) and abs(probs[0]) > 0:
                    raise ValueError("Cannot have a non-zero probability for family size 0 for negative deviation")
                if abs(1.0 - sum(probs)) > 0.01 * abs(sum(probs)):
                    raise ValueError("Sum of probabilities must be equal to one")
                if not dists:
                    dists.append(probs)
                while len(dists) <= size:
                    dists.append(list(dists[-1]))
                dists[size] = probs
    return ErrorModel(max_cnt, deviations, dists)


def read_rootdist(path: str) -> Dict[int, int]:
    out: Dict[int, int] = {}
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if len(tok) >= 2:
                out[int(tok[0])] = int(tok[1])
    return out
