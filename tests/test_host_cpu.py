"""CPU-only checks of the host side: the C ABI library loads and exports what include/cafe_b200.h declares,
refuses to compute without a GPU, and the schedule the kernels walk reproduces inference_prune when it is
interpreted on the CPU with oracle-built matrices (this validates slot allocation, spilling and child order)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from cafexp_b200 import engine, hostio, synth
from oracle import binding as orc

OPS = {0: "LEAF_SET", 1: "LEAF_MUL", 2: "GEMM_SET", 3: "GEMM_MUL", 4: "SPILL", 5: "FILL", 6: "RESCALE", 7: "ROOT"}


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    header = open(os.path.join(ROOT, "include", "cafe_b200.h")).read()
    declared = set(re.findall(r"\b(cafe_b200_[a-z_]+)\s*\(", header))
    declared -= {"cafe_b200_limits"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/cafe_b200.h but not exported"
    assert declared == set(engine.EXPORTS) | {"cafe_b200_plan_schedule"}
    assert lib.cafe_b200_abi_version() == 2
    lim = engine.limits()
    assert lim["families_per_tile"] == 48 and lim["max_matrix_size"] >= 512


@pytest.mark.skipif(engine.load_library().cafe_b200_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    tree = hostio.flatten_tree(hostio.parse_newick("(A:1,B:1);"))
    with pytest.raises(engine.CafeB200Error, match="no CUDA device"):
        engine.Engine(tree, np.array([[1, 2]], np.int32), 10, 8)


def test_create_rejects_bad_input():
    lib = engine.load_library()
    handle = C.c_void_p()
    assert lib.cafe_b200_create(C.byref(handle), None, None, 0, 0, 0, 0, 0) == -1


def plan(tree, n_slots):
    lib = engine.load_library()
    lib.cafe_b200_plan_schedule.restype = C.c_int
    ip = C.POINTER(C.c_int)
    ts, _keep = engine.tree_struct(tree)
    cap = 16 * tree.n_nodes + 64
    ops = np.zeros((cap, 4), np.int32)
    n_ops = C.c_int()
    n_spill = C.c_int()
    rc = lib.cafe_b200_plan_schedule(C.byref(ts), n_slots, ops.ctypes.data_as(ip), cap, C.byref(n_ops), C.byref(n_spill))
    assert rc == 0
    return ops[:n_ops.value], n_spill.value


def plan_program(tree):
    """(ops [n][7], leaf node ids, stack depth) of the pruning kernel's stack-machine program."""
    lib = engine.load_library()
    ip = C.POINTER(C.c_int)
    ts, _keep = engine.tree_struct(tree)
    cap = 4 * tree.n_nodes + 8
    ops = np.zeros((cap, 7), np.int32)
    leaves = np.zeros(cap, np.int32)
    n_ops, n_leaf, depth = C.c_int(), C.c_int(), C.c_int()
    rc = lib.cafe_b200_plan_program(C.byref(ts), ops.ctypes.data_as(ip), cap, C.byref(n_ops), leaves.ctypes.data_as(ip), cap, C.byref(n_leaf),
                                    C.byref(depth))
    assert rc == 0
    return ops[:n_ops.value], leaves[:n_leaf.value], depth.value


def interpret_program(tree, ops, leaves, depth, counts_row, lambdas, mf, mrf):
    """Run the stack-machine program on the CPU for one family: what the pruning kernel does, in numpy."""
    n = max(mf, mrf) + 1
    mats = {v: orc.build_matrix(n, lambdas[tree.lambda_index[v]], tree.branch[v])[:, :mf + 1] for v in range(tree.n_nodes - 1)}
    column = lambda leaf: mats[leaf][:, counts_row[tree.leaf_col[leaf]]]
    vec = None
    stack = {}
    for typ, node, flags, st, lb, n_pre, n_post in ops:
        if typ == 0:                                        # LEAVES
            assert n_pre >= 1 and n_post == 0
            vec = column(leaves[lb]).copy()
            for q in range(1, n_pre):
                vec = vec * column(leaves[lb + q])
        elif typ == 1:                                      # GEMM
            assert vec is not None
            acc = mats[node] @ vec[:mf + 1]
            vec = None                                      # consumed
            if n_pre:
                assert not (flags & 1), "leaves before the first internal child only"
                pre = column(leaves[lb]).copy()
                for q in range(1, n_pre):
                    pre = pre * column(leaves[lb + q])
                acc = pre * acc
            if flags & 1:
                assert 0 <= st < depth
                acc = stack.pop(st) * acc
            for q in range(n_post):
                acc = acc * column(leaves[lb + n_pre + q])
            if flags & 2:
                assert st not in stack and 0 <= st < depth
                stack[st] = acc
            else:
                vec = acc
        else:                                               # ROOT
            assert vec is not None and not stack
            return vec[1:mrf + 1]
    raise AssertionError("no ROOT op")


def interpret(tree, ops, n_slots, counts_row, lambdas, mf, mrf):
    """Run the schedule on the CPU for one family: what the pruning kernel does, in numpy."""
    n = max(mf, mrf) + 1
    mats = {}
    for v in range(tree.n_nodes - 1):
        mats[v] = orc.build_matrix(n, lambdas[tree.lambda_index[v]], tree.branch[v])[:, :mf + 1]
    slots = [None] * n_slots
    scratch = {}
    for typ, a, b, node in ops:
        name = OPS[int(typ)]
        if name in ("LEAF_SET", "LEAF_MUL"):
            col = mats[node][:, counts_row[tree.leaf_col[node]]]
            slots[a] = col.copy() if name == "LEAF_SET" else slots[a] * col
        elif name == "GEMM_SET":
            assert a == b
            slots[a] = mats[node] @ slots[a][:mf + 1]
        elif name == "GEMM_MUL":
            assert a != b and slots[a] is not None and slots[b] is not None
            slots[a] = slots[a] * (mats[node] @ slots[b][:mf + 1])
            slots[b] = None
        elif name == "SPILL":
            scratch[b] = slots[a]
            slots[a] = None
        elif name == "FILL":
            assert slots[a] is None
            slots[a] = scratch.pop(b)
        elif name == "ROOT":
            return slots[a][1:mrf + 1]
    raise AssertionError("no ROOT op")


TREES = {
    "cherry": "(A:1,B:1);",
    "abcd": "((A:1,B:1):1,(C:1,D:1):1);",
    "caterpillar": "((((A:1,B:2):1.5,C:3):2,D:4):1,E:7);",
    "tri": "((A:2,B:1.5,C:3):1,D:4,(E:1,F:1):2);",
    "balanced16": "((((A:1,B:1):1,(C:1,D:1):1):1,((E:1,F:1):1,(G:1,H:1):1):1):1,(((I:1,J:1):1,(K:1,L:1):1):1,((M:1,N:1):1,(O:1,P:1):1):1):1);",
}


@pytest.mark.parametrize("name", sorted(TREES))
@pytest.mark.parametrize("n_slots", [2, 3, 4, 8])
def test_schedule_reproduces_inference_prune(name, n_slots):
    tree = hostio.flatten_tree(hostio.parse_newick(TREES[name]))
    ops, n_spill = plan(tree, n_slots)
    types = [OPS[int(t)] for t in ops[:, 0]]
    assert types[-1] == "ROOT" and types.count("ROOT") == 1
    internal_children = sum(1 for v in range(tree.n_nodes - 1) if tree.leaf_col[v] < 0)
    assert types.count("GEMM_SET") + types.count("GEMM_MUL") == internal_children
    assert types.count("LEAF_SET") + types.count("LEAF_MUL") == tree.n_leaves
    assert types.count("SPILL") == types.count("FILL")
    assert all(0 <= a < n_slots for t, a, b, _ in ops if OPS[int(t)] != "SPILL" or True)
    if name == "balanced16" and n_slots == 2:
        assert n_spill > 0, "a balanced tree with two slots must spill"
    if n_slots >= 5:
        assert n_spill == 0
    rng = np.random.default_rng(3)
    mf, mrf = 14, 10
    lam = [0.04]
    for _ in range(3):
        row = rng.integers(0, 7, tree.n_leaves).astype(np.int32)
        got = interpret(tree, ops, n_slots, row, lam, mf, mrf)
        want = orc.inference_prune(tree, row, lam, mf, mrf)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", sorted(TREES))
def test_program_reproduces_inference_prune(name):
    tree = hostio.flatten_tree(hostio.parse_newick(TREES[name]))
    ops, leaves, depth = plan_program(tree)
    assert ops[-1, 0] == 2 and (ops[:, 0] == 2).sum() == 1
    internal_children = sum(1 for v in range(tree.n_nodes - 1) if tree.leaf_col[v] < 0)
    assert (ops[:, 0] == 1).sum() == internal_children
    assert len(leaves) == tree.n_leaves and sorted(leaves) == [v for v in range(tree.n_nodes) if tree.leaf_col[v] >= 0]
    assert ((ops[:, 2] & 2) != 0).sum() == ((ops[:, 2] & 1) != 0).sum(), "every parked product is popped"
    expected_depth = {"cherry": 0, "abcd": 1, "caterpillar": 0, "tri": 1, "balanced16": 3}[name]
    assert depth == expected_depth
    rng = np.random.default_rng(3)
    mf, mrf = 14, 10
    lam = [0.04]
    for _ in range(3):
        row = rng.integers(0, 7, tree.n_leaves).astype(np.int32)
        got = interpret_program(tree, ops, leaves, depth, row, lam, mf, mrf)
        want = orc.inference_prune(tree, row, lam, mf, mrf)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


def test_program_on_config5_tree():
    newick = synth.random_ultrametric_newick(100, 12345)
    tree = hostio.flatten_tree(hostio.parse_newick(newick))
    ops, leaves, depth = plan_program(tree)
    assert (ops[:, 0] == 1).sum() == 98 and len(leaves) == 100
    assert depth <= 4, "the parked stack of the benchmark tree fits tensor memory (4 entries at N = 151, three groups)"
    assert (ops[(ops[:, 0] == 1), 5] == 0).all(), "binary tree: leaves are always multiplied after the internal child"


def test_schedule_on_config5_tree_spills_at_most_once_with_four_slots():
    newick = synth.random_ultrametric_newick(100, 12345)
    tree = hostio.flatten_tree(hostio.parse_newick(newick))
    assert tree.n_nodes == 199 and tree.n_leaves == 100
    ops, n_spill = plan(tree, 4)          # 4 slots is what fits in 227 KB at N = 151
    types = [OPS[int(t)] for t in ops[:, 0]]
    assert n_spill <= 1 and types.count("SPILL") <= 1
    assert types.count("GEMM_SET") + types.count("GEMM_MUL") == 98
    assert plan(tree, 5)[1] == 0


def test_nary_children_keep_newick_order():
    tree = hostio.flatten_tree(hostio.parse_newick(TREES["tri"]))
    ops, _ = plan(tree, 8)
    root = tree.n_nodes - 1
    kids = list(tree.child_list[tree.child_offset[root]:tree.child_offset[root + 1]])
    # the ops that fold each root child into the root accumulator appear in Newick order
    fold = [int(node) for t, a, b, node in ops if int(node) in kids and OPS[int(t)] in ("LEAF_SET", "LEAF_MUL", "GEMM_SET", "GEMM_MUL")]
    assert fold == kids


def test_mammal_host_preparation(mammal):
    """Root filter, size limits and traversal order as the reference computes them (12 653 -> 10 956 families,
    mf 140, mrf 112)."""
    assert mammal["counts_all"].shape == (12653, 12)
    assert mammal["counts"].shape[0] == 10956
    assert np.array_equal(np.flatnonzero(mammal["keep"]), mammal["gold"]["keep"])
    assert (mammal["mf"], mammal["mrf"]) == (140, 112)
    assert mammal["tree"].n_nodes == 23 and mammal["tree2"].n_lambdas == 2
    names = mammal["tree2"].names
    lam_idx = dict(zip(names, mammal["tree2"].lambda_index))
    assert lam_idx["chimp"] == 1 and lam_idx["human"] == 1 and lam_idx["chimphuman"] == 1 and lam_idx["orang"] == 0


def test_host_parameter_mirror_matches_reference_and_oracle():
    """cafexp_b200/params.py (discrete gamma, root priors — what a Python host feeds the C ABI) against the compiled
    reference's values (tests/golden/scalars.json) and, bit for bit, against the oracle."""
    import json
    from cafexp_b200 import params
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scalars.json")))
    for case in gold["gamma"]:
        freq, rate = params.get_gamma(case["k"], case["alpha"])
        assert np.allclose(rate, case["rate"], rtol=1e-13, atol=0) and np.allclose(freq, case["freq"], rtol=1e-15, atol=0), case
        f2, r2 = orc.get_gamma(case["k"], case["alpha"])
        assert np.array_equal(rate, r2) and np.array_equal(freq, f2)
    for case in gold["poisson"]:
        got = params.prior_poisson(case["lambda"], case["n"], None, case["n"] + 2)
        assert np.allclose(got, case["prior"], rtol=1e-15, atol=0), case
    rd = {1: 5, 2: 3, 7: 1}
    assert np.array_equal(params.prior_uniform(30), orc.prior_uniform(30))
    assert np.array_equal(params.prior_uniform(30, rd, 40), orc.prior_uniform(30, rd, 40))
    assert np.array_equal(params.prior_poisson(2.5, 30, rd, 35), orc.prior_poisson(2.5, 30, rd, 35))


def test_header_is_plain_c(tmp_path):
    """include/cafe_b200.h is the drop-in boundary: it must compile as C99 (no C++, no torch / CUDA types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "abi.c"
    src.write_text('#include "cafe_b200.h"\nint main(void) { cafe_b200_limits l; cafe_b200_get_limits(&l); return cafe_b200_abi_version() + l.max_categories; }\n')
    res = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # and link-runs against the built library
    exe = tmp_path / "abi"
    res = subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", os.path.join(ROOT, "cafexp_b200"),
                          "-lcafe_b200", "-Wl,-rpath," + os.path.join(ROOT, "cafexp_b200")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert subprocess.run([str(exe)]).returncode > 0


def test_flat_family_reader_matches_the_python_reader(mammal, tmp_path):
    """cafe_b200_read_family_table (C, host only) against hostio.read_gene_families on the mammal table and on edge cases:
    extra species columns, shuffled column order, mixed case, CRLF line ends, blank lines."""
    flat = mammal["tree"]
    ids = [str(i) for i in np.load(os.path.join(ROOT, "tests", "golden", "mammal_counts.npz"))["ids"]]
    path = str(tmp_path / "fam.txt")
    hostio.write_gene_families(path, flat, ids, mammal["counts_all"])
    want_ids, want = hostio.read_gene_families(path, flat)
    got_ids, got = engine.read_family_table(path, flat.leaf_names)
    assert got_ids == want_ids == ids and np.array_equal(got, want) and np.array_equal(got, mammal["counts_all"])
    tree = hostio.flatten_tree(hostio.parse_newick("((A:1,b:1):1,C:2);"))
    odd = str(tmp_path / "odd.txt")
    with open(odd, "w", newline="") as fh:
        fh.write("Desc\tFamily ID\tc\tX\tB\ta\r\n(null)\tf1\t3\t99\t2\t1\r\n\r\n(null)\tf2\t0\t7\t12x\t5\r\n")
    got_ids, got = engine.read_family_table(odd, tree.leaf_names)
    by_species = [{"a": 1, "b": 2, "c": 3}, {"a": 5, "b": 12, "c": 0}]           # atoi semantics: "12x" -> 12 (src/io.cpp:186)
    want = np.array([[row[name.lower()] for name in tree.leaf_names] for row in by_species], np.int32)
    assert got_ids == ["f1", "f2"] and np.array_equal(got, want)
    bad = str(tmp_path / "bad.txt")
    open(bad, "w").write("Desc\tFamily ID\tA\tb\n(null)\tf1\t1\t2\n")
    with pytest.raises(engine.CafeB200Error, match="species missing"):
        engine.read_family_table(bad, tree.leaf_names)


def test_create_multi_rejects_bad_input():
    lib = engine.load_library()
    handle = C.c_void_p()
    tree = hostio.flatten_tree(hostio.parse_newick("(A:1,B:1);"))
    ts, _keep = engine.tree_struct(tree)
    counts = np.array([[1, 2]], np.int32)
    devs = np.array([0], np.int32)
    ip = C.POINTER(C.c_int)
    args = lambda **kw: (C.byref(handle), C.byref(ts), counts.ctypes.data_as(C.c_void_p), kw.get("bytes", 4), 1, 2, 10, 8,
                         kw.get("devs", devs).ctypes.data_as(ip) if kw.get("devs", devs) is not None else None, kw.get("n", 1))
    assert lib.cafe_b200_create_multi(*args(bytes=3)) == -1               # counts of 1, 2 or 4 bytes
    assert lib.cafe_b200_create_multi(*args(n=0)) == -1                   # at least one device
    assert lib.cafe_b200_create_multi(*args(devs=None)) == -1
    # without a GPU every well-formed create fails loudly with ERR_CUDA; with one, a duplicate device is an argument error
    rc = lib.cafe_b200_create_multi(*args(devs=np.array([0, 0], np.int32), n=2))
    assert rc == (-1 if lib.cafe_b200_device_count() > 0 else -2)


def _random_newick(rng, n_leaves, shape, max_children):
    nodes = [f"L{i}" for i in range(n_leaves)]
    rng.shuffle(nodes)
    while len(nodes) > 1:
        k = min(int(rng.integers(2, max_children + 1)), len(nodes))
        pick = [len(nodes) - 1] + list(range(k - 1)) if shape == "caterpillar" else list(rng.choice(len(nodes), size=k, replace=False))
        kids = [nodes[i] for i in pick]
        for i in sorted(pick, reverse=True):
            nodes.pop(i)
        rng.shuffle(kids)                                                  # leaves before / between / after internal children
        nodes.append("(" + ",".join(f"{c}:{float(rng.integers(1, 40))}" for c in kids) + ")")
    return nodes[0] + ";"


@pytest.mark.parametrize("seed", range(24))
def test_program_on_random_trees(seed):
    """The pruning program on random binary / n-ary / caterpillar trees, interpreted on the CPU against inference_prune:
    factor order (leaves before, between and after internal children), stack discipline, depth."""
    rng = np.random.default_rng(100 + seed)
    n_leaves = int(rng.choice([2, 3, 5, 9, 17, 30]))
    tree = hostio.flatten_tree(hostio.parse_newick(_random_newick(rng, n_leaves, str(rng.choice(["random", "caterpillar"])), int(rng.choice([2, 3, 4])))))
    ops, leaves, depth = plan_program(tree)
    assert sorted(leaves) == [v for v in range(tree.n_nodes) if tree.leaf_col[v] >= 0]
    # stack discipline: a pop finds what the matching push left, indices stay below depth, nothing is left at the root
    live = set()
    for typ, node, flags, st, lb, n_pre, n_post in ops:
        if typ != 1:
            continue
        if flags & 1:
            assert st in live
            live.discard(st)
        if flags & 2:
            assert st not in live and st < depth
            live.add(st)
    assert not live
    assert depth == (max((st + 1 for typ, _, flags, st, *_ in ops if typ == 1 and flags), default=0))
    if all(tree.child_offset[v + 1] - tree.child_offset[v] <= 2 for v in range(tree.n_nodes)):
        assert depth <= max(0, int(np.ceil(np.log2(max(n_leaves, 2)))) - 1), "binary trees: Sethi-Ullman depth"
        assert (ops[ops[:, 0] == 1, 5] == 0).all()                         # leaves always after the internal child
    mf, mrf = 12, 9
    for _ in range(2):
        row = rng.integers(0, 6, tree.n_leaves).astype(np.int32)
        got = interpret_program(tree, ops, leaves, depth, row, [0.03], mf, mrf)
        np.testing.assert_allclose(got, orc.inference_prune(tree, row, [0.03], mf, mrf), rtol=1e-12, atol=0)
