import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: CPU test taking more than ~20 s")


def load_json(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def fnum(v):
    """ref_harness prints non-finite doubles as strings."""
    return float(v) if isinstance(v, str) else v


@pytest.fixture(scope="session")
def mammal():
    """Inputs of BASELINE.json configs 1-4 (tests/golden/mammal_*), already root-filtered like the reference CLI."""
    from cafexp_b200 import hostio
    inp = load_json("mammal_inputs.json")
    flat = hostio.flatten_tree(hostio.parse_newick(inp["tree"]))
    assert flat.leaf_names == inp["leaf_names"]
    flat2 = hostio.flatten_tree(hostio.parse_newick(inp["tree"]), hostio.parse_newick(inp["lambda_tree"], True))
    z = np.load(os.path.join(GOLD, "mammal_counts.npz"))
    counts = z["counts"].astype(np.int32)
    keep = hostio.exists_at_root(flat, counts)
    out = np.load(os.path.join(GOLD, "mammal_outputs.npz"))
    meta = json.loads(str(out["meta"]))
    mf, mrf = hostio.family_size_limits(counts)
    return {"inputs": inp, "tree": flat, "tree2": flat2, "counts_all": counts, "counts": counts[keep], "keep": keep,
            "gold": out, "meta": meta, "mf": mf, "mrf": mrf}
