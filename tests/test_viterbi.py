"""Branch probabilities of reconstructed size changes (compute_viterbi_sum, src/gene_family_reconstructor.cpp:361-400):
SURVEY section 8(f) rank 3.  Golden: tests/golden/mammal_viterbi.npz (scripts/make_golden.py --only viterbi)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT
from cafexp_b200 import hostio
from oracle import binding as orc


@pytest.fixture(scope="module")
def vit(mammal):
    z = np.load(os.path.join(GOLD, "mammal_viterbi.npz"))
    meta = json.loads(str(z["meta"]))
    tree = mammal["tree"]
    assert list(tree.names) == meta["node_order"]          # the flattened tree visits nodes in the reference's order
    return {"meta": meta, "tree": tree, "sizes": z["node_sizes"].astype(np.int32), "probs": z["probabilities"],
            "families": np.ascontiguousarray(mammal["counts"][: meta["limit"]])}


def test_oracle_branch_probabilities_match_reference(vit):
    m = vit["meta"]
    got = orc.branch_probabilities(vit["tree"], vit["sizes"], [m["lambda"]], m["max_family_size"], m["max_root_family_size"])
    assert np.array_equal(got < 0, vit["probs"] < 0)
    ok = vit["probs"] >= 0
    assert np.allclose(got[ok], vit["probs"][ok], rtol=1e-13, atol=0)
    # leaves carry the observed counts; the root and unchanged nodes have no value
    leaf = vit["tree"].leaf_col >= 0
    assert np.array_equal(vit["sizes"][:, leaf], vit["families"][:, vit["tree"].leaf_col[leaf]])
    assert (vit["probs"][:, -1] < 0).all()
    sel = np.zeros(len(vit["sizes"]), np.uint8)
    sel[::3] = 1
    part = orc.branch_probabilities(vit["tree"], vit["sizes"], [m["lambda"]], m["max_family_size"], m["max_root_family_size"], selected=sel)
    assert (part[sel == 0] < 0).all() and np.array_equal(part[sel == 1], got[sel == 1])


@pytest.mark.gpu
def test_gpu_branch_probabilities(vit):
    from cafexp_b200 import engine
    m = vit["meta"]
    with engine.Engine(vit["tree"], vit["families"], m["max_family_size"], m["max_root_family_size"]) as eng:
        got = eng.branch_probabilities([m["lambda"]], vit["sizes"])
        sel = np.zeros(len(vit["sizes"]), np.uint8)
        sel[1::2] = 1
        part = eng.branch_probabilities([m["lambda"]], vit["sizes"], selected=sel)
        bad = vit["sizes"].copy()
        bad[0, 0] = m["max_family_size"] + 1
        with pytest.raises(engine.CafeB200Error, match="COUNT_RANGE"):
            eng.branch_probabilities([m["lambda"]], bad)
    assert np.array_equal(got < 0, vit["probs"] < 0)
    ok = vit["probs"] >= 0
    # matrix entries differ from the host's by CUDA's exp() (< 1 ulp); the selection p < p* only flips at exact ties
    assert np.allclose(got[ok], vit["probs"][ok], rtol=1e-12, atol=0)
    assert (part[sel == 0] < 0).all() and np.array_equal(part[sel == 1], got[sel == 1])


@pytest.mark.gpu
def test_reference_host_code_with_cuda_branch_probabilities(vit, mammal, tmp_path):
    """ref_harness_cuda eval --recon --viterbi --cuda 1: reconstruction and branch probabilities from the device, driven by
    the reference's host code (compute_branch_probabilities_cuda, integration/cuda_models.cpp)."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness_cuda")
    if not os.path.exists(harness):
        pytest.fail("oracle/_ref/ref_harness_cuda is missing: run `make -C oracle refcuda` in the build container")
    m = vit["meta"]
    tree_path, fam_path = tmp_path / "tree.txt", tmp_path / "fam.txt"
    tree_path.write_text(mammal["inputs"]["tree"] + "\n")
    ids = np.load(os.path.join(GOLD, "mammal_counts.npz"))["ids"]
    hostio.write_gene_families(str(fam_path), mammal["tree"], [str(i) for i in ids], mammal["counts_all"])
    dump = tmp_path / "vit.bin"
    res = subprocess.run([harness, "eval", "--cuda", "1", "--tree", str(tree_path), "--fam", str(fam_path), "--lambda", repr(m["lambda"]),
                          "--limit", str(m["limit"]), "--recon", "--viterbi", "--dumprecon", str(tmp_path / "r.rec"), "--dumpviterbi", str(dump)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    raw = dump.read_bytes()
    F, nn = vit["sizes"].shape
    sizes = np.frombuffer(raw, np.int32, F * nn, 0).reshape(F, nn)
    probs = np.frombuffer(raw, np.float64, F * nn, sizes.nbytes).reshape(F, nn)
    assert np.array_equal(sizes, vit["sizes"])                      # reconstructed ancestral counts: bit-exact
    assert np.array_equal(probs < 0, vit["probs"] < 0)
    ok = vit["probs"] >= 0
    assert np.allclose(probs[ok], vit["probs"][ok], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_gpu_branch_probabilities_sharded_over_two_devices(vit):
    """cafe_b200_branch_probabilities and cafe_b200_reconstruct through a two-device context equal the one-device results."""
    from cafexp_b200 import engine
    if engine.device_count() < 2:
        pytest.skip("needs two visible CUDA devices")
    m = vit["meta"]
    sel = np.zeros(len(vit["sizes"]), np.uint8)
    sel[::3] = 1
    out = []
    for dev in (0, [0, 1]):
        with engine.Engine(vit["tree"], vit["families"], m["max_family_size"], m["max_root_family_size"], device=dev) as eng:
            prior = orc.prior_uniform(m["max_root_family_size"], None, min(m["max_family_size"], m["max_root_family_size"]) + 1)
            out.append((eng.branch_probabilities([m["lambda"]], vit["sizes"]), eng.branch_probabilities([m["lambda"]], vit["sizes"], selected=sel),
                        eng.reconstruct([[m["lambda"]]], prior)))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
