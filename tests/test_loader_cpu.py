"""The adjacency loader keeps the reference's input format (utils/data_utils.py:79-197): checked field by field
against the reference's own loader when /root/reference is present (build container), and always against its
structural invariants."""
import os
import sys

import pytest
import torch

from qagnn_b200 import data as Dt

REF = "/root/reference"


def _equal_nested(a, b):
    return len(a) == len(b) and all(len(x) == len(y) and all(torch.equal(p, q) for p, q in zip(x, y)) for x, y in zip(a, b))


@pytest.mark.parametrize("max_node_num", [200, 30])
def test_loader_matches_reference_loader(tmp_path, max_node_num):
    if not os.path.isdir(os.path.join(REF, "utils")):
        pytest.skip("reference tree not available (GPU box)")
    path = str(tmp_path / "dev.graph.adj.pk")
    Dt.synth_adj_pickle(path, 20, seed=4)
    ours = Dt.load_sparse_adj_data_with_contextnode(path, max_node_num, 5, None, use_cache=False, write_cache=False)
    sys.path.insert(0, REF)
    try:
        from oracle.ref_shim import _install_stubs
        _install_stubs()
        from utils import data_utils as RD
        ref = RD.load_sparse_adj_data_with_contextnode(path, max_node_num, 5, None)
    finally:
        sys.path.remove(REF)
    for a, b, name in zip(ours[:4], ref[:4], ("concept_ids", "node_type_ids", "node_scores", "adj_lengths")):
        assert a.dtype == b.dtype and torch.equal(a, b), name
    assert _equal_nested(ours[4][0], ref[4][0]) and _equal_nested(ours[4][1], ref[4][1])
    # the cache the reference just wrote is readable by our loader and gives the same answer
    cached = Dt.load_sparse_adj_data_with_contextnode(path, max_node_num, 5, None, use_cache=True)
    assert torch.equal(cached[0], ref[0]) and _equal_nested(cached[4][0], ref[4][0])


def test_loader_invariants_and_packing(tmp_path):
    path = str(tmp_path / "x.graph.adj.pk")
    Dt.synth_adj_pickle(path, 10, seed=1)
    n = 50
    cids, ntypes, scores, lens, (ei, et) = Dt.load_sparse_adj_data_with_contextnode(path, n, 5, None, use_cache=False,
                                                                                   write_cache=False)
    assert cids.shape == (2, 5, n) and ntypes.shape == (2, 5, n) and scores.shape == (2, 5, n, 1) and lens.shape == (2, 5)
    assert (cids[..., 0] == 0).all() and (ntypes[..., 0] == 3).all()
    for q in range(2):
        for c in range(5):
            L = int(lens[q, c]); e, t = ei[q][c], et[q][c]
            assert e.dtype == torch.int64 and e.shape[0] == 2 and e.shape[1] == t.numel() and e.shape[1] % 2 == 0
            assert int(e.max()) < n and (cids[q, c, L:] == 1).all() and (ntypes[q, c, L:] == 2).all()
            half = e.shape[1] // 2  # inverse half = swapped endpoints, relation + 19
            assert torch.equal(e[:, half:], e[:, :half].flip(0)) and torch.equal(t[half:], t[:half] + 19)
    packed = Dt.pack_adj(ei, et, n, pin=False)
    assert packed.edge_index.shape[1] == sum(x.shape[1] for r in ei for x in r)
