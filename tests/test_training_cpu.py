"""Host logic of the training path (qagnn_b200/training.py) that needs no GPU: the combo one-hot table and the
combo-histogram BatchNorm of the shared edge encoder against a naive per-edge evaluation (what the reference does,
modeling_qagnn.py:419-433 with the module in .train())."""
import torch

from oracle import qagnn_oracle as O
from qagnn_b200 import training as TR


def _edge_features(inp, T, R):
    """[E+N, R+1+2T] one-hot edge features exactly as GATConvE.forward builds them (:419-432)."""
    ei, et, nt = inp["edge_index"], inp["edge_type"], inp["node_type"].view(-1)
    N = nt.numel()
    ev = torch.nn.functional.one_hot(et, R + 1).float()
    sv = torch.zeros(N, R + 1); sv[:, R] = 1
    hv = torch.nn.functional.one_hot(nt[ei[0]], T).float()
    tv = torch.nn.functional.one_hot(nt[ei[1]], T).float()
    self_ht = torch.nn.functional.one_hot(nt, T).float()
    return torch.cat([torch.cat([ev, sv]), torch.cat([torch.cat([hv, tv], 1), torch.cat([self_ht, self_ht], 1)])], 1)


def test_combo_table_rows_are_the_per_edge_features():
    T, R = 4, 38
    inp = O.synth_graph_batch(3, 20, 60, 64, R, seed=3, realistic=True)
    prep = O.graph_prep_oracle(inp["edge_index"], inp["edge_type"], inp["node_type"].view(-1), T, R)
    tab = TR.combo_onehot_table(T, R, torch.device("cpu"))
    assert tab.shape == (R * T * T + T, R + 1 + 2 * T)
    assert torch.equal(tab[torch.as_tensor(prep["combo"])], _edge_features(inp, T, R))


def test_histogram_batchnorm_equals_per_edge_batchnorm_and_running_stats():
    T, R, D, k = 4, 38, 32, 3
    torch.manual_seed(0)
    inp = O.synth_graph_batch(4, 16, 50, D, R, seed=5)
    prep = O.graph_prep_oracle(inp["edge_index"], inp["edge_type"], inp["node_type"].view(-1), T, R)
    combo = torch.as_tensor(prep["combo"])

    def make():
        torch.manual_seed(1)
        enc = torch.nn.Sequential(torch.nn.Linear(R + 1 + 2 * T, D), torch.nn.BatchNorm1d(D), torch.nn.ReLU(), torch.nn.Linear(D, D))
        enc[1].running_mean.normal_(); enc[1].running_var.uniform_(0.5, 2.0)
        return enc.train()
    ref, mine = make(), make()
    feats = _edge_features(inp, T, R)
    for _ in range(k):  # the reference calls the shared module once per layer, on one row per edge
        e_ref = ref(feats)
    g = torch.randn_like(e_ref)
    (e_ref * g).sum().backward()
    tab = TR.edge_table_train(mine, TR.combo_onehot_table(T, R, torch.device("cpu")),
                              torch.bincount(combo, minlength=R * T * T + T), k)
    e_mine = tab[combo]
    assert torch.allclose(e_mine, e_ref, atol=1e-5, rtol=1e-5)
    (e_mine * g * k).sum().backward()  # the reference back-propagates through each of its k identical calls
    # (only the last call's graph reached the loss above, so compare against a single call's gradient)
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mine.named_parameters()):
        assert torch.allclose(p2.grad / k, p1.grad, atol=2e-4, rtol=1e-4), n1
    assert torch.allclose(mine[1].running_mean, ref[1].running_mean, atol=1e-6)
    assert torch.allclose(mine[1].running_var, ref[1].running_var, atol=1e-5, rtol=1e-5)
    assert int(mine[1].num_batches_tracked) == int(ref[1].num_batches_tracked) == k
