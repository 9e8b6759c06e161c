"""TEST / BASELINE INFRASTRUCTURE ONLY - multi-threaded CPU port of the reference's path, op for op in torch.

The reference IS torch code run on whatever device holds the model; on the GPU box's host cores this port is
what `bench.py --impl reference` and the `cpu_baseline` leg time (kind = "port": /root/reference does not
travel to the GPU box and `lavis` is not importable as a package).  It follows the reference's own op sequence
(cast, strided norm, stable sort / per-group topk loop, scatter, index_put) so the timing is the reference's
algorithm, not a tuned rewrite.  Pinned against the golden fixtures in tests/test_oracle_vs_golden.py.
Never imported by vlmc/.
"""
import torch


class WandaStat:
    """wanda_pruner.py:56-81."""

    def __init__(self, columns):
        self.scaler_row = torch.zeros(columns)
        self.nsamples = 0

    def add_batch(self, inp):
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        inp = inp.reshape(-1, inp.shape[-1]).t()
        self.scaler_row *= self.nsamples / (self.nsamples + b)
        self.nsamples += b
        inp = inp.type(torch.float32)
        self.scaler_row += torch.norm(inp, p=2, dim=1) ** 2 / self.nsamples


def wanda_select(W, scaler_row, sparsity, prune_n=0, prune_m=0, whole_matrix=False):
    """wanda_pruner.py:318-341 / :664-687.  Returns (keep mask, importance score); W is zeroed in place."""
    metric = torch.abs(W) * torch.sqrt(scaler_row.reshape(1, -1))
    score = metric.abs().mean().item()
    prune = torch.zeros_like(metric) == 1
    if prune_n != 0:
        for ii in range(0, metric.shape[1], prune_m):
            grp = metric[:, ii:ii + prune_m].float()
            prune.scatter_(1, ii + torch.topk(grp, prune_n, dim=1, largest=False)[1], True)
    elif whole_matrix:
        thres = torch.sort(metric.flatten())[0][int(metric.numel() * sparsity)]
        prune = metric < thres
    else:
        order = torch.sort(metric, dim=-1, stable=True)[1]
        prune.scatter_(1, order[:, :int(metric.shape[1] * sparsity)], True)
    W[prune] = 0
    return ~prune, score


def lora_merge(W, A, B, scaling, mask):
    """lora.py:384-387 + train.py:634-637."""
    W += ((B @ A) * scaling) * mask
    W[~mask] = 0
    return W
