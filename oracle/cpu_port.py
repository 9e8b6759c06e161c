"""TEST / BASELINE INFRASTRUCTURE ONLY - multi-threaded CPU port of the reference's path, op for op in torch.

FALLBACK of `bench.py --impl reference` and the `cpu_baseline` leg (kind = "port"), used only when the unmodified
reference files are absent (neither /root/reference nor the verbatim copy under baseline/_ref that oracle/fetch_ref.py
makes; with them the reference's own composite pruner is timed, oracle/ref_arm.py, kind = "reference").  It follows the
reference's own op sequence (cast, strided norm, stable sort / per-group topk loop, scatter, index_put) so the timing
is the reference's algorithm, not a tuned rewrite.  Pinned against the golden fixtures the reference produced in
tests/test_reference_arm.py.  Never imported by vlmc/.
"""
import torch


class WandaStat:
    """wanda_pruner.py:56-81."""

    def __init__(self, columns, device=None):
        self.scaler_row = torch.zeros(columns, device=device)    # device: only bench.py's "same GPU, torch eager" figure
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


class SparseGPTPort:
    """sparsegpt_pruner.py:55-215: Hessian accumulation + fasterprune, the reference's op sequence on CPU tensors."""

    def __init__(self, weight):
        self.W = weight                      # [R, C] in its own dtype, pruned in place
        self.rows, self.columns = weight.shape
        self.H = torch.zeros((self.columns, self.columns))
        self.nsamples = 0

    def add_batch(self, inp):                                       # :68-79
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        inp = inp.reshape(-1, inp.shape[-1]).t()
        self.H *= self.nsamples / (self.nsamples + b)
        self.nsamples += b
        inp = (2 / self.nsamples) ** 0.5 * inp.float()
        self.H += inp.matmul(inp.t())

    @staticmethod
    def _chol(H, damp, upper):                                      # :114-128 / :146-157: damp only after a failure
        while True:
            L, info = torch.linalg.cholesky_ex(H, upper=upper)
            if int(info) == 0 and not torch.isnan(L).any():
                return L
            H[torch.arange(H.shape[0]), torch.arange(H.shape[0])] += damp

    def fasterprune(self, sparsity, prune_n=0, prune_m=0, blocksize=128, percdamp=.01):
        W = self.W.clone().float()
        H = self.H
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        W[:, dead] = 0
        L = self._chol(H, percdamp * torch.mean(torch.diag(H)), False)
        H = torch.cholesky_inverse(L)
        Hinv = self._chol(H, percdamp * torch.mean(torch.abs(torch.diag(H))), True)
        score = torch.mean(W ** 2 / (torch.diag(Hinv).reshape(1, -1)) ** 2).item()
        for i1 in range(0, self.columns, blocksize):                # :169-210
            i2 = min(i1 + blocksize, self.columns)
            count = i2 - i1
            W1 = W[:, i1:i2].clone()
            Q1 = torch.zeros_like(W1)
            Err1 = torch.zeros_like(W1)
            Hinv1 = Hinv[i1:i2, i1:i2]
            if prune_n == 0:
                tmp = W1 ** 2 / (torch.diag(Hinv1).reshape(1, -1)) ** 2
                thresh = torch.sort(tmp.flatten())[0][int(tmp.numel() * sparsity)]
                mask1 = tmp <= thresh
            else:
                mask1 = torch.zeros_like(W1) == 1
            for i in range(count):
                w = W1[:, i]
                d = Hinv1[i, i]
                if prune_n != 0 and i % prune_m == 0:
                    tmp = W1[:, i:i + prune_m] ** 2 / (torch.diag(Hinv1)[i:i + prune_m].reshape(1, -1)) ** 2
                    mask1.scatter_(1, i + torch.topk(tmp, prune_n, dim=1, largest=False)[1], True)
                q = w.clone()
                q[mask1[:, i]] = 0
                Q1[:, i] = q
                err1 = (w - q) / d
                W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0))
                Err1[:, i] = err1
            W[:, i1:i2] = Q1
            W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])
        self.W.copy_(W.to(self.W.dtype))
        return score


class DSnoTStat:
    """dsnot_pruner.py:58-101."""

    def __init__(self, columns):
        self.scaler_row = torch.zeros(columns)
        self.sum_metric_row = torch.zeros(columns)
        self.mean = torch.zeros(columns)
        self.var = torch.zeros(columns)
        self.nsamples = 0
        self.ntokens = 0

    def add_batch(self, inp):
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        inp = inp.reshape(-1, inp.shape[-1]).t().type(torch.float32)
        mean_inp = torch.mean(inp, dim=1, keepdim=True)
        var_inp = torch.var(inp, dim=1, unbiased=False, keepdim=True)
        n = inp.shape[1]
        self.var = var_inp if self.ntokens == 0 else (self.var * self.ntokens + var_inp * n) / (self.ntokens + n)
        self.mean = mean_inp if self.ntokens == 0 else (self.mean * self.ntokens + mean_inp * n) / (self.ntokens + n)
        self.ntokens += n
        self.scaler_row *= self.nsamples / (self.nsamples + b)
        self.sum_metric_row *= self.nsamples / (self.nsamples + b)
        self.nsamples += b
        self.scaler_row += torch.norm(inp, p=2, dim=1) ** 2 / self.nsamples
        self.sum_metric_row += torch.sum(inp, dim=1) / self.nsamples


def global_get_mask(importance_scores, p, max_sparsity_per_layer):
    """LayerSparsity.get_mask (layer_single_base_pruner.py:149-176) with the reference's own op sequence on CPU tensors:
    per-tensor topk + where + index_put for the protected entries, cat of every score, topk, one compare per tensor."""
    for k, v in importance_scores.items():
        num_to_set = int(v.numel() * (1 - max_sparsity_per_layer))
        if num_to_set > 0:
            threshold = torch.topk(v.flatten(), num_to_set, largest=True)[0][-1]
            v[torch.where(v >= threshold)] = torch.finfo(v.dtype).max
    all_scores = torch.cat([t.flatten() for t in importance_scores.values()])
    threshold = torch.topk(all_scores, int(p * all_scores.numel()), largest=False)[0][-1]
    return {k: (v > threshold).type(v.dtype) for k, v in importance_scores.items()}
