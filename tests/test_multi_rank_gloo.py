"""CPU, world_size 2 over gloo: the sharding / collective logic of vlmc.parallel with the numpy oracle
standing in for the CUDA kernels.  What is checked is the protocol: token shards + ONE sum all-reduce reproduce the
single-rank statistics, row shards + all-gather reproduce the single-rank masks / weights, the SparseGPT block
threshold is a k-th value over ALL shards' rows, DSnoT's executed-cycle count is a MAX over shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, fn_name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        globals()[fn_name](rank, out)
    finally:
        dist.destroy_process_group()


def _run(fn_name):
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(_free_port(), fn_name, out), nprocs=WORLD, join=True)
    assert all(out.get(r) == "ok" for r in range(WORLD)), dict(out)


def _data(seed, n_seq=6, S=24, C=128, R=40):
    g = torch.Generator().manual_seed(seed)
    xs = [(torch.randn(S, C, generator=g) * (1 + 0.1 * j)).numpy().astype(np.float32) for j in range(n_seq)]
    W = (torch.randn(R, C, generator=g) * 0.05).numpy().astype(np.float32)
    return xs, W


def _wanda(rank, out):
    from oracle import oracle
    from vlmc import parallel
    xs, W = _data(1)
    N, C = len(xs), W.shape[1]
    # single-rank truth: the reference's running mean over all sequences
    s_ref, n = np.zeros(C, np.float32), 0
    for x in xs:
        s_ref, n = oracle.wanda_add_batch(s_ref, n, x, 1)
    keep_ref, W_ref, mean_ref = oracle.wanda_nm(W, s_ref, 2, 4)

    def accum(x_local, state, n_before, b):          # stands in for native.sqnorm_accum(x, s, 0, N)
        s, _ = oracle.wanda_add_batch(state.numpy(), n_before, np.concatenate(x_local), b)
        state.copy_(torch.from_numpy(s))
    a, b = parallel.sample_range(N, rank, WORLD)
    s = parallel.sharded_accumulate(accum, xs[a:b], torch.zeros(C), N)
    assert np.abs(s.numpy() - s_ref).max() <= 1e-5 * np.abs(s_ref).max()

    def select(Wr, scal, keep_rows):                 # stands in for native.wanda_nm on the row shard
        k, Wp, m = oracle.wanda_nm(Wr.numpy(), scal.numpy(), 2, 4)
        Wr.copy_(torch.from_numpy(Wp))
        keep_rows.copy_(torch.from_numpy(k))
        return torch.tensor([m], dtype=torch.float32)
    Wt = torch.from_numpy(W.copy())
    keep, mean = parallel.prune_linear_row_sharded(Wt, torch.from_numpy(s_ref), select, rank, WORLD)
    assert np.array_equal(keep.numpy(), keep_ref) and np.array_equal(Wt.numpy(), W_ref)
    assert abs(mean.item() - mean_ref) < 1e-5 * mean_ref
    out[rank] = "ok"


def _wanda_packed(rank, out):
    """Row shards + ONE all-gather of bit-packed masks for several linears: same masks / weights as one rank."""
    from oracle import oracle
    from vlmc import parallel
    g = torch.Generator().manual_seed(7)
    shapes = [(8, 64), (12, 32), (4, 128)]
    Ws = [(torch.randn(R, C, generator=g) * 0.05).numpy().astype(np.float32) for R, C in shapes]
    scal = [(torch.rand(C, generator=g) * 9 + 1).numpy() for _, C in shapes]
    refs = [oracle.wanda_nm(W, s, 2, 4) for W, s in zip(Ws, scal)]

    def make_select(i):
        def select(Wr, keep_rows):                   # stands in for native.wanda_nm on the row shard
            k, Wp, m = oracle.wanda_nm(Wr.numpy(), scal[i], 2, 4)
            Wr.copy_(torch.from_numpy(Wp))
            keep_rows.copy_(torch.from_numpy(k))
            return torch.tensor([m], dtype=torch.float32)
        return select

    def pack(keep_rows, bits):                       # stands in for native.mask_pack
        bits.copy_(torch.from_numpy(np.packbits(keep_rows.numpy().astype(np.uint8), axis=1, bitorder="little")))

    def apply(Wf, bits, keep, rows_per_seg, seg_stride):     # stands in for native.mask_apply_packed
        R, C = Wf.shape
        b = bits.numpy()
        rows = [b[g * seg_stride: g * seg_stride + rows_per_seg * (C // 8)].reshape(rows_per_seg, C // 8)
                for g in range(R // rows_per_seg)]
        k = np.unpackbits(np.concatenate(rows), axis=1, bitorder="little").astype(bool)
        keep.copy_(torch.from_numpy(k))
        Wf.mul_(torch.from_numpy(k.astype(np.float32)))
    Wt = [torch.from_numpy(W.copy()) for W in Ws]
    res = parallel.prune_block_rows_packed(Wt, [make_select(i) for i in range(3)], pack, apply, rank, WORLD)
    for (keep, mean), W, (keep_ref, W_ref, mean_ref) in zip(res, Wt, refs):
        assert np.array_equal(keep.numpy(), keep_ref) and np.array_equal(W.numpy(), W_ref)
        assert abs(mean.item() - mean_ref) < 1e-5 * mean_ref

    # the batched calls (vlmc_wanda_rowselect_batch / vlmc_mask_pack_batch / vlmc_mask_apply_packed_batch on the GPU): one
    # select, one pack and one expand call for ALL linears around the same single all-gather - same results
    calls = {"select": 0, "pack": 0, "apply": 0}

    def select_batch(W_rows, keep_rows):
        calls["select"] += 1
        return torch.cat([make_select(i)(w, k) for i, (w, k) in enumerate(zip(W_rows, keep_rows))])

    def pack_batch(keep_rows_list, bits_list):
        calls["pack"] += 1
        for k, b in zip(keep_rows_list, bits_list):
            pack(k, b)

    def apply_batch(Wfs, bits_views, keeps, rows_per_seg, seg_stride):
        calls["apply"] += 1
        for Wf, b, k, rps in zip(Wfs, bits_views, keeps, rows_per_seg):
            apply(Wf, b, k, rps, seg_stride)
    Wt = [torch.from_numpy(W.copy()) for W in Ws]
    res = parallel.prune_block_rows_packed(Wt, [None] * 3, None, None, rank, WORLD, select_batch_fn=select_batch,
                                           pack_batch_fn=pack_batch, apply_batch_fn=apply_batch)
    assert calls == {"select": 1, "pack": 1, "apply": 1}
    for (keep, mean), W, (keep_ref, W_ref, mean_ref) in zip(res, Wt, refs):
        assert np.array_equal(keep.numpy(), keep_ref) and np.array_equal(W.numpy(), W_ref)
        assert abs(mean.item() - mean_ref) < 1e-5 * mean_ref
    out[rank] = "ok"


def _sparsegpt(rank, out):
    from oracle import oracle
    from vlmc import parallel
    xs, W = _data(2, n_seq=8, S=64, C=256, R=48)
    N, (R, C) = len(xs), W.shape
    H_ref, n = np.zeros((C, C), np.float32), 0
    for x in xs:
        H_ref, n = oracle.sparsegpt_add_batch(H_ref, n, x, 1)
    U_ref, dead_ref, _ = oracle.sparsegpt_inverse_factor(H_ref)
    W_ref, _, _ = oracle.sparsegpt_fasterprune(W, "f32", None, 0.5, U=U_ref, dead=dead_ref)

    def accum(x_local, state, n_before, b):          # raw partial sum with the global divisor 2/N
        X = np.concatenate(x_local).astype(np.float64)
        state.copy_(torch.from_numpy((X.T @ X * (2.0 / b)).astype(np.float32)))
    a, b = parallel.sample_range(N, rank, WORLD)
    H = parallel.sharded_accumulate(accum, xs[a:b], torch.zeros(C, C), N)
    assert np.abs(H.numpy() - H_ref).max() <= 1e-5 * np.abs(H_ref).max()

    # two Hessians -> one factorisation per rank, factors broadcast
    calls = []

    def factor(Hm):
        calls.append(1)
        U, dead, _ = oracle.sparsegpt_inverse_factor(Hm.numpy())
        return torch.from_numpy(U), torch.from_numpy(dead.astype(np.uint8))
    facs = parallel.factor_all([torch.from_numpy(H_ref), torch.from_numpy(H_ref * 2)], factor,
                               lambda Hm: (torch.empty(C, C), torch.empty(C, dtype=torch.uint8)), rank, WORLD)
    assert len(calls) == 1
    assert np.array_equal(facs[0][0].numpy(), U_ref)
    assert np.allclose(facs[1][0].numpy(), U_ref / np.sqrt(2), rtol=1e-4, atol=1e-6)

    # row-sharded sweep: the block threshold must be the k-th value over the rows of BOTH shards
    def sweep(Wr, U, dead, rows_total, reduce_sum):
        def kth_over_all_shards(tmp, k_unused):
            k = int(rows_total * tmp.shape[1] * 0.5)
            parts = [None] * WORLD
            dist.all_gather_object(parts, tmp.ravel())
            allv = np.concatenate(parts)
            return np.partition(allv, k)[k]
        Wp, _, _ = oracle.sparsegpt_fasterprune(Wr.numpy(), "f32", None, 0.5, U=U.numpy(), dead=dead.numpy().astype(bool),
                                                kth_fn=kth_over_all_shards)
        Wr.copy_(torch.from_numpy(Wp))
    Wt = torch.from_numpy(W.copy())
    parallel.obs_rows_sharded(Wt, torch.from_numpy(U_ref), torch.from_numpy(dead_ref.astype(np.uint8)), sweep, rank, WORLD)
    assert np.array_equal(Wt.numpy(), W_ref)
    out[rank] = "ok"


def _sparsegpt_linears(rank, out):
    """Whole linears per rank (vlmc.parallel.prune_linears_task_parallel): every linear is pruned by exactly one rank,
    by the unsharded algorithm, and every rank ends up with every pruned weight."""
    from oracle import oracle
    from vlmc import parallel
    g = torch.Generator().manual_seed(11)
    shapes = [(24, 128), (40, 128), (16, 256)]
    Ws, Hs = [], []
    for R, C in shapes:
        X = torch.randn(4 * C, C, generator=g).numpy().astype(np.float64)
        Hs.append((X.T @ X * (2.0 / 4)).astype(np.float32))
        Ws.append((torch.randn(R, C, generator=g) * 0.05).numpy().astype(np.float32))
    refs = []
    for W, H in zip(Ws, Hs):
        U, dead, _ = oracle.sparsegpt_inverse_factor(H.copy())
        refs.append(oracle.sparsegpt_fasterprune(W, "f32", None, 0.5, U=U, dead=dead)[0])
    Wt = [torch.from_numpy(W.copy()) for W in Ws]
    ran = []

    def run(indices):
        for i in indices:
            ran.append(i)
            U, dead, _ = oracle.sparsegpt_inverse_factor(Hs[i].copy())
            Wt[i].copy_(torch.from_numpy(oracle.sparsegpt_fasterprune(Ws[i], "f32", None, 0.5, U=U, dead=dead)[0]))
    owners = parallel.prune_linears_task_parallel(Wt, run, rank, WORLD)
    assert sorted(ran) == [i for i, o in enumerate(owners) if o == rank] and len(set(owners)) == WORLD
    for W, ref in zip(Wt, refs):
        assert np.array_equal(W.numpy(), ref)
    out[rank] = "ok"


def _dsnot(rank, out):
    from oracle import oracle
    from vlmc import parallel
    g = torch.Generator().manual_seed(3)
    R, C = 24, 320
    W = (torch.randn(R, C, generator=g) * 0.05).numpy().astype(np.float32)
    scal = (torch.rand(C, generator=g) * 50 + 1).numpy()
    summ = (torch.randn(C, generator=g) * 20).numpy()
    var = (torch.rand(C, generator=g) + 0.2).numpy()
    k = round(C * 0.6)
    keep_ref, cyc_ref = oracle.dsnot_refine(W, scal, summ, var, sparsity_num=k, ref_fixup=False)
    s, e = parallel.row_range(R, rank, WORLD)
    # a shard on its own may stop earlier than the whole matrix: the executed-cycle count is a MAX over shards
    _, cyc_local = oracle.dsnot_refine(W[s:e], scal, summ, var, sparsity_num=k, ref_fixup=False)
    t = parallel.allreduce_max(torch.tensor([cyc_local]))
    assert int(t.item()) == cyc_ref
    keep_local, _ = oracle.dsnot_refine(W[s:e], scal, summ, var, sparsity_num=k, ref_fixup=False,
                                        force_cycles=int(t.item()))
    full = torch.zeros(R, C, dtype=torch.uint8)
    full[s:e] = torch.from_numpy(keep_local.astype(np.uint8))
    parallel.gather_rows(full, rank, WORLD)
    assert np.array_equal(full.numpy().astype(bool), keep_ref)
    # running means merge: per-rank means weighted by the sample counts
    means = [torch.full((4,), float(rank + 1))]
    parallel.merge_running_means(means, n_local=rank + 1, n_total=3)
    assert torch.allclose(means[0], torch.full((4,), (1 * 1 + 2 * 2) / 3.0))
    out[rank] = "ok"


@pytest.mark.parametrize("fn", ["_wanda", "_wanda_packed", "_sparsegpt", "_sparsegpt_linears", "_dsnot"])
def test_two_ranks(fn, built_lib):
    _run(fn)


def _importance_dp(rank, out):
    """SURVEY 8f-4 across ranks: round-robin batches + ONE all-reduce of the packed accumulator reproduce the sequential
    loop of compute_importance_scores (layer_single_base_pruner.py:440-463), including its stop rule on ragged batch
    lengths (no batch is started once accum_samples >= num_samples) and the batch count used as the divisor."""
    from oracle import oracle
    from vlmc import parallel
    g = torch.Generator().manual_seed(21)
    shapes = [(6, 10), (3, 5)]
    lens = [4, 4, 2, 4, 1, 4, 4, 3, 4]
    batches = [{"len": n, "grads": [torch.randn(s, generator=g) * 0.1 for s in shapes]} for n in lens]
    w = [torch.randn(s, generator=g) for s in shapes]
    for num_samples in (1, 8, 10, 11, 14, 100):
        # sequential truth (the reference's loop)
        used, accum = [], 0
        for b in batches:
            if accum >= num_samples:
                break
            accum += b["len"]
            used.append(b)
        sizes = [int(np.prod(s)) for s in shapes]
        packed = torch.zeros(sum(sizes))
        views = [packed[sum(sizes[:i]):sum(sizes[:i + 1])].view(s) for i, s in enumerate(shapes)]

        def accum_fn(grads):                         # stands in for native.importance_accum(acc, grads, "obd")
            for v, gr in zip(views, grads):
                v += gr * gr
        nb = parallel.importance_accumulate_data_parallel(
            batches, num_samples, lambda d: (d["grads"], d["len"]), accum_fn, packed, rank, WORLD)
        assert nb == len(used), (num_samples, nb, len(used))
        for i, s in enumerate(shapes):
            want = oracle.importance_scores_first_order(w[i].numpy(), [b["grads"][i].numpy() for b in used], "obd")
            got = (w[i] * w[i]) * (views[i] / nb)
            assert np.abs(got.numpy() - want).max() <= 1e-6 * np.abs(want).max(), (num_samples, i)
    out[rank] = "ok"


def test_importance_scores_data_parallel():
    _run("_importance_dp")


def _driver_merge(rank, out):
    """layerwise.merge_statistics_across_ranks (data-parallel calibration in the drop-in driver): per-rank running means
    over strided sample shares merge into the single-rank statistics - Wanda's scaler_row, DSnoT's four vectors (mean /
    var token-weighted, var = mean of per-call variances) and a SparseGPT Hessian accumulated under the global divisor."""
    import types
    from oracle import oracle
    from vlmc.compression.pruners import layerwise
    xs, _ = _data(5, n_seq=7)
    N, C = len(xs), xs[0].shape[1]
    st_ref = dict(scaler_row=np.zeros(C, np.float32), sum_metric_row=np.zeros(C, np.float32), mean=np.zeros(C, np.float32),
                  var=np.zeros(C, np.float32), nsamples=0, ntokens=0)
    H_ref, n = np.zeros((C, C), np.float32), 0
    for x in xs:
        st_ref = oracle.dsnot_add_batch(st_ref, x, 1)
        H_ref, n = oracle.sparsegpt_add_batch(H_ref, n, x, 1)
    mine = xs[rank::WORLD]
    st = dict(scaler_row=np.zeros(C, np.float32), sum_metric_row=np.zeros(C, np.float32), mean=np.zeros(C, np.float32),
              var=np.zeros(C, np.float32), nsamples=0, ntokens=0)
    H = np.zeros((C, C), np.float64)
    for x in mine:
        st = oracle.dsnot_add_batch(st, x, 1)
        H += (2.0 / N) * x.astype(np.float64).T @ x.astype(np.float64)         # what SparseGPT.add_batch does under _global_n
    wd = types.SimpleNamespace(scaler_row=torch.from_numpy(st["scaler_row"].copy()), sum_metric_row=torch.from_numpy(st["sum_metric_row"].copy()),
                               mean=torch.from_numpy(st["mean"].copy()).reshape(-1, 1), var=torch.from_numpy(st["var"].copy()).reshape(-1, 1),
                               nsamples=st["nsamples"], ntokens=st["ntokens"])
    ww = types.SimpleNamespace(scaler_row=torch.from_numpy(st["scaler_row"].copy()), nsamples=st["nsamples"])
    ws = types.SimpleNamespace(H=torch.from_numpy(H.astype(np.float32)), nsamples=st["nsamples"])
    follower = types.SimpleNamespace(H=ws.H, nsamples=st["nsamples"])          # shares the leader's Hessian: reduced once
    layerwise.merge_statistics_across_ranks([wd, ww, ws, follower], N)
    for k in ("scaler_row", "sum_metric_row", "mean", "var"):
        got = getattr(wd, k).reshape(-1).numpy()
        assert np.abs(got - st_ref[k]).max() <= 1e-5 * np.abs(st_ref[k]).max(), k
    assert np.abs(ww.scaler_row.numpy() - st_ref["scaler_row"]).max() <= 1e-5 * np.abs(st_ref["scaler_row"]).max()
    assert np.abs(ws.H.numpy() - H_ref).max() <= 1e-5 * np.abs(H_ref).max() and follower.H is ws.H
    assert wd.nsamples == ww.nsamples == ws.nsamples == follower.nsamples == N and wd.ntokens == st_ref["ntokens"]
    out[rank] = "ok"


def test_driver_statistics_merge_across_ranks():
    _run("_driver_merge")
