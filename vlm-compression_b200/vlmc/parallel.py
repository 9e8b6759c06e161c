"""Multi-GPU plan for the path (new: the reference runs independent replicas).

  phase 1  calibration tokens are split across ranks; every rank accumulates its own running means
           with the same kernels; ONE all-reduce(SUM) of a packed fp32 buffer per block merges them
           (mean over all samples = sum_r mean_r * n_r / sum_r n_r)
  phase 2  output rows are split across ranks for mask selection (rows are independent for the per-row
           and n:m rules); an all-gather returns the pruned weights and masks to every rank

torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests) is the plumbing; the statistics and
selection stay in the CUDA kernels.  The functions take the kernel entry points as arguments so the CPU
tests can drive the sharding / collective logic with the numpy oracle standing in for the device code.
"""
import torch
import torch.distributed as dist


def row_range(rows, rank, world):
    """Contiguous, balanced split of `rows` output rows; the first rows % world ranks get one extra."""
    base, extra = divmod(rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def sample_range(n_samples, rank, world):
    return row_range(n_samples, rank, world)


def merge_running_means(stats, n_local, group=None, n_total=None):
    """All-reduce a list of per-rank running means (each the mean over this rank's n_local samples) into the
    mean over all samples, in place, with one packed collective.  Returns the global sample count
    (pass n_total when every rank already knows it: that avoids a host sync)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return n_local
    dev = stats[0].device
    flat = torch.cat([s.reshape(-1).float() * float(n_local) for s in stats] +
                     [torch.full((1,), float(n_local), device=dev)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    n_total = float(flat[-1].item()) if n_total is None else float(n_total)
    off = 0
    for s in stats:
        k = s.numel()
        s.copy_((flat[off:off + k] / n_total).reshape(s.shape))
        off += k
    return int(round(n_total))


def merge_token_means(means, ntok_local, group=None, ntok_total=None):
    """Same for token-weighted means (DSnoT mean / var)."""
    return merge_running_means(means, ntok_local, group, ntok_total)


def gather_rows(full, rank, world, group=None):
    """All-gather the row shards of `full` [R, ...] in place: rank r owns rows row_range(R, r, world)."""
    if not dist.is_initialized() or world == 1:
        return full
    R = full.shape[0]
    if R % world == 0:
        s, e = row_range(R, rank, world)
        dist.all_gather_into_tensor(full, full[s:e].clone(), group=group)
    else:
        for r in range(world):
            s, e = row_range(R, r, world)
            if e > s:
                dist.broadcast(full[s:e], src=dist.get_global_rank(group, r) if group else r, group=group)
    return full


def prune_linear_row_sharded(weight, scaler_row, select_fn, rank, world, group=None):
    """Phase 2 for one linear.  select_fn(W_rows, scaler_row, keep_rows_out) -> score_mean (1-elem tensor) runs
    the selection kernel on this rank's row shard in place; returns (keep_mask [R, C], score_mean tensor)."""
    R, C = weight.shape
    s, e = row_range(R, rank, world)
    keep = torch.empty((R, C), dtype=torch.bool, device=weight.device)
    mean = torch.zeros(1, dtype=torch.float32, device=weight.device)
    if e > s:
        mean = select_fn(weight[s:e], scaler_row, keep[s:e]) * float(e - s)
    if dist.is_initialized() and world > 1:
        gather_rows(weight, rank, world, group)
        gather_rows(keep.view(torch.uint8), rank, world, group)
        dist.all_reduce(mean, op=dist.ReduceOp.SUM, group=group)
    return keep, mean / float(R)


def exchange_rows_packed(weight, keep, pack_fn, apply_fn, rank, world, group=None):
    """One linear whose row shard [row_range(R, rank, world)] of `keep` (mask bytes) and `weight` (zeroed) is final on
    this rank: all-gather the shard masks as bits and expand / apply the other shards locally (see
    prune_block_rows_packed).  Falls back to gathering bytes when the rows do not split evenly."""
    R, C = weight.shape
    if not dist.is_initialized() or world == 1:
        return keep
    if R % world or C % 16:
        gather_rows(weight, rank, world, group)
        gather_rows(keep.view(torch.uint8), rank, world, group)
        return keep
    s, e = row_range(R, rank, world)
    n = (e - s) * (C // 8)
    allbits = torch.empty(world * n, dtype=torch.uint8, device=weight.device)
    mine = allbits[rank * n:(rank + 1) * n]
    pack_fn(keep[s:e], mine.view(e - s, C // 8))
    dist.all_gather_into_tensor(allbits, mine.clone(), group=group)
    apply_fn(weight, allbits, keep, R // world, n)
    return keep


def prune_block_rows_packed(weights, select_fns, pack_fn, apply_fn, rank, world, group=None, select_batch_fn=None,
                            pack_batch_fn=None, apply_batch_fn=None):
    """Phase 2 for ALL linears of a block with ONE collective.  Every rank runs select_fns[i](W_rows, keep_rows) ->
    score_mean on its row shard of weights[i] (mask bytes + zeroed rows, in place), packs those mask rows to bits
    (pack_fn(keep_rows, bits_out)), all ranks all-gather the bits of all linears in one call (1 bit per weight: the
    weights are replicated, only the decision travels) and expand them with apply_fn(W, bits, keep,
    rows_per_seg, seg_stride), which also zeroes the pruned weights of the local replica.  Returns [(keep [R, C] bool, score_mean)].
    Needs R % world == 0 and C % 16 == 0 for every linear (callers fall back to prune_linear_row_sharded otherwise).
    select_batch_fn(W_rows_list, keep_rows_list) -> score means [len] (optional): selects on ALL row shards in one call
    (vlmc_wanda_rowselect_batch) instead of one select_fns[i] call per linear; same results.  pack_batch_fn(keep_rows_list,
    bits_list) and apply_batch_fn(weights, bits_views, keeps, rows_per_seg_list, seg_stride) (optional) do the same for the
    packing and the expand-and-zero pass: one launch each instead of one per linear."""
    dev = weights[0].device
    shapes = [tuple(w.shape) for w in weights]
    assert all(R % world == 0 and C % 16 == 0 for R, C in shapes)
    sizes = [(R // world) * (C // 8) for R, C in shapes]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    mine = torch.empty(offs[-1], dtype=torch.uint8, device=dev)
    keeps, means = [], torch.zeros(len(weights), dtype=torch.float32, device=dev)
    ranges = [row_range(R, rank, world) for R, _ in shapes]
    keeps = [torch.empty((R, C), dtype=torch.bool, device=dev) for R, C in shapes]
    if select_batch_fn is not None:
        m_all = select_batch_fn([w[s:e] for w, (s, e) in zip(weights, ranges)], [k[s:e] for k, (s, e) in zip(keeps, ranges)])
        means = m_all.reshape(-1) * (1.0 / float(world))          # R % world == 0: every shard is exactly 1 / world of its rows
    for i, (w, (R, C)) in enumerate(zip(weights, shapes)):
        s, e = ranges[i]
        keep = keeps[i]
        if select_batch_fn is None:
            m = select_fns[i](w[s:e], keep[s:e])
            means[i:i + 1] = m.reshape(1) * (float(e - s) / float(R))
        if pack_batch_fn is None:
            pack_fn(keep[s:e], mine[offs[i]:offs[i + 1]].view(e - s, C // 8))
    if pack_batch_fn is not None:
        pack_batch_fn([k[s:e] for k, (s, e) in zip(keeps, ranges)],
                      [mine[offs[i]:offs[i + 1]].view(ranges[i][1] - ranges[i][0], shapes[i][1] // 8) for i in range(len(weights))])
    if dist.is_initialized() and world > 1:
        allbits = torch.empty(world * offs[-1], dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allbits, mine, group=group)
        dist.all_reduce(means, op=dist.ReduceOp.SUM, group=group)
        # buffer layout [rank][linear shard]: ONE expand-and-zero call per linear covers the rows of every rank
        # (re-applying the own shard is idempotent)
        if apply_batch_fn is not None:
            apply_batch_fn(weights, [allbits[offs[i]:] for i in range(len(weights))], keeps, [R // world for R, _ in shapes], offs[-1])
        else:
            for i, (w, (R, C)) in enumerate(zip(weights, shapes)):
                apply_fn(w, allbits[offs[i]:], keeps[i], R // world, offs[-1])
    return [(k, means[i:i + 1]) for i, k in enumerate(keeps)]


# ---------------------------------------------------------------------------------------------------------------
# SparseGPT: tokens sharded for H, factorisations spread over ranks, rows sharded for the OBS sweep
# ---------------------------------------------------------------------------------------------------------------
def assign_factorisations(columns, world):
    """Which rank factorises which Hessian: longest-processing-time-first on the C^3 cost.
    columns: list of C per linear (in layer order).  Returns a list of owner ranks, deterministic on every rank."""
    load = [0.0] * world
    owner = [0] * len(columns)
    for i in sorted(range(len(columns)), key=lambda i: (-columns[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += float(columns[i]) ** 3
    return owner


def chain_cost_ms(R, C):
    """Rough duration of one linear's SparseGPT chain on one B200 (measured, r01): the blocked Cholesky + triangular
    inverse (latency-bound panels: 4.1 ms at C=4096, 20.6 ms at C=11008) plus the OBS sweep (C/128 dependent blocks of
    ~0.13 ms, wider with more rows).  Only the ORDER of the costs matters for the assignment."""
    return 4.1 * (C / 4096.0) ** 1.63 + 0.13 * ((C + 127) // 128) * max(1.0, R / 4096.0) ** 0.5


def assign_linears(shapes, world):
    """Which rank runs the factorisation + sweep chain of which linear: longest-processing-time-first on
    chain_cost_ms.  shapes: [(R, C)] in layer order.  Returns a list of owner ranks, deterministic on every rank."""
    load = [0.0] * world
    owner = [0] * len(shapes)
    for i in sorted(range(len(shapes)), key=lambda i: (-chain_cost_ms(*shapes[i]), i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += chain_cost_ms(*shapes[i])
    return owner


def prune_linears_task_parallel(weights, run_fn, rank, world, group=None, extra=None):
    """SparseGPT phase 2 + 3 on several GPUs: WHOLE linears are spread over the ranks (assign_linears) - each chain
    (factorisation, then column-block sweep) stays on one GPU, so nothing is exchanged per column block and no factor
    travels - and the pruned weights are broadcast from their owners.  run_fn(indices) prunes weights[i] in place for
    the linears this rank owns.  extra: optional list of per-linear tensors (masks) broadcast alongside.
    Returns the owner list."""
    owners = assign_linears([tuple(w.shape) for w in weights], world)
    mine = [i for i, o in enumerate(owners) if o == rank]
    if mine:
        run_fn(mine)
    for i, o in enumerate(owners):
        broadcast_from(weights[i], o, group)
        if extra is not None and extra[i] is not None:
            broadcast_from(extra[i], o, group)
    return owners


def allreduce_sum(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_max(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def broadcast_from(t, owner, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=dist.get_global_rank(group, owner) if group else owner, group=group)
    return t


def sharded_accumulate(accum_fn, x_local, state, n_total, group=None):
    """Phase 1 for any statistic that is a plain mean over samples (scaler_row, H): every rank runs ONE accumulation
    over its own sequences with the divisor of the whole calibration set (accum_fn(x_local, state, n_before=0,
    b=n_total)), so the partial results simply add: one SUM all-reduce, no rescaling pass."""
    accum_fn(x_local, state, 0, n_total)
    return allreduce_sum(state, group)


def factor_all(Hs, factor_fn, alloc_fn, rank, world, group=None):
    """Phase 2 of SparseGPT: the (sequential) Cholesky-inverse chains of one block are independent of each other, so
    they are spread over the ranks and the factors broadcast.  factor_fn(H) -> (U, dead) runs the damping loop;
    alloc_fn(H) -> (U, dead) returns empty receive buffers.  Returns [(U, dead)] for every Hessian on every rank."""
    owners = assign_factorisations([H.shape[0] for H in Hs], world)
    out = []
    for H, owner in zip(Hs, owners):
        out.append(factor_fn(H) if owner == rank else alloc_fn(H))
    for (U, dead), owner in zip(out, owners):
        broadcast_from(U, owner, group)
        broadcast_from(dead, owner, group)
    return out


def obs_rows_sharded(weight, U, dead, sweep_fn, rank, world, group=None):
    """Phase 3: rank r sweeps rows row_range(R, r, world) of `weight` in place; sweep_fn(W_rows, U, dead, rows_total,
    reduce_sum) exchanges only the block histograms (unstructured) through reduce_sum.  Pruned rows are gathered."""
    R = weight.shape[0]
    s, e = row_range(R, rank, world)
    if e > s:
        sweep_fn(weight[s:e], U, dead, R, lambda t: allreduce_sum(t, group))
    return gather_rows(weight, rank, world, group)


def importance_accumulate_data_parallel(batches, num_samples, grads_fn, accum_fn, packed_acc, rank, world, group=None):
    """SURVEY 8f-4 across ranks: the first-order statistic of LayerSparsity.compute_importance_scores
    (layer_single_base_pruner.py:440-458) is a plain sum over calibration batches, so batches are dealt round-robin
    (batch i -> rank i % world) from a loader every rank iterates identically, each rank accumulates its own, and ONE SUM
    all-reduce of the packed accumulator merges them.  The reference's stop rule - no batch is started once
    accum_samples >= num_samples (:442-443) - is evaluated on the global, in-order sample count: after every round of
    `world` batches the ranks exchange their batch lengths and a rank drops a batch the sequential loop would not have
    started.  grads_fn(batch) -> (grads, batch_len); accum_fn(grads) adds into the views of `packed_acc`.
    Returns the number of batches that count (the divisor of :463)."""
    accum_samples, num_batches = 0, 0
    it = iter(batches)
    done = False
    while not done:
        mine, n_round = None, 0
        for r in range(world):
            try:
                d = next(it)
            except StopIteration:
                done = True
                break
            n_round += 1
            if r == rank:
                mine = d
        if n_round == 0:
            break
        if accum_samples >= num_samples:                     # the sequential loop would have stopped before this round
            break
        lens = torch.zeros(world, dtype=torch.int64, device=packed_acc.device)
        grads = None
        if mine is not None:
            grads, blen = grads_fn(mine)
            lens[rank] = int(blen)
        allreduce_sum(lens, group)
        for r, blen in enumerate(lens.tolist()[:n_round]):
            if accum_samples >= num_samples:
                done = True
                break
            accum_samples += blen
            num_batches += 1
            if r == rank:
                accum_fn(grads)
    allreduce_sum(packed_acc, group)
    return num_batches
