"""Stream-level scheduling of the independent chains of one transformer block.

The reference prunes the linears of a block one after the other (sparsegpt_pruner.py:441-452: for name in subset:
fasterprune; free).  Each SparseGPT.fasterprune is a SEQUENTIAL chain - blocked Cholesky (C/128 dependent panels),
then the column-block OBS sweep (C/128 dependent blocks) - whose kernels are small (one CTA for a diagonal block, a
few dozen tiles for a panel) and leave most of the 148 SMs idle.  The chains of different linears are independent
of each other, so here they run CONCURRENTLY, one CUDA stream per chain, forked from and joined to the caller's
stream with events.  Nothing about the arithmetic changes: the same kernels in the same per-chain order, so the
results are bit-identical to the one-after-the-other schedule (tests/test_gpu_parity.py).

  factor_concurrent    all distinct Hessians of a block: dead channels + damp value + chol_inv_upper, ONE host
                       sync for all status words, the reference's damp-and-retry loop (:114-128) only for the
                       ones that failed
  sweep_concurrent     all OBS sweeps of a block
  sparsegpt_block      both, pipelined per Hessian, for a list of (W, H) items whose H may be shared (q/k/v, gate/up):
                       factorised once, its sweeps start as soon as it is done (guarded by the status word)

Outputs that outlive the fork (U, dead, scores) are allocated on the caller's stream BEFORE the fork, so the caching
allocator never hands their memory to another stream; per-stream scratch comes from native.workspace (keyed by stream).
"""
import torch

from vlmc import native

_pools = {}


def side_streams(device, n):
    """n persistent side streams of `device` (created once: native.workspace keys its scratch by stream)."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    pool = _pools.setdefault(idx, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


class Fork:
    """with Fork(device, n) as f:  for i in range(n): with f.stream(i): ...enqueue chain i...
    On exit the caller's stream waits for every side stream."""

    def __init__(self, device, n):
        self.device = torch.device(device)
        self.n = n

    def __enter__(self):
        self.main = torch.cuda.current_stream(self.device)
        self.streams = side_streams(self.device, self.n)
        start = torch.cuda.Event()
        start.record(self.main)
        for s in self.streams:
            s.wait_event(start)
        return self

    def stream(self, i):
        return torch.cuda.stream(self.streams[i % self.n])

    def __exit__(self, *exc):
        for s in self.streams:
            done = torch.cuda.Event()
            done.record(s)
            self.main.wait_event(done)
        return False


def clamp_infinite(M):
    """sparsegpt_pruner.py:101-109 (H) and :133-141 (H^-1): +inf entries become quantile(M, 0.999), then -inf entries
    quantile(M, 0.001) of the matrix as it then stands (torch.quantile over ALL entries, the infinite ones included).
    Returns the number of entries replaced.  The reference spins forever on a NaN (its quantile is NaN and the Cholesky
    loop never succeeds): here that raises.  Unlike torch.quantile there is no 16 M element limit (C >= 4096)."""
    pos, neg, nan = native.matrix_nonfinite_count(M)
    if nan:
        raise RuntimeError("the matrix holds NaN entries: the reference's damping loop never terminates on it")
    if pos:
        native.matrix_replace_inf(M, native.matrix_quantile(M, 0.999), negative=False)
    if neg:
        native.matrix_replace_inf(M, native.matrix_quantile(M, 0.001), negative=True)
    return pos + neg


def second_stage(U, percdamp=0.01, status=None, max_retries=64):
    """The reference's second stage in its own order (sparsegpt_pruner.py:131-157) starting from the fused factor U of
    vlmc_chol_inv_upper: Hinv = U^T U (= cholesky_inverse), the +-inf clamp, damp2 = percdamp * mean|diag Hinv| and the
    damp-and-retry loop around cholesky(Hinv, upper=True).  U is overwritten with the result.  Returns the number of
    damping steps taken.  Mathematically the identity when nothing is clamped or damped."""
    return second_stage_from_inverse(native.gram_upper(U), U, percdamp, status, max_retries)


def second_stage_from_inverse(Hinv, U=None, percdamp=0.01, status=None, max_retries=64):
    """sparsegpt_pruner.py:133-157 on a given H^-1 (modified in place: clamp, damping).  Writes U (allocated when None);
    returns the number of damping steps.  Read the factor from the U you passed in."""
    clamp_infinite(Hinv)
    damp2 = native.diag_abs_mean(Hinv, percdamp)
    U = torch.empty_like(Hinv) if U is None else U
    status = torch.empty(1, dtype=torch.int32, device=U.device) if status is None else status
    for steps in range(max_retries + 1):
        native.chol_upper(Hinv, U, status)
        st = int(status.item())
        if st == 0:
            return steps
        if st & native.NONFINITE:
            raise RuntimeError("H^-1 holds NaN entries: the reference's damping loop never terminates on it")
        native.hessian_add_damp(Hinv, damp2)
    raise RuntimeError("H^-1 stayed non-positive-definite after damping")


def resolve_factor(H, U, status, damp, percdamp=0.01, max_retries=64, exact_reference_order=False):
    """Host side of one factorisation AFTER its first vlmc_chol_inv_upper attempt was enqueued: reads the status word and
    runs the reference's control flow for whatever it reports -
      VLMC_NONFINITE    clamp the +-inf entries of H (:101-109), recompute damp = percdamp * mean(diag H) (:111), retry
      VLMC_NOT_POSDEF   H[diag] += damp, cumulatively per retry (:114-128)
      VLMC_HUGE_FACTOR  (or exact_reference_order) the second stage in the reference's own order (:131-157)
    Returns True when anything had to be redone (the caller then sweeps again)."""
    st = int(status.item())
    redone = False
    retries = 0
    while st & (native.NONFINITE | native.NOT_POSDEF):
        if st & native.NONFINITE:
            if clamp_infinite(H) == 0:
                raise RuntimeError("non-finite Hessian without +-inf entries")      # unreachable: NaN raises above
            native.hessian_prepare(H, percdamp, damp)
        else:
            native.hessian_add_damp(H, damp)
        retries += 1
        if retries > max_retries:
            raise RuntimeError("Hessian stayed non-positive-definite after damping")
        native.chol_inv_upper(H, U, status)
        st = int(status.item())
        redone = True
    if (st & native.HUGE_FACTOR) or exact_reference_order:
        second_stage(U, percdamp, status, max_retries)
        redone = True
    return redone


def factor_concurrent(Hs, percdamp=0.01, Us=None, max_retries=64, exact_reference_order=False):
    """[(U, dead)] for a list of DISTINCT Hessians (each modified in place like the reference: dead diagonal -> 1,
    damping added only after a failed attempt, sparsegpt_pruner.py:95-96,111-128).  Us: optional output buffers."""
    dev = Hs[0].device
    n = len(Hs)
    Us = [torch.empty_like(H) for H in Hs] if Us is None else Us
    damps = torch.empty(n, dtype=torch.float32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    deads = [torch.empty(H.shape[0], dtype=torch.uint8, device=dev) for H in Hs]
    order = sorted(range(n), key=lambda i: -Hs[i].shape[0])      # longest chain first
    # several chains at once: the Cholesky look-ahead's side streams only add contention (measured), one chain keeps it
    with Fork(dev, n) as f, native.chol_lookahead(n == 1):
        for slot, i in enumerate(order):
            with f.stream(slot):
                native.hessian_prepare(Hs[i], percdamp, damps[i:i + 1], deads[i])
                native.chol_inv_upper(Hs[i], Us[i], status[i:i + 1])
    failed = [i for i, s in enumerate(status.tolist()) if s != 0 or exact_reference_order]      # the ONE host sync of the phase
    for i in failed:                                                     # :101-157 for the ones that need it
        resolve_factor(Hs[i], Us[i], status[i:i + 1], damps[i:i + 1], percdamp, max_retries, exact_reference_order)
    return list(zip(Us, deads))


def sweep_concurrent(items, blocksize=128):
    """items: [(W, U, dead, sparsity, prune_n, prune_m)], W pruned in place (obs_sweep).  Returns the importance
    scores as one [n] device tensor."""
    dev = items[0][0].device
    n = len(items)
    scores = torch.empty(n, dtype=torch.float32, device=dev)
    order = sorted(range(n), key=lambda i: -items[i][0].shape[1])
    with Fork(dev, n) as f:
        for slot, i in enumerate(order):
            W, U, dead, sparsity, pn, pm = items[i]
            with f.stream(slot):
                native.obs_sweep(W, U, sparsity, pn, pm, dead=dead, blocksize=blocksize, score=scores[i:i + 1])
    return scores


def sparsegpt_block(items, percdamp=0.01, blocksize=128, Us=None, max_retries=64, exact_reference_order=False):
    """items: [(W, H, sparsity, prune_n, prune_m)].  Items that hold THE SAME H tensor (linears fed by the same
    activations) are factorised once.  Returns (scores [n] device tensor, {id(H): (U, dead)}).

    Pipelined: every distinct H is prepared and factorised on its own stream and the sweeps of ITS linears start as soon
    as that factorisation is done (each on its own stream, behind an event) - they do not wait for the longest
    factorisation of the block.  The sweeps are enqueued before the host knows whether the factorisation succeeded:
    they carry its device status word and leave W untouched if it failed (vlmc_obs_sweep_guarded).  One host sync at
    the end reads all status words; a failed H then goes through the reference's damp-and-retry loop
    (sparsegpt_pruner.py:114-128) and its linears are swept again.  Same kernels in the same per-chain order as
    SparseGPT.fasterprune: bit-identical weights."""
    distinct, seen, members = [], {}, []
    for i, (_, H, *_rest) in enumerate(items):
        if id(H) not in seen:
            seen[id(H)] = len(distinct)
            distinct.append(H)
            members.append([])
        members[seen[id(H)]].append(i)
    dev = distinct[0].device
    nH, n = len(distinct), len(items)
    Us = [torch.empty_like(H) for H in distinct] if Us is None else Us
    damps = torch.empty(nH, dtype=torch.float32, device=dev)
    status = torch.empty(nH, dtype=torch.int32, device=dev)
    deads = [torch.empty(H.shape[0], dtype=torch.uint8, device=dev) for H in distinct]
    scores = torch.empty(n, dtype=torch.float32, device=dev)
    order = sorted(range(nH), key=lambda h: -distinct[h].shape[0])      # longest chain first
    with Fork(dev, nH + n) as f, native.chol_lookahead(nH == 1):       # see factor_concurrent
        for slot, h in enumerate(order):
            with f.stream(slot):
                native.hessian_prepare(distinct[h], percdamp, damps[h:h + 1], deads[h])
                native.chol_inv_upper(distinct[h], Us[h], status[h:h + 1])
                factored = torch.cuda.Event()
                factored.record(f.streams[slot])
            for i in ([] if exact_reference_order else members[h]):     # exact order: swept after the second stage
                W, _, sparsity, pn, pm = items[i]
                f.streams[nH + i].wait_event(factored)
                with f.stream(nH + i):
                    native.obs_sweep(W, Us[h], sparsity, pn, pm, dead=deads[h], blocksize=blocksize,
                                     score=scores[i:i + 1], fail_flag=status[h:h + 1])
    failed = [h for h, st in enumerate(status.tolist()) if st != 0 or exact_reference_order]     # the ONE host sync of the block
    for h in failed:                                                     # :101-157 for the ones that need it
        resolve_factor(distinct[h], Us[h], status[h:h + 1], damps[h:h + 1], percdamp, max_retries, exact_reference_order)
        for i in members[h]:
            W, _, sparsity, pn, pm = items[i]
            native.obs_sweep(W, Us[h], sparsity, pn, pm, dead=deads[h], blocksize=blocksize, score=scores[i:i + 1])
    return scores, {id(H): (Us[h], deads[h]) for h, H in enumerate(distinct)}
