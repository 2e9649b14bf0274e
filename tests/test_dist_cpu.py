"""world_size-2 `gloo` test of the host logic of the multi-GPU path: the fit's partial sums (Gram, X^T y, shifted
moments) of row shards add up to the single-process result after ONE all-reduce of one packed buffer, and the chain
shards of sample() cover all chains with disjoint global chain ids."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import bf_oracle as o
    from bayesfast_b200.runtime import shard_bounds, dist_info
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    r, w, g = dist_info(True)
    assert (r, w) == (rank, world)
    n, N = 4, 240
    rng = np.random.default_rng(3)
    x = rng.normal(size=(N, n)) + 0.2
    y = rng.normal(size=(N, 1))
    lo, hi = shard_bounds(N, rank, world)
    blocks = lambda xs: np.concatenate([o.design_block(k, xs) for k in ('linear', 'quadratic', 'cubic-2')], axis=1)
    A = blocks(x[lo:hi])
    P = A.shape[1]
    shift = np.zeros(n)
    buf = np.concatenate([(A.T @ A).ravel(), (A.T @ y[lo:hi]).ravel(), (x[lo:hi] - shift).sum(axis=0),
                          ((x[lo:hi] - shift).T @ (x[lo:hi] - shift)).ravel(), [hi - lo]])
    t = torch.from_numpy(buf)
    dist.all_reduce(t, group=g)                      # the single exchange step of the fit
    G = buf[:P * P].reshape(P, P)
    b = buf[P * P:P * P + P]
    coef = np.linalg.solve(G, b)
    s1 = buf[P * P + P:P * P + P + n]
    s2 = buf[P * P + P + n:-1].reshape(n, n)
    cnt = buf[-1]
    cov = (s2 - np.outer(s1, s1) / cnt) / (cnt - 1)
    # the product's own small collectives of the sharded fit (bayesfast_b200/runtime.py, fit.py): row count, largest radius,
    # the common shift of the moments, the center_max vote with a rank that has no valid candidate
    from bayesfast_b200.runtime import allreduce_values, pick_center
    from bayesfast_b200.fit import _common_shift
    tot = int(allreduce_values([hi - lo], 'sum', g, 0, 'int64')[0])
    mx = float(allreduce_values([1.5 + rank], 'max', g, 0)[0])
    sh = _common_shift(x[lo:hi], n, rank, world, g, 0)
    c1 = pick_center(7.5 if rank == 1 else -np.inf, np.full(n, 1. + rank), g, 0)          # rank 0: invalid logp
    c2 = pick_center(-np.inf, np.zeros(n), g, 0)                                           # nobody has a candidate
    extra = dict(tot=tot, mx=mx, shift=sh, c1=c1, c2_none=c2 is None)
    # chain sharding of sample()
    lo_c, hi_c = shard_bounds(37, rank, world)
    q.put((rank, coef, cov, cnt, (lo_c, hi_c), extra))
    dist.barrier()
    dist.destroy_process_group()


def test_fit_allreduce_and_chain_sharding_gloo(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, N = 4, 240
    rng = np.random.default_rng(3)
    x = rng.normal(size=(N, n)) + 0.2
    y = rng.normal(size=(N, 1))
    cfgs = [dict(order=k, input_mask=np.arange(n), output_mask=np.arange(1)) for k in ('linear', 'quadratic', 'cubic-2')]
    ref = oracle.fit(cfgs, n, 1, x, y)
    packed = np.concatenate([ref[0][0], ref[1][0][np.triu_indices(n)], ref[2][0].ravel()])
    for rank, coef, cov, cnt, _, ex in res:
        assert ex['tot'] == N and ex['mx'] == 2.5 and ex['c2_none']
        assert np.array_equal(ex['shift'], x[0]) and np.array_equal(ex['c1'], np.full(n, 2.))
        assert cnt == N
        assert np.allclose(coef, packed, rtol=1e-8, atol=1e-10)
        assert np.allclose(cov, np.cov(x, rowvar=False), rtol=1e-10)
    assert np.array_equal(res[0][1], res[1][1])          # bit-identical on both ranks: no broadcast needed
    assert res[0][4] == (0, 19) and res[1][4] == (19, 37)
