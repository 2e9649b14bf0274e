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
    # chain sharding of sample()
    lo_c, hi_c = shard_bounds(37, rank, world)
    q.put((rank, coef, cov, cnt, (lo_c, hi_c)))
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
    for rank, coef, cov, cnt, _ in res:
        assert cnt == N
        assert np.allclose(coef, packed, rtol=1e-8, atol=1e-10)
        assert np.allclose(cov, np.cov(x, rowvar=False), rtol=1e-10)
    assert np.array_equal(res[0][1], res[1][1])          # bit-identical on both ranks: no broadcast needed
    assert res[0][4] == (0, 19) and res[1][4] == (19, 37)
