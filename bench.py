#!/usr/bin/env python
"""
bench.py -- NUTS leapfrog-steps x chains / s, cubic-2 PolyModel surrogate, d=26 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one complete NUTS run (n_iter=1500, n_warmup=500: reference defaults, sample_trace.py:499-512)
of 4096 chains per GPU on the 26-D DES-Y1-shaped synthetic posterior (SURVEY.md 8d config 3 = BASELINE.json
configs[2]: 32768 chains over 8 GPUs) through its fitted cubic-2 surrogate: ONE kernel launch (+ device-side reset of
the chain state).  Weak scaling: every rank runs its own 4096 chains (global chain ids -> distinct Philox streams), no
data-path collective.  The dominant kernel is nuts_dmma_kernel (bfb_sampler_dmma.cu: 8 chains per warp as the rows of
FP64 m8n8k4 DMMAs).  4096 chains are 512 warps -- fewer than the 592 warp schedulers of a B200 -- so the line also
carries `more_chains` (north_star: "at least 4096 chains per GPU"): the same run with 16384 chains on this GPU, and
`eval_kernel`: the batched surrogate evaluation kernel alone (bfb_eval_dmma.cu) against the same FP64 peak.
`pipeline_kernel`: SURVEY 8f rank 1 -- DES-Y1-shaped surrogate -> Gaussian-likelihood pipeline (n = 26, m = 457 outputs) on the
tensor cores: batched logp + gradient (bfb_lik_dmma.cu) and a 4096-chain NUTS run (model variant bit 3), same FP64 peak.

  value : sum(tree_size) of all ranks / max-over-ranks device time of K steps, inputs resident in HBM
  e2e   : the same through bayesfast_b200.sample() with host x_0 and all samples + stats copied back
  roofline : algorithmic FP64 flops (8n^2+24n per chain-leapfrog) / kernel time (CUDA events on
             the launching stream) vs the FP64 peak measured in this run (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline : the oracle (C restatement of the reference path, OpenMP over chains) on this box's host cores,
             bounded sample of the same workload.  `--impl reference` runs only that.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DIM, ORDER, CHAINS_PER_GPU, N_ITER, N_WARMUP, SEED = 26, 'cubic-2', 4096, 1500, 500, 20261017
FLOPS_PER_LEAF = 8 * N_DIM * N_DIM + 24 * N_DIM        # 6032 at n=26
WORKLOAD = 'des_y1_shaped_d26_cubic2_nuts_4096_chains_per_gpu_n_iter1500_n_warmup500'


def clocks_sampler(stop, out, device):
    q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    while not stop.is_set():
        try:
            r = subprocess.run(['nvidia-smi', '-i', str(device), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(',')
            out.append([v.strip() for v in r])
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return None
    sm = sorted(float(s[0]) for s in samples if s[0].replace('.', '').isdigit())
    names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
    reasons = [nm for i, nm in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith('active') for s in samples)]
    return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(samples[0][1]) if samples[0][1] else None,
                power_w_max=max(float(s[2]) for s in samples if s[2].replace('.', '').isdigit()), reasons=reasons,
                samples=len(samples))


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_spec_cpu(prob):
    """surrogate fitted on the CPU with the oracle (scipy lstsq), for the reference arm"""
    from oracle import bf_oracle as o
    n = prob['n']
    cfgs = [dict(order=k, input_mask=np.arange(n), output_mask=np.arange(1)) for k in ('linear', 'quadratic', 'cubic-2')]
    coefs = o.fit(cfgs, n, 1, prob['x_fit'], prob['y_fit'])
    for c, a in zip(cfgs, coefs):
        c['coef'] = a
    mu, hess, alpha = o.bound_from_points(prob['x_fit'])
    spec = dict(n=n, m=1, configs=cfgs, use_bound=False, input_scales=None, use_decay=False, transform_ranges=None)
    xm = prob['x_fit'][int(np.argmax(prob['y_fit'][:, 0]))]
    f_mu = o.OracleDensity(spec).poly_eval_batch(xm[None])[0][0]
    spec.update(use_bound=True, mu=mu, hess=hess, alpha=alpha, f_mu=f_mu)
    return spec


def run_cpu_baseline(spec, prob, budget_s=15., n_threads=0):
    """the oracle port on all host cores: bounded sample = as many full chains (n_iter=1500) as fit in ~budget_s"""
    from oracle import bf_oracle as o
    cores = cpu_cores() if n_threads <= 0 else n_threads
    od = o.OracleDensity(spec)
    n = prob['n']
    cfg = dict(n_iter=N_ITER, n_warmup=N_WARMUP, seed=SEED, n_threads=cores)
    step0 = 1. / n**0.25
    t0 = time.time()
    r = od.run('NUTS', cfg, prob['x_0'][:cores], step0, np.ones(n))
    t_probe = time.time() - t0
    chains = int(max(cores, min(prob['x_0'].shape[0], cores * max(1, int(budget_s / max(t_probe, 1e-3))))))
    t0 = time.time()
    r = od.run('NUTS', cfg, prob['x_0'][:chains], step0, np.ones(n))
    dt = time.time() - t0
    leaves = int(r['tree_size'].sum())
    return dict(value=leaves / dt, unit='leapfrog-steps*chains/s', cores=cores, kind='port',
                sample='{} full chains (n_iter={}, n_warmup={}) of the same workload, {} leapfrogs in {:.1f} s, oracle/'
                       'libbf_oracle.so (C restatement, OpenMP over chains); the Python reference itself measured '
                       '3.2e4/s on 8 cores (BASELINE.md)'.format(chains, N_ITER, N_WARMUP, leaves, dt)), leaves, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--chains-per-gpu', type=int, default=CHAINS_PER_GPU)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=15.)
    ap.add_argument('--no-extras', action='store_true', help='skip the more_chains / eval_kernel supplementary measurements')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    from bayesfast_b200 import synthetic
    C = args.chains_per_gpu
    prob = synthetic.des_shaped(N_DIM, seed=1, n_chain=C * max(world, 1), order=ORDER)
    config = dict(workload=WORKLOAD, n=N_DIM, order=ORDER, n_param=synthetic.n_param(ORDER, N_DIM),
                  chains_per_gpu=C, n_iter=N_ITER, n_warmup=N_WARMUP, n_fit=int(prob['x_fit'].shape[0]),
                  step='one full NUTS run of all chains (one kernel launch)',
                  l2='256 MiB scratch memset between timed steps (L2 flush); outputs (1.8 GB/step) exceed L2',
                  parallelism='chains sharded over {} GPU(s), no data-path collective'.format(world))

    if args.impl == 'reference':
        if rank != 0:
            return
        spec = oracle_spec_cpu(prob)
        vals = []
        for s in range(args.warmup + args.steps):
            cb, leaves, dt = run_cpu_baseline(spec, prob, budget_s=min(args.cpu_budget, 10.))
            if s >= args.warmup:
                vals.append((leaves, dt))
        leaves = sum(v[0] for v in vals)
        dt = sum(v[1] for v in vals)
        cb['value'] = leaves / dt
        print(json.dumps(dict(metric='nuts_leapfrog_steps_x_chains_per_s', value=leaves / dt, unit='leapfrog-steps*chains/s',
                              impl='reference', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                              ms_per_step=dt / max(args.steps, 1) * 1e3, higher_is_better=True, scaling='weak',
                              vs_baseline=None, dtype='f64', data='synthetic', config=config, cpu_baseline=cb,
                              e2e=dict(value=leaves / dt, unit='leapfrog-steps*chains/s', h2d_bytes_per_step=0,
                                       d2h_bytes_per_step=0), gpu_launches=0)))
        return

    import torch
    import torch.distributed as dist
    import bayesfast_b200 as bfb
    from bayesfast_b200 import _cabi
    numa_cpus = None
    if world > 1:
        from bayesfast_b200.runtime import bind_to_gpu_numa
        numa_cpus = bind_to_gpu_numa(local)       # host buffers of this rank on the NUMA node of its GPU
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = local

    # ---- setup (untimed): fit the surrogate on the GPU, build the density ----
    sur = bfb.PolyModel(ORDER, input_size=N_DIM, output_size=1, device=dev)
    t0 = time.time()
    sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
    fit_s = time.time() - t0            # first call of the process: CUDA context, module load and allocations included
    t0 = time.time()
    sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
    fit_warm_s = time.time() - t0       # the same fit again (steady state: host copies, Gram kernel, solve, bound)
    den = bfb.Density(sur)
    h = den._sync(False)
    peak = max(h.fp64_peak(0) for _ in range(2))               # TFLOP/s, DFMA, measured now on this GPU
    x0 = np.ascontiguousarray(prob['x_0'][rank * C:(rank + 1) * C])
    trace_kw = dict(n_chain=C, n_iter=N_ITER, n_warmup=N_WARMUP, x_0=x0, random_generator=SEED)
    cfg = bfb.NTrace(**trace_kw)._cfg_dict(SEED, rank * C)
    h.sampler_init(cfg, x0, 1. / N_DIM**0.25, np.ones(N_DIM), x0)
    S = C * N_ITER
    d_out = dict(samples=torch.empty(S * N_DIM, dtype=torch.float64, device='cuda:%d' % dev))
    for k in _cabi.FLOAT_STATS:
        d_out[k] = torch.empty(S, dtype=torch.float64, device='cuda:%d' % dev)
    for k in _cabi.INT_STATS:
        d_out[k] = torch.empty(S, dtype=torch.int32, device='cuda:%d' % dev)
    ptrs = {k: v.data_ptr() for k, v in d_out.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:%d' % dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        h.synchronize()

    def one_step():
        h.sampler_reset()
        r = h.sampler_run('NUTS', N_ITER, out_ptrs=ptrs)
        return r['total_tree_size'], h.last_kernel_ms()

    for _ in range(args.warmup):
        one_step()
        flush.zero_()
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, dev), daemon=True)
    th.start()
    l0 = h.launch_count()
    leaves = 0
    kern_ms = []
    step_ms = 0.
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        lv, kms = one_step()                  # returns after the stream has been synchronised
        barrier()
        step_ms += (time.perf_counter() - t0) * 1e3
        leaves += lv
        kern_ms.append(kms)
    launches = h.launch_count() - l0
    kernel_family = h.sampler_last_path()
    # device time of the timed steps: CUDA events around the kernel on the launching stream (+ the reset copies,
    # which the wall clock above includes); report the event time, max over ranks
    dev_ms = float(sum(kern_ms))
    if world > 1:
        t = torch.tensor([dev_ms, step_ms], dtype=torch.float64, device='cuda:%d' % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, step_ms = float(t[0]), float(t[1])
        t = torch.tensor([leaves, launches], dtype=torch.int64, device='cuda:%d' % dev)
        dist.all_reduce(t)
        leaves_all, launches_all = int(t[0]), int(t[1])
    else:
        leaves_all, launches_all = leaves, launches
    value = leaves_all / (step_ms * 1e-3)

    # ---- e2e: the public API, host buffers in and out ----
    e2e_leaves, e2e_ms, h2d, d2h, e2e_kernel_ms = 0, 0., 0, 0, 0.
    n_e2e = max(1, min(args.steps, 3))
    for i in range(1 + n_e2e):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        tt = bfb.sample(den, dict(trace_kw), verbose=False)
        barrier()
        if i > 0:
            e2e_ms += (time.perf_counter() - t0) * 1e3
            e2e_kernel_ms += tt.kernel_ms
            e2e_leaves += tt.total_tree_size
            h2d = x0.nbytes * 2 + 8 * C + x0.nbytes
            d2h = sum(v.nbytes for k, v in tt.arrays.items() if k not in ('samples_original', 'logp_original'))
        del tt
    stop.set()
    th.join(timeout=2)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device='cuda:%d' % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
        t = torch.tensor([e2e_leaves], dtype=torch.int64, device='cuda:%d' % dev)
        dist.all_reduce(t)
        e2e_leaves = int(t[0])
    e2e_value = e2e_leaves / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    k_ms = float(np.mean(kern_ms))
    achieved = FLOPS_PER_LEAF * (leaves / args.steps) / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get('sampler_kernel_dram_bytes_per_launch')
    except Exception:
        pass
    hbm_bytes = (leaves / args.steps) * 0 + C * N_ITER * (N_DIM * 8 + 7 * 8 + 3 * 4)
    roofline = dict(bound='tensor', pipe='FP64 pipe: the kernel evaluates the surrogate with m8n8k4 DMMAs; DFMA and DMMA issue to the same '
                    'pipe (measured 37.0 vs 37.2 TFLOP/s, 31 when interleaved)', kernel='nuts_%s_kernel' % kernel_family, achieved=achieved, peak=peak, unit='TFLOP/s',
                    frac=achieved / peak, traffic=traffic,
                    peak_source='bfb_fp64_peak DFMA microbenchmark run just before the timed region on this GPU '
                                '(MEASURED_PEAKS.json has no FP64 entry; nominal 37 TFLOP/s)',
                    algorithmic_flops_per_leapfrog=FLOPS_PER_LEAF, kernel_ms=k_ms,
                    hbm=dict(algorithmic_bytes_per_launch=hbm_bytes, achieved_gbs=hbm_bytes / (k_ms * 1e-3) / 1e9,
                             peak_gbs=peaks.get('hbm_gbs'), note='outputs only; the tree state stays in shared memory'))
    extras = {}
    if not args.no_extras:
        # (a) the same run with 16384 chains on this GPU (two warps per scheduler instead of one): kernel-only value
        C2 = 16384
        prob2 = synthetic.des_shaped(N_DIM, seed=1, n_chain=C2, order=ORDER)
        cfg2 = bfb.NTrace(n_chain=C2, n_iter=N_ITER, n_warmup=N_WARMUP, x_0=prob2['x_0'], random_generator=SEED)._cfg_dict(SEED, 0)
        h.sampler_init(cfg2, prob2['x_0'], 1. / N_DIM**0.25, np.ones(N_DIM), prob2['x_0'])
        best = None
        for i in range(3):
            h.sampler_reset()
            flush.zero_()
            torch.cuda.synchronize(dev)
            r2 = h.sampler_run('NUTS', N_ITER, out_ptrs={})
            ms2 = h.last_kernel_ms()
            if i > 0 and (best is None or ms2 < best[1]):
                best = (r2['total_tree_size'], ms2)
        tf2 = FLOPS_PER_LEAF * best[0] / (best[1] * 1e-3) / 1e12
        extras['more_chains'] = dict(chains_per_gpu=C2, value=best[0] / (best[1] * 1e-3), unit='leapfrog-steps*chains/s',
                                     kernel_ms=best[1], outputs='none written (kernel only)', kernel=h.sampler_last_path(),
                                     roofline_frac=tf2 / peak, achieved_tflops=tf2)
        # (a2) lock-step HMC (hmc.py:16-49, n_int_step = 32) on the same chains: the tensor-core integrator without the NUTS
        # tree bookkeeping (hmc_dmma_kernel), kernel only
        hm = {}
        for Ch, x0h in ((C, x0), (C2, prob2['x_0'])):
            cfgh = bfb.HTrace(n_chain=Ch, n_iter=300, n_warmup=100, x_0=x0h, n_int_step=32, random_generator=SEED)._cfg_dict(SEED, 0)
            h.sampler_init(cfgh, x0h, 1. / N_DIM**0.25, np.ones(N_DIM), x0h)
            h.sampler_run('HMC', 100, out_ptrs={})                         # step-size adaptation
            rh = h.sampler_run('HMC', 100, out_ptrs={})
            msh = h.last_kernel_ms()
            tfh = FLOPS_PER_LEAF * rh['total_tree_size'] / (msh * 1e-3) / 1e12
            hm[str(Ch)] = dict(chains_per_gpu=Ch, value=rh['total_tree_size'] / (msh * 1e-3), unit='leapfrog-steps*chains/s',
                               kernel_ms=msh, roofline_frac=tfh / peak, achieved_tflops=tfh)
        extras['hmc_kernel'] = dict(kernel='hmc_dmma_kernel', n_int_step=32, iterations=100, outputs='none written (kernel only)', runs=hm)
        # (b) the surrogate evaluation kernel alone: logp + gradient of 2^22 device-resident points
        Ce = 1 << 22
        Xe = torch.randn(Ce, N_DIM, dtype=torch.float64, device='cuda:%d' % dev) @ torch.tensor(np.linalg.cholesky(prob['cov']).T, device='cuda:%d' % dev)
        Xe = Xe.contiguous()
        lpe = torch.empty(Ce, dtype=torch.float64, device='cuda:%d' % dev)
        ge = torch.empty(Ce, N_DIM, dtype=torch.float64, device='cuda:%d' % dev)
        torch.cuda.synchronize(dev)
        mse = []
        for i in range(6):
            h.logp_and_grad_batch_dev(Xe.data_ptr(), Ce, lpe.data_ptr(), ge.data_ptr())
            mse.append(h.last_kernel_ms())
        mse = float(np.mean(mse[1:]))
        fe = (8 * N_DIM * N_DIM + 15 * N_DIM) * Ce
        extras['eval_kernel'] = dict(kernel='eval_dmma_kernel', points=Ce, ms=mse, points_per_s=Ce / mse * 1e3,
                                     algorithmic_flops_per_point=8 * N_DIM * N_DIM + 15 * N_DIM,
                                     roofline=dict(bound='tensor', achieved=fe / mse / 1e9, peak=peak, unit='TFLOP/s', frac=fe / mse / 1e9 / peak),
                                     hbm_gbs=Ce * (2 * N_DIM + 1) * 8 / mse / 1e6)
        del Xe, lpe, ge
        # (c) SURVEY 8f rank 1: DES-Y1-shaped surrogate -> Gaussian-likelihood pipeline (n = 26, m = 457 block-quadratic outputs,
        # dense inverse covariance) on the tensor cores: batched logp + gradient (bfb_lik_dmma.cu) and NUTS (model variant bit 3)
        try:
            from bayesfast_b200 import _cabi
            from bayesfast_b200.density import whiten_spec
            pspec, plik = synthetic.des_pipeline(N_DIM, 457, seed=0)
            hp = _cabi.Handle(dev)
            hp.set_model(whiten_spec(pspec, plik))
            Cp = 1 << 16
            Xp = (torch.randn(Cp, N_DIM, dtype=torch.float64, device='cuda:%d' % dev) * 0.3).contiguous()
            lpp = torch.empty(Cp, dtype=torch.float64, device='cuda:%d' % dev)
            gp = torch.empty(Cp, N_DIM, dtype=torch.float64, device='cuda:%d' % dev)
            torch.cuda.synchronize(dev)
            msp = []
            for i in range(5):
                hp.logp_and_grad_batch_dev(Xp.data_ptr(), Cp, lpp.data_ptr(), gp.data_ptr())
                msp.append(hp.last_kernel_ms())
            msp = float(np.mean(msp[1:]))
            fl = 457 * (2 * N_DIM * N_DIM + 5 * N_DIM) + 9 * N_DIM
            x0p = np.random.default_rng(1).normal(size=(C, N_DIM)) * 0.2
            cfgp = bfb.NTrace(n_chain=C, n_iter=200, n_warmup=100, x_0=x0p, random_generator=SEED)._cfg_dict(SEED, 0)
            hp.sampler_init(cfgp, x0p, 1. / N_DIM**0.25, np.ones(N_DIM), x0p)
            rp = hp.sampler_run('NUTS', 200, out_ptrs={})
            msn = hp.last_kernel_ms()
            extras['pipeline_kernel'] = dict(
                workload='des_y1_shaped_pipeline_n26_m457_gaussian_likelihood', algorithmic_flops_per_evaluation=fl,
                eval=dict(kernel='lik_eval_dmma_kernel' if hp.eval_last_path() == 'lik_dmma' else hp.eval_last_path(), points=Cp, ms=msp,
                          points_per_s=Cp / msp * 1e3, roofline=dict(bound='tensor', achieved=fl * Cp / msp / 1e9, peak=peak,
                                                                     unit='TFLOP/s', frac=fl * Cp / msp / 1e9 / peak)),
                nuts=dict(kernel='nuts_%s_kernel' % hp.sampler_last_path(), chains_per_gpu=C, iterations=200, kernel_ms=msn,
                          value=rp['total_tree_size'] / msn * 1e3, unit='leapfrog-steps*chains/s',
                          roofline_frac=fl * rp['total_tree_size'] / msn / 1e9 / peak))
            hp.close()
            del Xp, lpp, gp
        except Exception as exc:                                   # supplementary measurement: never fails the bench line
            extras['pipeline_kernel'] = dict(error=repr(exc))
    out = dict(metric='nuts_leapfrog_steps_x_chains_per_s', value=value, unit='leapfrog-steps*chains/s', n_gpus=world,
               steps=args.steps, warmup=args.warmup, ms_per_step=step_ms / args.steps, higher_is_better=True,
               scaling='weak', vs_baseline=None, dtype='f64', data='synthetic', config=config,
               e2e=dict(value=e2e_value, unit='leapfrog-steps*chains/s', h2d_bytes_per_step=int(h2d),
                        d2h_bytes_per_step=int(d2h), ms_per_step=e2e_ms / n_e2e, kernel_ms_per_step=e2e_kernel_ms / n_e2e,
                        numa_cpus_rank0=(len(numa_cpus) if numa_cpus else None)),
               gpu_launches=int(launches_all), roofline=roofline, clocks=summarize_clocks(samples),
               fit=dict(seconds=fit_s, seconds_warm=fit_warm_s, kernel_ms=getattr(sur, '_fit_kernel_ms', None), n=N_DIM,
                        P=config['n_param'], N=config['n_fit'], rel_resid=getattr(sur, '_fit_rel_resid', None)),
               mean_tree_size=leaves / args.steps / (C * N_ITER), kernel='nuts_%s_kernel' % kernel_family, **extras)
    if world == 1 and not args.no_cpu_baseline:
        spec = den.to_spec()
        out['cpu_baseline'] = run_cpu_baseline(spec, prob, budget_s=args.cpu_budget)[0]
        out['cpu_baseline']['gpu_over_cpu_e2e'] = e2e_value / out['cpu_baseline']['value']
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
