// bfb_sampler_team.cu -- NUTS / HMC with EIGHT chains per TEAM of four warps (bfb_team.cuh): the tensor-core sampler for
// linear + quadratic (+ cubic-2) surrogates with or without radial bound, input_size <= 32.
//
// Reference restated here: samplers/hmc_utils/base_hmc.py:62-85, samplers/nuts.py:27-217, samplers/hmc.py:16-49,
// hmc_utils/integration.py:28-95, hmc_utils/metrics.py:73-91,186-211,333-371, hmc_utils/step_size.py:10-51.
#include "bfb_team.cuh"
#include "bfb_nuts_common.cuh"
#include <cstring>
#include <cstdlib>

static __device__ __noinline__ double team_draw_normal(uint64_t seed, uint64_t chain, uint64_t t) { return bfb_draw_normal(seed, chain, t); }
static __device__ __noinline__ double team_draw_uniform(uint64_t seed, uint64_t chain, uint64_t t) { return bfb_draw_uniform(seed, chain, t); }

// Work queue (same layout as queue_init_kernel, bfb_nuts_common.cuh), but the first `head0` units are pre-assigned: team
// slot s = blockIdx.x + gridDim.x * team takes unit s without an atomic, so that the teams of ALL blocks get work when there
// are fewer groups than team slots (4096 chains = 512 groups on 148 x 4 = 592 slots).
static __global__ void team_queue_init_kernel(int *queue, int n_groups, int n_units, int head0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { queue[0] = head0; queue[1] = n_groups; }
    if (i < n_groups) queue[2 + i] = 0;
    if (i < n_units) queue[2 + n_groups + i] = (i < n_groups) ? i : -1;
}

// shared memory of a team of the HMC kernel: exchange buffers (x | x - mu | x^2) | reduction buffers | control words
template <int NR, int MV>
__host__ __device__ constexpr int team_base_doubles() { return 3 * TeamShape<NR, MV>::SLOT + TeamShape<NR, MV>::RED_DOUBLES + 24; }

// next work unit of the team (leader pops, everybody reads after the barrier); false when the queue is exhausted
__device__ __forceinline__ bool team_next_unit(int *queue, volatile int *ring, int n_units, bool &first, int slot, volatile int *tctl,
                                               int bar_id, bool leader, int &group)
{
    if (leader) {
        int idx = first ? slot : atomicAdd(queue, 1);
        int grp = -1;
        if (idx < n_units) { while ((grp = ring[idx]) < 0) __nanosleep(100); }
        tctl[0] = grp;
    }
    first = false;
    team_bar(bar_id);
    group = tctl[0];
    team_bar(bar_id);              // everybody has read the word before the leader may write the next one
    __threadfence();
    return group >= 0;
}

// ----------------------------------------------------------------------------------------------------------------------
// HMC (samplers/hmc.py:16-49): n_int_step leapfrogs per iteration for every chain, no tree.
// ----------------------------------------------------------------------------------------------------------------------
template <int NR, int MV, int G>
__global__ void __launch_bounds__(128 * G, 1) hmc_team_kernel(DevModel M, bfb_sampler_cfg cfg, ChainState st, RunOutDevF out,
                                                              int base_iter, int chunk_iters, int n_groups, int n_units,
                                                              int *__restrict__ queue)
{
    using TS = TeamShape<NR, MV>;
    constexpr int NRW = TS::NRW;
    extern __shared__ double smem[];
    double *tab = smem, *msm = smem + TS::TAB_DOUBLES;
    for (int i = threadIdx.x; i < TS::TAB_DOUBLES; i += blockDim.x) tab[i] = M.tfrag[i];
    if (threadIdx.x < 32) { msm[threadIdx.x] = M.use_bound ? M.mu[threadIdx.x] : 0.; msm[32 + threadIdx.x] = M.lin[threadIdx.x]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, team = wib >> 2, w = wib & 3, gi = lane >> 2, lg = lane & 3;
    double *tsm = smem + TS::TAB_DOUBLES + 64 + (size_t)team * team_base_doubles<NR, MV>();
    double *xbuf = tsm, *red = tsm + 3 * TS::SLOT;
    const double *mu_t = msm;
    volatile int *tctl = reinterpret_cast<volatile int *>(red + TS::RED_DOUBLES);
    double *bcast = red + TS::RED_DOUBLES + 8;         // [2][8 chains]: value and energy of the end point, from the leader warp
    const int bar_id = 1 + team;
    const double *tab_w = tab + (size_t)w * NR * TS::NTW * 32;
    const bool leader = (w == 0 && lane == 0), scribe = (w == 0 && lg == 0);
    const int n = M.n;
    const DmmaConsts K = dmma_consts(M);
    volatile int *qv = queue;
    volatile int *ring = queue + 2 + n_groups;
    int rbuf = 0;
    bool first = true;
    const int slot = blockIdx.x + gridDim.x * team;
#pragma unroll 1
    for (;;) {
    int group;
    if (!team_next_unit(queue, ring, n_units, first, slot, tctl, bar_id, leader, group)) break;
    const int chunk = qv[2 + group];
    const int it_lo = chunk * chunk_iters;
    const int it_hi = min(out.n_iter, it_lo + chunk_iters);
    const int64_t c_raw = (int64_t)group * 8 + gi;
    const bool exists = c_raw < st.C;
    const int64_t c = exists ? c_raw : st.C - 1;
    const size_t vb = (size_t)c * M.np;
    double q[NRW], p[NRW], g[NRW], var[NRW];
#pragma unroll
    for (int i = 0; i < NRW; ++i) {
        const int j = 4 * (NRW * w + i) + lg;
        q[i] = st.q[vb + j]; g[i] = st.g[vb + j]; var[i] = st.var[vb + j]; p[i] = 0.;
    }
    const uint64_t seed = cfg.seed, chain_id = (uint64_t)(cfg.chain0 + c);
    int64_t t = st.t_draw[c];
    const int it0 = base_iter;
    double logp_q = st.logp[c];
    double log_step = st.log_step[c], log_bar = st.log_bar[c], hbar = st.hbar[c];
    const double mu_da = st.mu_da[c];
    int64_t count = st.count[c], n_samples = st.n_samples[c], previous_update = st.previous_update[c];
    int adapt_window = st.adapt_window[c];
    double fg_n = st.fg_n[c], bg_n = st.bg_n[c];
    int status = exists ? st.status[c] : 9;
    unsigned long long tree_total = 0;

#pragma unroll 1
    for (int it = it_lo; it < it_hi; ++it) {
        const bool warm = (it0 + it) < cfg.n_warmup;
        // momentum: metrics.py:83-86 (every lane draws the normals of its own dimensions)
        double qs[NRW], gs[NRW], part = 0.;
#pragma unroll
        for (int i = 0; i < NRW; ++i) {
            const int j = 4 * (NRW * w + i) + lg;
            const double z = (j < n) ? team_draw_normal(seed, chain_id, (uint64_t)(t + j)) : 0.;
            p[i] = (j < n) ? z / sqrt(var[i]) : 0.;
            part = fma(p[i], var[i] * p[i], part);
            qs[i] = q[i]; gs[i] = g[i];
        }
        double z0 = 0., z1 = 0., z2 = 0.;
        team_sum4(part, z0, z1, z2, red, rbuf, bar_id, lane, w);
        const double ke0 = part;
        bool live = status == 0;
        double E0 = 0.5 * ke0 - logp_q;
        if (live) {
            t += n;
            if (!isfinite(E0)) { status = 2; live = false; }                 // base_hmc.py:72-76
        }
        const double eps = warm ? exp(log_step) : exp(log_bar);               // step_size.py:25-29
        const double dt = 0.5 * eps;
        double lp = logp_q, E = E0;
#pragma unroll 1
        for (int s_ = 0; s_ < cfg.n_int_step; ++s_) {
            // integration.py:68-95
#pragma unroll
            for (int i = 0; i < NRW; ++i) {
                p[i] = fma(dt, g[i], p[i]);
                q[i] = fma(eps, var[i] * p[i], q[i]);
                const int e = (NRW * w + i) * 32 + lane;
                xbuf[e] = q[i]; xbuf[TS::SLOT + e] = q[i] - mu_t[4 * (NRW * w + i) + lg]; xbuf[2 * TS::SLOT + e] = q[i] * q[i];
            }
            team_bar(bar_id);
            double gn[NRW], ke2;
            team_logp_grad<NR, MV>(tab_w, msm, xbuf, red, rbuf, bar_id, lane, w, K, live, q, p, var, dt, lp, gn, ke2);
#pragma unroll
            for (int i = 0; i < NRW; ++i) { g[i] = gn[i]; p[i] = fma(dt, gn[i], p[i]); }
            E = 0.5 * ke2 - lp;
        }
        // the value and the kinetic energy are complete in the leader warp only (team_logp_grad)
        if (cfg.n_int_step > 0) {
            if (scribe) { bcast[gi] = lp; bcast[8 + gi] = E; }
            team_bar(bar_id);
            lp = bcast[gi]; E = bcast[8 + gi];
        }
        // HMC._hamiltonian_step, hmc.py:31-49
        double dE;
        int diverging = 0;
        if (isfinite(E)) { dE = E0 - E; diverging = fabs(dE) > cfg.max_change; }
        else { dE = -INFINITY; diverging = 1; }
        double accept_stat;
        { const double e_ = exp(dE); accept_stat = e_ < 1. ? e_ : 1.; }
        const double ua = team_draw_uniform(seed, chain_id, (uint64_t)t);
        bool accepted = false;
        if (live && !diverging) { t += 1; accepted = !(ua >= accept_stat); }
        const double s_logp = lp, s_energy = E;
        if (accepted) logp_q = lp;
        else {
#pragma unroll
            for (int i = 0; i < NRW; ++i) { q[i] = qs[i]; g[i] = gs[i]; }
        }
        if (live) {
            tree_total += (unsigned long long)cfg.n_int_step;
            if (warm && cfg.adapt_step_size) {          // step_size.py:31-45
                const double cnt = (double)count;
                const double ww = 1. / (cnt + cfg.t0);
                hbar = ((1. - ww) * hbar + ww * (cfg.target_accept - accept_stat));
                log_step = mu_da - hbar * sqrt(cnt) / cfg.gamma;
                const double mk = pow(cnt, -cfg.k);
                log_bar = mk * log_step + (1. - mk) * log_bar;
                count += 1;
            }
            if (warm && cfg.adapt_metric) {             // metrics.py:186-211, 351-357; Welford state in global memory
                const int64_t delta = n_samples - previous_update;
                const bool upd = ((delta + 1) % cfg.update_window == 0);
                const bool swap = delta >= adapt_window;
                fg_n += 1.; bg_n += 1.;
#pragma unroll
                for (int i = 0; i < NRW; ++i) {
                    const int j = 4 * (NRW * w + i) + lg;
                    if (j < n) {
                        double fgm = st.fg_mean[vb + j], fgr = st.fg_raw[vb + j], bgm = st.bg_mean[vb + j], bgr = st.bg_raw[vb + j];
                        double od = q[i] - fgm;
                        fgm += od / fg_n;
                        fgr += 1. * od * (q[i] - fgm);
                        od = q[i] - bgm;
                        bgm += od / bg_n;
                        bgr += 1. * od * (q[i] - bgm);
                        if (upd) var[i] = fgr / fg_n;
                        if (swap) { fgm = bgm; fgr = bgr; bgm = 0.; bgr = 0.; }
                        st.fg_mean[vb + j] = fgm; st.fg_raw[vb + j] = fgr; st.bg_mean[vb + j] = bgm; st.bg_raw[vb + j] = bgr;
                    }
                }
                if (swap) { fg_n = bg_n; bg_n = 10.; previous_update = n_samples; if (cfg.doubling) adapt_window *= 2; }
                n_samples += 1;
            }
            const size_t o = (size_t)c * out.n_iter + it;
            if (out.o.samples) {
#pragma unroll
                for (int i = 0; i < NRW; ++i) { const int j = 4 * (NRW * w + i) + lg; if (j < n) out.o.samples[o * n + j] = q[i]; }
            }
            if (scribe) {
                if (out.o.logp) out.o.logp[o] = s_logp;
                if (out.o.energy) out.o.energy[o] = s_energy;
                if (out.o.tree_depth) out.o.tree_depth[o] = accepted ? 1 : 0;
                if (out.o.tree_size) out.o.tree_size[o] = cfg.n_int_step;
                if (out.o.mean_tree_accept) out.o.mean_tree_accept[o] = accept_stat;
                if (out.o.step_size) out.o.step_size[o] = exp(log_step);
                if (out.o.step_size_bar) out.o.step_size_bar[o] = exp(log_bar);
                if (out.o.energy_change) out.o.energy_change[o] = dE;
                if (out.o.max_energy_change) out.o.max_energy_change[o] = 0.;
                if (out.o.diverging) out.o.diverging[o] = diverging;
            }
        }
    }
    // ---- persist chain state ----
    if (exists && st.status[c] == 0) {
#pragma unroll
        for (int i = 0; i < NRW; ++i) {
            const int j = 4 * (NRW * w + i) + lg;
            st.q[vb + j] = q[i]; st.g[vb + j] = g[i]; st.var[vb + j] = var[i];
        }
    }
    team_bar(bar_id);                // every warp has read st.status before the scribe overwrites it
    if (exists && scribe && st.status[c] == 0) {
        st.logp[c] = logp_q; st.fg_n[c] = fg_n; st.bg_n[c] = bg_n;
        st.log_step[c] = log_step; st.log_bar[c] = log_bar; st.hbar[c] = hbar;
        st.count[c] = count; st.n_samples[c] = n_samples; st.previous_update[c] = previous_update;
        st.adapt_window[c] = adapt_window; st.t_draw[c] = t; st.iter[c] = it0 + it_hi;
        st.status[c] = status;
        if (tree_total) atomicAdd(st.tree_total, tree_total);
    }
    __threadfence();
    team_bar(bar_id);
    if (leader) {
        qv[2 + group] = chunk + 1;
        if ((chunk + 1) * chunk_iters < out.n_iter) {
            const int ti = atomicAdd(queue + 1, 1);
            __threadfence();
            ring[ti] = group;
        }
    }
    }   // unit loop
}

static int team_queue_setup(bfb_context *h, int n_groups, int n_iter, int slots, int &chunk_iters, int &n_units)
{
    chunk_iters = (n_iter + 5) / 6;
    if (chunk_iters < 16) chunk_iters = n_iter < 16 ? n_iter : 16;
    if (const char *e = getenv("BFB200_CHUNK_ITERS")) { int v = atoi(e); if (v >= 1) chunk_iters = v; }
    const int n_chunks = (n_iter + chunk_iters - 1) / chunk_iters;
    const int64_t n_units64 = (int64_t)n_groups * n_chunks;
    BFB_REQUIRE(n_units64 < (1ll << 31), BFB_ERR_ARG, "too many work units");
    n_units = (int)n_units64;
    const size_t qlen = 2 + (size_t)n_groups + (size_t)n_units64;
    if (qlen > h->queue_len) {
        if (h->queue) cudaFree(h->queue);
        h->queue = nullptr; h->queue_len = 0;
        BFB_CUDA(cudaMalloc((void **)&h->queue, sizeof(int) * qlen));
        h->queue_len = qlen;
    }
    team_queue_init_kernel<<<(unsigned)((n_units64 + 255) / 256), 256, 0, h->stream>>>(h->queue, n_groups, n_units, slots);
    h->launches++;
    return BFB_OK;
}

template <int NR, int MV, int G>
static int launch_hmc_team(bfb_context *h, const bfb_run_out &o, int n_iter)
{
    using TS = TeamShape<NR, MV>;
    const int64_t C = h->cs.C;
    const int n_groups = (int)((C + 7) / 8);
    const size_t smem = sizeof(double) * (TS::TAB_DOUBLES + 64 + (size_t)G * team_base_doubles<NR, MV>());
    BFB_CUDA(cudaFuncSetAttribute(hmc_team_kernel<NR, MV, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RunOutDevF od;
    od.o = o; od.n_iter = n_iter;
    int blocks = h->sm_count;
    if ((int64_t)blocks > n_groups) blocks = n_groups;
    int chunk_iters, n_units, rc;
    if ((rc = team_queue_setup(h, n_groups, n_iter, blocks * G, chunk_iters, n_units))) return rc;
    hmc_team_kernel<NR, MV, G><<<blocks, 128 * G, smem, h->stream>>>(h->dm, h->scfg, h->cs, od, (int)h->iters_done, chunk_iters,
                                                                    n_groups, n_units, h->queue);
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

template <int NR, int MV>
static int launch_hmc_team_g(bfb_context *h, const bfb_run_out &o, int n_iter)
{
    int G = 4;
    if (const char *e = getenv("BFB200_TEAMS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 6) G = v; }
    switch (G) {
    case 1: return launch_hmc_team<NR, MV, 1>(h, o, n_iter);
    case 2: return launch_hmc_team<NR, MV, 2>(h, o, n_iter);
    case 3: return launch_hmc_team<NR, MV, 3>(h, o, n_iter);
    case 5: return launch_hmc_team<NR, MV, 5>(h, o, n_iter);
    case 6: return launch_hmc_team<NR, MV, 6>(h, o, n_iter);
    }
    return launch_hmc_team<NR, MV, 4>(h, o, n_iter);
}

// returns 1 if this path does not apply (caller tries the next kernel), 0 on launch, <0 on error
int bfb_launch_hmc_team(bfb_context *h, const bfb_run_out &o, int n_iter)
{
    const DevModel &M = h->dm;
    if (M.epilogue || !M.tfrag || M.frag_nr == 0 || M.frag_ext || M.has_c3) return 1;
    if (const char *e = getenv("BFB200_SAMPLER")) { if (strcmp(e, "team")) return 1; }
    const int mv = M.has_c2 ? 1 : 0;
#define BFB_CASE(NR_, MV_) if (M.frag_nr == NR_ && mv == MV_) return launch_hmc_team_g<NR_, MV_>(h, o, n_iter);
    BFB_CASE(4, 0) BFB_CASE(4, 1) BFB_CASE(7, 0) BFB_CASE(7, 1) BFB_CASE(8, 0) BFB_CASE(8, 1)
#undef BFB_CASE
    return 1;
}
