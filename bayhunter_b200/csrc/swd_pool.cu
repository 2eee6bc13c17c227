// swd_pool.cu -- the dispersion search with the lanes of a whole CTA as one pool (full batches).
//
// swd_kernel gives a warp S models of one curve and deals the warp's 32 lanes to their pending candidates.
// Here a CTA of 4 warps owns all chains of M models of ONE wave type -- the first-root and second-root chains
// of the group curve and the chain of the phase curve, 3 M <= 128 chains -- and every round deals ALL its 128
// lanes to ALL pending candidates: the same search state machine (swd_core.cuh), the same dealing rule (a
// refining chain one lane, the spare lanes evenly to the walking chains) and exactly the reference's candidate
// sequence, so the results are bit-identical to swd_kernel's.  What changes:
//   * the chains sit on the first 3 M threads, so the divergent bookkeeping (publish, consume) runs in 2-3 of
//     the 4 warps while the evaluation -- the convergent, fp64-bound part -- is packed densely over all four;
//   * group chains (32 chains on 32 lanes in swd_kernel: no lane to spare) borrow the lanes phase chains leave;
//   * a model's layer records are built once for its three chains;
//   * 125 registers (Rayleigh), M = 28: 4 CTAs = 16 warps per SM, 586 CTAs on 592 slots at B = 8192.
// Measured (joint5): B = 8192 3.64 -> 3.45 ms per evaluation (M = 28), B = 16384 7.14 -> 6.47 (M = 32), B = 4096
// 2.55 -> 2.42 (M = 14: the kGuess instantiation, whose spare lanes go to the refining chains' next midpoints,
// swd_core.cuh: refine_guess); below ~7 k (model, wave type) pairs the CTAs no longer fill the device and
// swd_kernel with its fitted layout is faster, so the engine picks this kernel by rule (engine.cu) --
// profiles/r02_swd_restructure.txt sections 12 and 15.
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace bh {

namespace {

#ifndef BH_POOL_WARPS
#define BH_POOL_WARPS 4
#endif
#ifndef BH_POOL_MIN_BLOCKS
#define BH_POOL_MIN_BLOCKS 4
#endif
#ifndef BH_POOL_GUESS_MAX_CHAINS
#define BH_POOL_GUESS_MAX_CHAINS 72
#endif
constexpr int kPoolWarps = BH_POOL_WARPS;
constexpr int kPoolLanes = kPoolWarps * 32;
constexpr int kPoolMaxModels = kPoolLanes;      // links / first-root mailboxes per CTA

struct PoolShared {
  double c[kPoolLanes], clow[kPoolLanes], c2[kPoolLanes], omega[kPoolLanes], del[kPoolLanes];
  double omA[SWD_MAX_PERIODS], omB[SWD_MAX_PERIODS], omP[SWD_MAX_PERIODS];
  SearchLink link[kPoolMaxModels];
  int stage[kPoolLanes], idir[kPoolLanes], nlay[kPoolLanes], col[kPoolLanes], owner_at[kPoolLanes];
  unsigned act[kPoolWarps], brk[kPoolWarps];       // per warp: lanes that want a value / more than one
  unsigned startbits[kPoolWarps];                  // bit g set: a chain's run of lanes starts at pool lane g
  int any_wait;
  double tab[SWD_TAB_ROWS * kPoolLanes];           // Neville tableaus, one column per lane
};

// kWave: 1 Love, 2 Rayleigh (all curves of a launch are of one wave type)
// kGuess: refining chains take lanes for refinement guesses when the CTA has lanes to spare (M well below 28)
template <int kWave, bool kGuess>
__global__ void __launch_bounds__(kPoolLanes, BH_POOL_MIN_BLOCKS)
swd_pool_kernel(SwdLaunch p, int M) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wave = kWave;
  const int b0 = (int)blockIdx.x * M;
  const int nmod = min(M, p.B - b0);
  // the launch's curves: group (gc) and / or phase (pc)
  int gc = -1, pc = -1;
  for (int c = 0; c < p.ncurves; ++c) { if (p.igr[c]) gc = c; else pc = c; }
  const int stride = p.row_stride, lcap = p.lcap;
  const int* __restrict__ perm = p.perm;
#define BH_MODEL(j) (perm ? perm[b0 + (j)] : b0 + (j))

  double* rec = reinterpret_cast<double*>(smem_raw);       // [field][layer][M]
  const int fs = lcap * M;
  PoolShared* ws = reinterpret_cast<PoolShared*>(rec + (size_t)SWD_REC_FIELDS * fs);

  if (gc >= 0) for (int k = t; k < p.kmax[gc]; k += kPoolLanes) swd_period_omegas(1, p.periods[gc][k], &ws->omA[k], &ws->omB[k]);
  if (pc >= 0) for (int k = t; k < p.kmax[pc]; k += kPoolLanes) { double u; swd_period_omegas(0, p.periods[pc][k], &ws->omP[k], &u); }
  if (t < kPoolMaxModels) { ws->link[t].na = 0; ws->link[t].a_failed = 0; ws->link[t].del1st = 0.0; }
  if (t < kPoolWarps) ws->startbits[t] = 0u;
  if (t == 0) ws->any_wait = 0;

  // chain layout: [0, M) first roots of the group curve, [M, 2M) its second roots, then M phase chains
  const int ngrp = gc >= 0 ? 2 * M : 0;
  const bool rider = t >= ngrp;                            // a chain of the phase curve
  const int mycurve = rider ? pc : gc;
  const int role = (!rider && t >= M) ? 1 : 0;
  const int sidx = rider ? t - ngrp : (role ? t - M : t);
  bool owner = mycurve >= 0 && sidx < nmod && t < ngrp + (pc >= 0 ? M : 0);
  const int mymodel = owner ? BH_MODEL(sidx) : 0;
  if (owner) {
    const int n = p.nlay[mymodel];
    owner = n > p.nlay_lo && n <= p.nlay_hi;
  }
  const int igr = owner ? p.igr[mycurve] : 0, kmax = owner ? p.kmax[mycurve] : 1;
  const double* __restrict__ periods = owner ? p.periods[mycurve] : nullptr;
  const int tid = owner ? p.target_id[mycurve] : 0;
  if (!__syncthreads_or((int)owner)) {
    if (t == 0 && p.done) atomicAdd(p.done, kPoolWarps);
    return;
  }
  Search s;
  SearchCtx ctx;
  ctx.omA = rider ? ws->omP : ws->omA; ctx.omB = ws->omB;
  ctx.link = (igr > 0 && owner) ? &ws->link[sidx] : nullptr;
  {
    double* r = p.roots + ((size_t)mymodel * p.curve_stride + (owner ? p.curve_off[mycurve] : 0)) * 2;
    ctx.ra = r; ctx.rb = r + kmax;
  }
  int myL = 0;
  if (owner) {
    myL = min(p.nlay[mymodel], lcap);
    if (search_setup(s, p.rows + (size_t)mymodel * stride, 1, myL, kmax, role, ws->tab + t, kPoolLanes)) {
      if (role == 0) search_begin_a(s, ctx);
    } else if (role == 0 && ctx.link) {
      ctx.link->a_failed = 1;
    }
  } else {
    s.stage = ST_DONE;
  }
  ws->nlay[t] = myL;
  ws->col[t] = sidx;
  // ---- fp64 layer records of the CTA's models (one set serves all chains of a model) ----
  for (int e = t; e < lcap * M; e += kPoolLanes) {
    const int m = e % M, l = e / M;
    if (m < nmod) {
      const int n = p.nlay[BH_MODEL(m)];
      const int L = (n > p.nlay_lo && n <= p.nlay_hi) ? min(n, lcap) : 0;
      if (l < L) swd_make_rec(wave, p.rows[(size_t)BH_MODEL(m) * stride + l], l == L - 1, rec + (size_t)l * M + m, fs);
    }
  }
  __syncthreads();

  unsigned long long consumed = 0, evaluated = 0;
  unsigned rounds = 0;
  const int max_spec = p.max_spec;
  const double dc = fabs((double)0.005f);
  const unsigned below = (1u << lane) - 1u;

  // (A variant with three barriers per round -- the next round's ballots taken at the end of the consume step -- was
  //  slower: 3.51 vs 3.45 ms, 128 vs 125 registers.)
#ifdef BH_SWD_TIMING
  long long cyc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define BH_TICK(i) { const long long now_ = clock64(); cyc[i] += now_ - tick_; tick_ = now_; }
  long long tick_ = clock64();
#else
#define BH_TICK(i)
#endif
  for (;;) {
    // ---- phase A: owners say what they want, the CTA deals its lanes ----
    if (role) search_poll_b(s, ctx);
    const int want = search_nwant(s, kPoolLanes);
    const unsigned a_w = __ballot_sync(0xffffffffu, want > 0), b_w = __ballot_sync(0xffffffffu, want > 1);
    if (lane == 0) { ws->act[warp] = a_w; ws->brk[warp] = b_w; }
    if (s.stage == ST_WAIT) ws->any_wait = 1;      // a second-root chain whose first root may arrive this round
    BH_TICK(0)
    __syncthreads();
    BH_TICK(1)
    int nact = 0, nbr = 0, act_below = 0, rank = 0;
#pragma unroll
    for (int w = 0; w < kPoolWarps; ++w) {
      const unsigned a = ws->act[w], b = ws->brk[w];
      nact += __popc(a); nbr += __popc(b);
      if (w < warp) { act_below += __popc(a); rank += __popc(b); }
    }
    const bool waiting = ws->any_wait != 0;
    if (nact == 0) {
      if (!waiting) break;                         // uniform: everything read from shared memory after the barrier
      __syncthreads();
      if (t == 0) ws->any_wait = 0;
      __syncthreads();
      continue;
    }
    ++rounds;
    act_below += __popc(a_w & below);
    rank += __popc(b_w & below);
    // dealing rule of swd_core.cuh (deal_lanes), over the CTA's lanes
    const int nrf = nact - nbr;
    const int g = kGuess ? refine_guess_lanes(nrf, kPoolLanes - nact - BH_GUESS_WALK_EXTRA * nbr) : 0;
    const int extra = kPoolLanes - nact - g * nrf;
    const int quo = nbr ? __float2int_rz(__fdividef((float)extra + 0.5f, (float)nbr)) : 0;
    const int rem = extra - quo * nbr;
    const int per = quo + 1;
    const bool capped = per >= max_spec;
    int cnt = 0;
    if (want > 0) {
      cnt = 1 + g;
      if (want > 1) { cnt = per + (rank < rem ? 1 : 0); if (cnt > max_spec) cnt = max_spec; }
    }
    const int excl = (act_below - rank) * (1 + g) + (capped ? rank * max_spec : rank * per + (rank < rem ? rank : rem));
    const int total = nrf * (1 + g) + (capped ? nbr * max_spec : nbr * per + (nbr < rem ? nbr : rem));
    if (cnt > 0) {
      ws->c[t] = search_pending_c(s);
      ws->clow[t] = (kGuess && s.stage > ST_BR_STEP) ? s.c1 : s.clow;
      if (kGuess) ws->c2[t] = s.c2;
      ws->omega[t] = s.omega;
      ws->stage[t] = s.stage;
      ws->idir[t] = s.idir;
      ws->owner_at[excl] = t;
      atomicOr(&ws->startbits[excl >> 5], 1u << (excl & 31));
    }
    BH_TICK(2)
    __syncthreads();
    BH_TICK(3)

    // ---- phase B: every dealt lane evaluates one candidate ----
    if (t < total) {
      int w = warp;
      unsigned m = ws->startbits[w] & (0xffffffffu >> (31 - lane));
      while (m == 0u) { --w; m = ws->startbits[w]; }
      const int start = w * 32 + 31 - __clz(m);
      const int i = t - start;
      const int own = ws->owner_at[start];
      const double omega = ws->omega[own];
      const double c = candidate_from(ws->stage[own], ws->c[own], ws->idir[own], ws->clow[own], dc, i, kGuess ? ws->c2[own] : 0.0, true);
      ws->del[t] = secular_rec(wave, rec + ws->col[own], fs, M, ws->nlay[own], fm::div(omega, c), omega);
      evaluated += 1;
    }
    BH_TICK(4)
    __syncthreads();
    BH_TICK(5)

    // ---- phase C: owners consume their values in reference order ----
    if (t < kPoolWarps) ws->startbits[t] = 0u;
    if (t == 0) ws->any_wait = 0;
    if (cnt > 0) consumed += search_consume(s, &ws->del[excl], cnt, ctx, kGuess, true);
    BH_TICK(6)
    __syncthreads();
    BH_TICK(7)
  }
#ifdef BH_SWD_TIMING
  if (lane == 0 && (blockIdx.x % 59) == 7)
    printf("pool wave %d cta %d warp %d rounds %u: poll+ballot %lld | wait %lld | deal+publish %lld | wait %lld | evaluate %lld | wait %lld | consume %lld | wait %lld cycles per round\n",
           wave, (int)blockIdx.x, warp, rounds, cyc[0] / max(rounds, 1u), cyc[1] / max(rounds, 1u), cyc[2] / max(rounds, 1u), cyc[3] / max(rounds, 1u),
           cyc[4] / max(rounds, 1u), cyc[5] / max(rounds, 1u), cyc[6] / max(rounds, 1u), cyc[7] / max(rounds, 1u));
#endif

  // ---- curve values from the stored roots; validity flag ----
  __shared__ int done_flag[kPoolLanes];
  done_flag[t] = s.stage == ST_DONE;
  __syncthreads();
  if (owner && role == 0) {
    bool ok = done_flag[t] != 0;
    if (igr > 0) ok = ok && done_flag[t + M] != 0;
    double* __restrict__ my_curve = p.curves + (size_t)mymodel * p.curve_stride + p.curve_off[mycurve];
    if (ok)
      for (int k = 0; k < kmax; ++k)
        my_curve[k] = swd_curve_value(igr, periods[k], ctx.ra[k], igr > 0 ? ctx.rb[k] : 0.0);
    p.tstatus[(size_t)mymodel * kMaxTargets + tid] = ok ? 1 : 0;
  }
#undef BH_MODEL
  for (int d = 16; d > 0; d >>= 1) {
    evaluated += __shfl_down_sync(0xffffffffu, evaluated, d);
    consumed += __shfl_down_sync(0xffffffffu, consumed, d);
  }
  if (lane == 0 && p.done) atomicAdd(p.done, 1);
  if (lane == 0 && p.counters) {
    atomicAdd(&p.counters[0], consumed);
    atomicAdd(&p.counters[1], evaluated);
    if (warp == 0) {
      const int slot = gc >= 0 ? gc : pc;
      atomicAdd(&p.counters[2 + 2 * (p.counter_base + slot)], (unsigned long long)rounds);
      atomicMax(&p.counters[3 + 2 * (p.counter_base + slot)], (unsigned long long)rounds);
    }
  }
}

size_t pool_smem_bytes(int lcap, int M) {
  return (size_t)SWD_REC_FIELDS * lcap * M * sizeof(double) + sizeof(PoolShared);
}

template <int kWave, bool kGuess>
void launch_pool(const SwdLaunch& p, int M, cudaStream_t st) {
  const int nb = (p.B + M - 1) / M;
  const size_t smem = pool_smem_bytes(p.lcap, M);
  static KernelAttrs attrs;
  bh_configure_kernel(swd_pool_kernel<kWave, kGuess>, smem, attrs);
  static bool reported = false;
  if (!reported && getenv("BH_DEBUG")) {
    reported = true;
    int res = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, swd_pool_kernel<kWave, kGuess>, kPoolLanes, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, swd_pool_kernel<kWave, kGuess>);
    fprintf(stderr, "[bh] swd_pool_kernel<%d,%d>: %d CTAs of %d lanes, %d models each, smem %zu B, regs %d, local %zu B, max resident CTAs/SM %d\n",
            kWave, (int)kGuess, nb, kPoolLanes, M, smem, fa.numRegs, fa.localSizeBytes, res);
  }
  swd_pool_kernel<kWave, kGuess><<<nb, kPoolLanes, smem, st>>>(p, M);
}

}  // namespace

// The pool kernel takes a launch with at most one group and one phase curve, all of one wave type.
bool swd_pool_fits(const SwdLaunch& p) {
  int ng = 0, np = 0;
  for (int c = 0; c < p.ncurves; ++c) {
    if (p.wave[c] != p.wave[0]) return false;
    (p.igr[c] ? ng : np) += 1;
  }
  return p.ncurves > 0 && ng <= 1 && np <= 1;
}

// Models per CTA: about 7/8 of the pool's lanes own a chain (3 chains per model with a group and a phase curve),
// the rest speculate for the walking chains; `want` > 0 overrides.
int swd_pool_models(const SwdLaunch& p, int lcap, int want) {
  int cpm = 0;
  for (int c = 0; c < p.ncurves; ++c) cpm += p.igr[c] ? 2 : 1;
  if (cpm < 1) cpm = 1;
  int m = want > 0 ? want : (kPoolLanes * 7 / 8) / cpm;
  if (m * cpm > kPoolLanes) m = kPoolLanes / cpm;
  while (m > 1 && pool_smem_bytes(lcap, m) > 72 * 1024) --m;
  return m < 1 ? 1 : m;
}

size_t swd_pool_smem_bytes(int lcap, int M) { return pool_smem_bytes(lcap, M); }

int swd_pool_warp_count(const SwdLaunch& p, int M) { return kPoolWarps * ((p.B + M - 1) / M); }

void launch_swd_pool(const SwdLaunch& p, int M, cudaStream_t st) {
  if (p.ncurves <= 0 || p.B <= 0) return;
  // refinement guesses need 3 lanes per refining + 4 per walking chain: no room from ~24 models (72 chains) up
  int cpm = 0;
  for (int c = 0; c < p.ncurves; ++c) cpm += p.igr[c] ? 2 : 1;
  const bool guess = M * cpm <= BH_POOL_GUESS_MAX_CHAINS;
  if (p.wave[0] == 2) { if (guess) launch_pool<2, true>(p, M, st); else launch_pool<2, false>(p, M, st); }
  else { if (guess) launch_pool<1, true>(p, M, st); else launch_pool<1, false>(p, M, st); }
}

}  // namespace bh
