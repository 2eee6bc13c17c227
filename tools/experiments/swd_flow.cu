// swd_flow.cu -- the dispersion search without rounds: chain states in shared memory, warps as generic workers.
//
// swd_kernel and swd_pool_kernel advance their chains in rounds -- deal lanes, evaluate, consume, each behind a
// barrier -- so the fp64 pipe idles while the (divergent, latency-bound) bookkeeping of a round runs, and a
// round lasts as long as its slowest warp.  Here a CTA of 4 warps owns the chains of M models of ONE wave type
// as before (first-root and second-root chains of the group curve, the chain of the phase curve; 3 M <= 96),
// but a chain's state machine lives in shared memory and work moves through three queues:
//     evalQ   pending secular-function candidates (chain, index in the chain's published run)
//     walkQ   chains whose published candidates are all evaluated, walking a bracket (+ woken second-root chains)
//     refQ    the same, refining a root (nevill)
// Every warp loops: take up to 32 ready chains of ONE kind and consume their values (the warp runs one branch of
// the state machine, not all of them), else take up to 32 candidates and evaluate them, else nap.  A consumed
// chain publishes its next candidates at once, so no chain ever waits for another chain's bookkeeping.
// The candidate sequence each chain consumes is the reference's (swd_core.cuh), so the results are bit-identical
// to swd_kernel's; only the number of speculative evaluations depends on timing.
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace bh {

namespace {

constexpr int kFlowWarps = 4;
constexpr int kFlowThreads = kFlowWarps * 32;
constexpr int kFlowChains = 96;          // chain slots per CTA
constexpr int kFlowModels = 32;          // model columns per CTA (3 per model: <= 32)
constexpr int kFlowSlots = 4;            // result slots per chain (candidates in flight)
constexpr int kRingE = 512, kRingC = 128;
constexpr int kGenShift = 10;            // ring entry = generation << 10 | payload

struct FlowShared {
  Search st[kFlowChains];
  double res[kFlowChains * kFlowSlots];
  double omA[SWD_MAX_PERIODS], omB[SWD_MAX_PERIODS], omP[SWD_MAX_PERIODS];
  double tab[SWD_TAB_ROWS * kFlowChains];          // Neville tableaus, one column per chain
  SearchLink link[kFlowModels];
  int waitflag[kFlowModels];                       // 1: the model's second-root chain is parked until the next first root
  int outstanding[kFlowChains], pubn[kFlowChains];
  int nlay[kFlowModels], model[kFlowModels];
  int ringE[kRingE], ringW[kRingC], ringR[kRingC];
  unsigned headE, tailE, headW, tailW, headR, tailR;
  int live;                                        // chains that are not finished
  int error;
};

__device__ __forceinline__ unsigned ldv(const unsigned* p) { return *(const volatile unsigned*)p; }
__device__ __forceinline__ int ldv(const int* p) { return *(const volatile int*)p; }

// All 32 lanes call; a lane appends n entries payload0, payload0 + 1, ...
template <int R>
__device__ __forceinline__ void ring_push(int* buf, unsigned* tail, int n, int payload0, int lane) {
  int incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(tail, (unsigned)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  const unsigned pos = base + (unsigned)(incl - n);
  for (int i = 0; i < n; ++i) {
    const unsigned q = pos + i;
    *(volatile int*)&buf[q & (R - 1)] = (int)((((q / R) & 0xfffffu) << kGenShift) | (unsigned)(payload0 + i));
  }
}

// All 32 lanes call; returns how many entries the warp took (lane i < n holds one in *payload, the others -1).
template <int R>
__device__ __forceinline__ int ring_pop(int* buf, unsigned* head, const unsigned* tail, int lane, int* payload) {
  unsigned h = 0;
  int n = 0;
  if (lane == 0) {
    for (;;) {
      h = ldv(head);
      const int avail = (int)(ldv(tail) - h);
      if (avail <= 0) { n = 0; break; }
      n = avail < 32 ? avail : 32;
      if (atomicCAS(head, h, h + (unsigned)n) == h) break;
    }
  }
  n = __shfl_sync(0xffffffffu, n, 0);
  h = __shfl_sync(0xffffffffu, h, 0);
  *payload = -1;
  if (lane < n) {
    const unsigned q = h + lane;
    const int gen = (int)((q / R) & 0xfffffu);
    int v, spins = 0;
    for (;;) {                                       // reserved by its producer, written a few instructions later
      v = ldv(&buf[q & (R - 1)]);
      if ((int)((unsigned)v >> kGenShift) == gen) break;
      if (++spins > (1 << 22)) { v = -1; break; }
    }
    *payload = v < 0 ? -1 : (v & ((1 << kGenShift) - 1));
  }
  return n;
}

// kWave: 1 Love, 2 Rayleigh (all curves of a launch are of one wave type)
template <int kWave>
__global__ void __launch_bounds__(kFlowThreads, 4)
swd_flow_kernel(SwdLaunch p, int M, int spec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wave = kWave;
  const int b0 = (int)blockIdx.x * M;
  const int nmod = min(M, p.B - b0);
  int gc = -1, pc = -1;
  for (int c = 0; c < p.ncurves; ++c) { if (p.igr[c]) gc = c; else pc = c; }
  const int stride = p.row_stride, lcap = p.lcap;
  const int* __restrict__ perm = p.perm;
#define BH_MODEL(j) (perm ? perm[b0 + (j)] : b0 + (j))

  double* rec = reinterpret_cast<double*>(smem_raw);       // [field][layer][M]
  const int fs = lcap * M;
  FlowShared* ws = reinterpret_cast<FlowShared*>(rec + (size_t)SWD_REC_FIELDS * fs);

  if (gc >= 0) for (int k = t; k < p.kmax[gc]; k += kFlowThreads) swd_period_omegas(1, p.periods[gc][k], &ws->omA[k], &ws->omB[k]);
  if (pc >= 0) for (int k = t; k < p.kmax[pc]; k += kFlowThreads) { double u; swd_period_omegas(0, p.periods[pc][k], &ws->omP[k], &u); }
  if (t < kFlowModels) {
    ws->link[t].na = 0; ws->link[t].a_failed = 0; ws->link[t].del1st = 0.0;
    ws->waitflag[t] = 0;
    int L = 0, mdl = 0;
    if (t < nmod) {
      mdl = BH_MODEL(t);
      const int n = p.nlay[mdl];
      L = (n > p.nlay_lo && n <= p.nlay_hi) ? min(n, lcap) : 0;
    }
    ws->nlay[t] = L; ws->model[t] = mdl;
  }
  for (int i = t; i < kRingE; i += kFlowThreads) ws->ringE[i] = -1;
  for (int i = t; i < kRingC; i += kFlowThreads) { ws->ringW[i] = -1; ws->ringR[i] = -1; }
  if (t == 0) {
    ws->headE = ws->tailE = ws->headW = ws->tailW = ws->headR = ws->tailR = 0u;
    ws->live = 0; ws->error = 0;
  }
  __syncthreads();

  // chain layout: [0, M) first roots of the group curve, [M, 2M) its second roots, then M phase chains
  const int ngrp = gc >= 0 ? 2 * M : 0;
  const int nch = ngrp + (pc >= 0 ? M : 0);
  // context of chain c (all derived from its index)
  auto chain_ctx = [&](int c, SearchCtx& ctx, int& sidx, int& curve) {
    const bool rider = c >= ngrp;
    curve = rider ? pc : gc;
    sidx = rider ? c - ngrp : (c >= M ? c - M : c);
    ctx.omA = rider ? ws->omP : ws->omA; ctx.omB = ws->omB;
    ctx.link = rider ? nullptr : &ws->link[sidx];
    double* r = p.roots + ((size_t)ws->model[sidx] * p.curve_stride + p.curve_off[curve]) * 2;
    ctx.ra = r; ctx.rb = r + p.kmax[curve];
  };
  const int thrE = (spec >> 4) & 63, thrC = (spec >> 10) & 63, napn = 20 << ((spec >> 16) & 7);
  spec &= 15;
  const int spec_n = spec < 1 ? 1 : (spec > kFlowSlots ? kFlowSlots : spec);

  // ---- set up the chains; first-root and phase chains publish their first candidates ----
  {
    int n = 0;
    if (t < kFlowChains) {
      Search s = {};
      s.stage = ST_DONE;
      int live = 0;
      if (t < nch) {
        SearchCtx ctx; int sidx, curve;
        chain_ctx(t, ctx, sidx, curve);
        const int L = sidx < nmod ? ws->nlay[sidx] : 0;
        if (L > 0) {
          const int role = (t < ngrp && t >= M) ? 1 : 0;
          if (search_setup(s, p.rows + (size_t)ws->model[sidx] * stride, 1, L, p.kmax[curve], role, ws->tab + t, kFlowChains)) {
            live = 1;
            if (role == 0) { search_begin_a(s, ctx); n = spec_n; }
            else ws->waitflag[sidx] = 1;           // parked until the model's first first root
          } else if (role == 0 && ctx.link) {
            ctx.link->a_failed = 1;
          }
        }
      }
      ws->st[t] = s;
      ws->pubn[t] = n; ws->outstanding[t] = n;
      if (live) atomicAdd(&ws->live, 1);
    }
    __threadfence_block();
    if (warp < kFlowChains / 32) ring_push<kRingE>(ws->ringE, &ws->tailE, n, t * kFlowSlots, lane);
  }
  // ---- fp64 layer records of the CTA's models (one set serves all chains of a model) ----
  for (int e = t; e < lcap * M; e += kFlowThreads) {
    const int m = e % M, l = e / M;
    if (m < nmod) {
      const int L = ws->nlay[m];
      if (l < L) swd_make_rec(wave, p.rows[(size_t)ws->model[m] * stride + l], l == L - 1, rec + (size_t)l * M + m, fs);
    }
  }
  __syncthreads();

  unsigned long long consumed = 0, evaluated = 0;
  unsigned tasks = 0;
  const double dc = fabs((double)0.005f);
  int idle = 0;

  for (;;) {
    if (ldv(&ws->live) <= 0 || ldv(&ws->error)) break;
    unsigned dW = 0, dR = 0, dE = 0;
    if (lane == 0) {
      dW = ldv(&ws->tailW) - ldv(&ws->headW);
      dR = ldv(&ws->tailR) - ldv(&ws->headR);
      dE = ldv(&ws->tailE) - ldv(&ws->headE);
    }
    dW = __shfl_sync(0xffffffffu, dW, 0); dR = __shfl_sync(0xffffffffu, dR, 0); dE = __shfl_sync(0xffffffffu, dE, 0);
    int payload = -1, n = 0, kind = 0;
    // batches below the thresholds are left to fill up unless this warp has found nothing to do for a while
    const bool takeC = (int)(dR + dW) > 0 && ((int)dR >= thrC || (int)dW >= thrC || idle >= 2);
    const bool takeE = (int)dE > 0 && ((int)dE >= thrE || idle >= 4);
    if (takeC) {
      if (dR >= dW) n = ring_pop<kRingC>(ws->ringR, &ws->headR, &ws->tailR, lane, &payload);
      if (n == 0) n = ring_pop<kRingC>(ws->ringW, &ws->headW, &ws->tailW, lane, &payload);
      if (n == 0) n = ring_pop<kRingC>(ws->ringR, &ws->headR, &ws->tailR, lane, &payload);
      kind = 1;
    }
    if (n == 0 && takeE) { n = ring_pop<kRingE>(ws->ringE, &ws->headE, &ws->tailE, lane, &payload); kind = 2; }
    if (n == 0) {
      if (++idle > (1 << 22)) { if (lane == 0) ws->error = 1; break; }     // watchdog: a lost wake-up must not hang the device
      __nanosleep(napn);
      continue;
    }
    idle = 0;
    ++tasks;
    __threadfence_block();

    if (kind == 2) {
      // ---- evaluate one candidate per lane ----
      int done_chain = -1, done_stage = 0;
      if (payload >= 0) {
        const int c = payload / kFlowSlots, i = payload - c * kFlowSlots;
        const Search& s = ws->st[c];
        const int stage = s.stage;
        const double omega = s.omega;
        const double pend = (stage <= ST_BR_STEP) ? s.c1 : s.c3;
        const double cand = candidate_from(stage, pend, s.idir, s.clow, dc, i);
        const int col = c >= ngrp ? c - ngrp : (c >= M ? c - M : c);
        ws->res[payload] = secular_rec(wave, rec + col, fs, M, ws->nlay[col], fm::div(omega, cand), omega);
        evaluated += 1;
        __threadfence_block();
        if (atomicSub(&ws->outstanding[c], 1) == 1) { done_chain = c; done_stage = stage; }
      }
      __threadfence_block();
      ring_push<kRingC>(ws->ringW, &ws->tailW, (done_chain >= 0 && done_stage <= ST_BR_STEP) ? 1 : 0, done_chain, lane);
      ring_push<kRingC>(ws->ringR, &ws->tailR, (done_chain >= 0 && done_stage > ST_BR_STEP) ? 1 : 0, done_chain, lane);
    } else {
      // ---- consume: one ready chain per lane, all of one kind ----
      int npub = 0, wake = -1;
      const int c = payload;
      if (c >= 0) {
        Search s = ws->st[c];
        SearchCtx ctx; int sidx, curve;
        chain_ctx(c, ctx, sidx, curve);
        const int nval = ws->pubn[c];
        const int k0 = s.k;
        if (nval > 0) consumed += search_consume(s, &ws->res[c * kFlowSlots], nval, ctx);
        bool parked = false;
        if (s.role == 0) {
          // a first root (or the failure) is out: hand it to the model's second-root chain if that one is parked
          if (ctx.link && (s.k != k0 || s.stage == ST_FAILED)) {
            __threadfence_block();
            if (atomicExch(&ws->waitflag[sidx], 0) == 1) wake = M + sidx;
          }
        } else if (s.stage == ST_WAIT) {
          search_poll_b(s, ctx);
          if (s.stage == ST_WAIT) {
            // park: state out first, then the flag, then look again (the first-root chain may have published meanwhile)
            ws->st[c] = s; ws->pubn[c] = 0;
            __threadfence_block();
            atomicExch(&ws->waitflag[sidx], 1);
            __threadfence_block();
            if (s.k < ldv(&ctx.link->na) || ldv(&ctx.link->a_failed)) {
              if (atomicExch(&ws->waitflag[sidx], 0) == 1) search_poll_b(s, ctx);
              else parked = true;                    // the first-root chain's consumer took the flag: it wakes this chain
            } else {
              parked = true;
            }
          }
        }
        if (!parked) {
          npub = s.stage <= ST_BR_STEP ? spec_n : (s.stage < ST_WAIT ? 1 : 0);
          ws->st[c] = s;
          ws->pubn[c] = npub; ws->outstanding[c] = npub;
          if (s.stage >= ST_DONE) atomicSub(&ws->live, 1);
        }
      }
      __threadfence_block();
      ring_push<kRingE>(ws->ringE, &ws->tailE, npub, c * kFlowSlots, lane);
      ring_push<kRingC>(ws->ringW, &ws->tailW, wake >= 0 ? 1 : 0, wake, lane);
    }
  }
  __syncthreads();
  if (t == 0 && ws->error) printf("swd_flow_kernel: CTA %d gave up waiting (live %d)\n", (int)blockIdx.x, ws->live);

  // ---- curve values from the stored roots; validity flag ----
  if (t < nmod && ws->nlay[t] > 0) {
    const int mdl = ws->model[t];
    for (int q = 0; q < 2; ++q) {
      const int curve = q == 0 ? gc : pc;
      if (curve < 0) continue;
      const int igr = p.igr[curve], kmax = p.kmax[curve];
      const int ca = q == 0 ? t : ngrp + t;
      bool ok = ws->st[ca].stage == ST_DONE && !ws->error;
      if (igr > 0) ok = ok && ws->st[M + t].stage == ST_DONE;
      const double* ra = p.roots + ((size_t)mdl * p.curve_stride + p.curve_off[curve]) * 2;
      const double* rb = ra + kmax;
      double* __restrict__ my_curve = p.curves + (size_t)mdl * p.curve_stride + p.curve_off[curve];
      const double* __restrict__ periods = p.periods[curve];
      if (ok)
        for (int k = 0; k < kmax; ++k)
          my_curve[k] = swd_curve_value(igr, periods[k], ra[k], igr > 0 ? rb[k] : 0.0);
      p.tstatus[(size_t)mdl * kMaxTargets + p.target_id[curve]] = ok ? 1 : 0;
    }
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
    const int slot = gc >= 0 ? gc : pc;
    atomicAdd(&p.counters[2 + 2 * (p.counter_base + slot)], (unsigned long long)tasks);
    atomicMax(&p.counters[3 + 2 * (p.counter_base + slot)], (unsigned long long)tasks);
  }
}

size_t flow_smem_bytes(int lcap, int M) {
  return (size_t)SWD_REC_FIELDS * lcap * M * sizeof(double) + sizeof(FlowShared);
}

template <int kWave>
void launch_flow(const SwdLaunch& p, int M, int spec, cudaStream_t st) {
  const int nb = (p.B + M - 1) / M;
  const size_t smem = flow_smem_bytes(p.lcap, M);
  static KernelAttrs attrs;
  bh_configure_kernel(swd_flow_kernel<kWave>, smem, attrs);
  static bool reported = false;
  if (!reported && getenv("BH_DEBUG")) {
    reported = true;
    int res = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, swd_flow_kernel<kWave>, kFlowThreads, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, swd_flow_kernel<kWave>);
    fprintf(stderr, "[bh] swd_flow_kernel<%d>: %d CTAs of %d threads, %d models each, smem %zu B, regs %d, local %zu B, max resident CTAs/SM %d\n",
            kWave, nb, kFlowThreads, M, smem, fa.numRegs, fa.localSizeBytes, res);
  }
  swd_flow_kernel<kWave><<<nb, kFlowThreads, smem, st>>>(p, M, spec);
}

}  // namespace

// Models per CTA: 3 chains per model with a group and a phase curve, at most 96 chain slots and 32 model columns.
int swd_flow_models(const SwdLaunch& p, int lcap, int want) {
  int cpm = 0;
  for (int c = 0; c < p.ncurves; ++c) cpm += p.igr[c] ? 2 : 1;
  if (cpm < 1) cpm = 1;
  int m = want > 0 ? want : 28;
  if (m * cpm > kFlowChains) m = kFlowChains / cpm;
  if (m > kFlowModels) m = kFlowModels;
  while (m > 1 && flow_smem_bytes(lcap, m) > 56 * 1024) --m;
  return m < 1 ? 1 : m;
}

int swd_flow_warp_count(const SwdLaunch& p, int M) { return kFlowWarps * ((p.B + M - 1) / M); }

void launch_swd_flow(const SwdLaunch& p, int M, int spec, cudaStream_t st) {
  if (p.ncurves <= 0 || p.B <= 0) return;
  if (p.wave[0] == 2) launch_flow<2>(p, M, spec, st); else launch_flow<1>(p, M, spec, st);
}

}  // namespace bh
