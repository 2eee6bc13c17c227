// swd_kernel.cu -- batched SURF96-equivalent dispersion search on sm_100a.
//
// Mapping: one warp per CTA; a warp owns up to S models of ONE curve.  A phase
// curve is one serial chain of root searches per model (lane m < S keeps its
// state machine in registers); a group curve is two chains per model -- first
// roots on lane m, second roots on lane S + m, see swd_core.cuh -- so S <= 16
// there.  Every round the 32 lanes of the warp are dealt out to the pending
// secular-function candidates of those chains: one lane for a chain that is
// refining a root (the reference's Neville/bisection sequence is strictly
// serial), and the spare lanes to the chains that are still walking their
// bracket (candidates c1+dc, c1+2dc, ... are independent of earlier secular
// values, so evaluating them ahead is pure speculation that cannot change the
// result).  All lanes then run the secular function together -- the only
// expensive, and fully convergent, part of the kernel.
//
// Shared memory per warp: the fp64 layer records of its S models
// (swd_core.cuh: swd_make_rec), field-major [field][layer][search] so that the
// 32 lanes of a round read 32 distinct banks (or broadcast when several lanes
// speculate for the same search), plus one small mailbox (WarpShared).
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

// resident warps (= CTAs) per SM the register allocation must allow
// Register budget of the instantiations that carry the Rayleigh code, as "resident warps (= CTAs)
// per SM the allocation must allow":
//   12 -> 147 registers: the faster code (~10 %) whenever the whole grid is resident anyway (<= 12 warps per SM)
//   14 -> 128 registers, no spills: the whole grid of a full batch (~2048 warps on 148 SMs) is
//         resident at once instead of leaving a second wave (4.65 -> 4.2 ms at B = 8192)
// launch_one() picks per launch from the grid size (profiles/r01_variants.txt).
#ifndef BH_SWD_MIN_BLOCKS_RAYLEIGH
#define BH_SWD_MIN_BLOCKS_RAYLEIGH 12
#endif
#ifndef BH_SWD_MIN_BLOCKS_RAYLEIGH_DENSE
#define BH_SWD_MIN_BLOCKS_RAYLEIGH_DENSE 14
#endif
#ifndef BH_SWD_MIN_BLOCKS_LOVE
#define BH_SWD_MIN_BLOCKS_LOVE 16
#endif

namespace bh {

namespace {

struct WarpShared {
  double c[32];       // pending candidate base (c1 or c3) per owner
  double clow[32];    // walking chain: lower end of the window; refining chain: c1
  double c2[32];      // refining chain: c2
  double omega[32];
  double del[32];     // secular values per lane
  double omA[SWD_MAX_PERIODS], omB[SWD_MAX_PERIODS];   // per-period angular frequencies of this curve
  SearchLink link[16];                                 // role A -> role B mailboxes (group curves)
  int stage[32];
  int idir[32];
  int nlay[32];       // rows of the owner's model (constant per warp)
  int col[32];        // record column (model slot) of the owner
  int owner_at[32];   // owner lane of the candidate run that starts at this lane
  double tab[SWD_TAB_ROWS * 32];   // Neville tableaus, one column per lane
};

// kDirect: every chain evaluates only its own next candidate (no speculation, no
// dealing, no mailbox traffic) -- the mode for large batches, where all 32 lanes
// of a warp own a chain anyway.  Otherwise spare lanes are dealt to the chains
// that are walking a bracket (small batches, the single-model shims).
// kWave: 1 Love, 2 Rayleigh -- one instantiation per wave type, so that the Love
// chains do not pay for the Rayleigh code's registers; the two are launched on
// different streams and share the SMs.
template <bool kDirect, int kWave, int kMinBlocks>
__global__ void __launch_bounds__(32, kMinBlocks)
swd_kernel(SwdLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  int gw = blockIdx.x;                                // work item = global warp id
  if (kWave == 0 && p.queue) {
    // mixed launch: claim an item of the wave type this SM is dedicated to
    int item = -1;
    if (lane == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      int t = ((int)smid < p.sm_split) ? 0 : 1;
      if (atomicAdd(&p.queue[2 + smid], 1) >= p.type_quota[t]) t ^= 1;
      int i = atomicAdd(&p.queue[t], 1);
      if (p.type_begin[t] + i >= p.type_begin[t + 1]) {
        t ^= 1;
        i = atomicAdd(&p.queue[t], 1);
      }
      if (p.type_begin[t] + i < p.type_begin[t + 1]) item = p.type_begin[t] + i;
    }
    gw = __shfl_sync(0xffffffffu, item, 0);
    if (gw < 0) return;
  }
  int curve = 0;
  while (curve + 1 < p.ncurves && gw >= p.warp_begin[curve + 1]) ++curve;
  const int S = p.spw[curve];                         // models per warp (<= 16 for group curves)
  const int b0 = (gw - p.warp_begin[curve]) * S;      // first model of this warp
  const int nsearch = min(S, p.B - b0);
  const int wave = kWave ? kWave : p.wave[curve];     // kWave == 0: mixed launch, wave per curve
  const int igr = p.igr[curve], kmax = p.kmax[curve];
  const double* __restrict__ periods = p.periods[curve];
  const int tid = p.target_id[curve];
  const int stride = p.row_stride;                    // rows per model in p.rows
  const int lcap = p.lcap;                            // layer capacity of the records

  // model handled by slot j of this warp: models are dealt to warps in the order of p.perm
  // (sorted by layer count, so that the lanes of a warp run layer loops of similar length)
  const int* __restrict__ perm = p.perm;
#define BH_MODEL(j) (perm ? perm[b0 + (j)] : b0 + (j))

  // ---- shared-memory carve-up: [records 6*lcap*S doubles][WarpShared] ----
  double* rec = reinterpret_cast<double*>(smem_raw);
  const int fs = lcap * S;                            // field stride; layer stride = S
  WarpShared* ws = reinterpret_cast<WarpShared*>(rec + (size_t)SWD_REC_FIELDS * fs);

  // ---- per-period tables of this curve ----
  for (int k = lane; k < kmax; k += 32) swd_period_omegas(igr, periods[k], &ws->omA[k], &ws->omB[k]);
  if (lane < 16) { ws->link[lane].na = 0; ws->link[lane].a_failed = 0; ws->link[lane].del1st = 0.0; }

  // ---- owner state: lanes [0, nsearch) run the first-root chains (role A), lanes
  //      [S, S + nsearch) the second-root chains of group curves (role B) ----
  const int role = (igr > 0 && lane >= S) ? 1 : 0;
  const int sidx = role ? lane - S : lane;            // model slot within the warp
  bool owner = sidx < nsearch && (role == 0 || igr > 0) && lane < (igr > 0 ? 2 * S : S);
  // This launch takes the models whose row count lies in (nlay_lo, nlay_hi]: the record
  // capacity `lcap` follows the layer counts seen lately, deeper models go to a second launch
  // with full capacity (engine.cu).  Warps without a model in range leave at once.
  if (owner) {
    const int n = p.nlay[BH_MODEL(sidx)];
    owner = n > p.nlay_lo && n <= p.nlay_hi;
  }
  if (!__any_sync(0xffffffffu, owner)) {
    if (lane == 0 && p.done) atomicAdd(p.done, 1);
    return;
  }
  Search s;
  SearchCtx ctx;
  ctx.omA = ws->omA; ctx.omB = ws->omB;
  ctx.link = (igr > 0 && owner) ? &ws->link[sidx] : nullptr;
  {
    double* r = p.roots + ((size_t)BH_MODEL(owner ? sidx : 0) * p.curve_stride + p.curve_off[curve]) * 2;
    ctx.ra = r; ctx.rb = r + kmax;
  }
  int myL = 0;
  const int mymodel = owner ? BH_MODEL(sidx) : 0;
  __syncwarp();
  if (owner) {
    myL = p.nlay[mymodel];
    if (myL > lcap) myL = lcap;
    if (search_setup(s, p.rows + (size_t)mymodel * stride, 1, myL, kmax, role, ws->tab + lane, 32)) {
      if (role == 0) search_begin_a(s, ctx);
    } else if (role == 0 && ctx.link) {
      ctx.link->a_failed = 1;
    }
  } else {
    s.stage = ST_DONE;
  }
  ws->nlay[lane] = myL;
  ws->col[lane] = sidx;
  __syncwarp();

  // ---- derive the fp64 layer records of this warp's models ----
  for (int t = lane; t < lcap * S; t += 32) {
    const int m = t % S, l = t / S;
    if (m < nsearch) {
      const int L = ws->nlay[m];                      // 0 for slots outside this launch's range
      if (l < L) {
        const LayerRow r = p.rows[(size_t)BH_MODEL(m) * stride + l];
        swd_make_rec(wave, r, l == L - 1, rec + (size_t)l * S + m, fs);
      }
    }
  }
  __syncwarp();

  unsigned long long consumed = 0, evaluated = 0;
  unsigned rounds = 0;
#ifdef BH_SWD_TIMING
  long long cycA = 0, cycB = 0, cycC = 0, lanesB = 0, cycS[3] = {0, 0, 0}, nS[3] = {0, 0, 0};
#endif
  const int max_spec = p.max_spec;
  const double dc = fabs((double)0.005f);

  if (kDirect) {
    const double* __restrict__ myrec = rec + sidx;
    for (;;) {
      if (role) search_poll_b(s, ctx);
      const bool run = s.stage < ST_WAIT;
      if (!__any_sync(0xffffffffu, run)) {
        if (!__any_sync(0xffffffffu, s.stage == ST_WAIT)) break;
        __syncwarp();
        continue;
      }
      ++rounds;
      if (run) {
        const double c = candidate_from(s.stage, search_pending_c(s), s.idir, s.clow, dc, 0);
        const double v = secular_rec(wave, myrec, fs, S, myL, fm::div(s.omega, c), s.omega);
        consumed += search_consume(s, &v, 1, ctx);
      }
      __syncwarp();    // orders role A's published roots before role B's next poll
    }
    evaluated = consumed;
  } else
  for (;;) {
#ifdef BH_SWD_TIMING
    long long tA = clock64();
#endif
    // ---- phase A: owners publish, warp deals lanes ----
    if (role) search_poll_b(s, ctx);
    int want = search_nwant(s, 32);
    unsigned active = __ballot_sync(0xffffffffu, want > 0);
    if (active == 0) {
      // nothing to evaluate: finished, unless a role-B chain is still waiting for a
      // root that role A published in this very round
      if (__ballot_sync(0xffffffffu, s.stage == ST_WAIT) == 0) break;
      __syncwarp();
      continue;
    }
    ++rounds;
    unsigned bracket = __ballot_sync(0xffffffffu, want > 1);
    // lanes per chain and where each chain's run of lanes starts: closed form from the two ballots
    // (swd_core.cuh: deal_lanes) -- no shuffle scan, no integer division on the round's critical path
    const LaneDeal deal = deal_lanes(active, bracket, lane, max_spec);
    const int cnt = deal.cnt;
    const unsigned excl = (unsigned)deal.excl, total = (unsigned)deal.total;
    if (cnt > 0) {
      ws->c[lane] = search_pending_c(s);
      ws->clow[lane] = s.stage > ST_BR_STEP ? s.c1 : s.clow;
      ws->c2[lane] = s.c2;
      ws->omega[lane] = s.omega;
      ws->stage[lane] = s.stage;
      ws->idir[lane] = s.idir;
    }
    unsigned startmask = __reduce_or_sync(0xffffffffu, cnt > 0 ? (1u << excl) : 0u);
    if (cnt > 0) ws->owner_at[excl] = lane;
    __syncwarp();

#ifdef BH_SWD_TIMING
    long long tB = clock64();
#endif
    // ---- phase B: every dealt lane evaluates one candidate ----
    if ((unsigned)lane < total) {
      unsigned starts = startmask & (0xffffffffu >> (31 - lane));
      int start = 31 - __clz(starts);
      int i = lane - start;
      int own = ws->owner_at[start];
      double omega = ws->omega[own];
      double c = candidate_from(ws->stage[own], ws->c[own], ws->idir[own], ws->clow[own], dc, i, ws->c2[own]);
      int L = ws->nlay[own];
      double wvno = fm::div(omega, c);
      ws->del[lane] = secular_rec(wave, rec + ws->col[own], fs, S, L, wvno, omega);
      evaluated += 1;
    }
    __syncwarp();

#ifdef BH_SWD_TIMING
    long long tC = clock64();
#endif
    // ---- phase C: owners consume their values in reference order ----
#ifdef BH_SWD_TIMING
    { const int st0 = s.stage; const long long q0 = clock64();
      if (cnt > 0) consumed += search_consume(s, &ws->del[excl], cnt, ctx);
      const long long dq = clock64() - q0;
      if (cnt > 0) { if (st0 == ST_BR_FIRST) { cycS[0] += dq; nS[0]++; } else if (st0 == ST_BR_STEP) { cycS[1] += dq; nS[1]++; } else { cycS[2] += dq; nS[2]++; } } }
#else
    if (cnt > 0) consumed += search_consume(s, &ws->del[excl], cnt, ctx);
#endif
    __syncwarp();
#ifdef BH_SWD_TIMING
    long long tD = clock64();
    cycA += tB - tA; cycB += tC - tB; cycC += tD - tC; lanesB += total;
#endif
  }
#ifdef BH_SWD_TIMING
  for (int q = 0; q < 3; ++q)
    for (int d = 16; d > 0; d >>= 1) { cycS[q] += __shfl_down_sync(0xffffffffu, cycS[q], d); nS[q] += __shfl_down_sync(0xffffffffu, nS[q], d); }
  if (lane == 0 && ((gw - p.warp_begin[curve]) % 97) == 5)
    printf("swd stages curve %d wave %d igr %d: first %lld x %lld  step %lld x %lld  refine %lld x %lld (lane-calls x cycles per call)\n", curve, wave, igr,
           nS[0], cycS[0] / max(nS[0], 1LL), nS[1], cycS[1] / max(nS[1], 1LL), nS[2], cycS[2] / max(nS[2], 1LL));
  if (lane == 0 && ((gw - p.warp_begin[curve]) % 97) == 5)
    printf("swd timing curve %d wave %d igr %d warp %d: rounds %u  A %lld  B %lld  C %lld cycles per round, lanes %.1f\n", curve, wave, igr, gw,
           rounds, cycA / max(rounds, 1u), cycB / max(rounds, 1u), cycC / max(rounds, 1u), (double)lanesB / max(rounds, 1u));
#endif

  // ---- curve values from the stored roots; validity flag ----
  // role-A lanes write the curve; the second roots of a group curve were stored by
  // the role-B lane S places up (same warp: ordered by the __syncwarp above)
  {
    const bool done = s.stage == ST_DONE;
    const unsigned bdone = __ballot_sync(0xffffffffu, done);
    if (owner && role == 0) {
      bool ok = done;
      if (igr > 0) ok = ok && ((bdone >> (lane + S)) & 1u);
      double* __restrict__ my_curve = p.curves + (size_t)mymodel * p.curve_stride + p.curve_off[curve];
      if (ok)
        for (int k = 0; k < kmax; ++k)
          my_curve[k] = swd_curve_value(igr, periods[k], ctx.ra[k], igr > 0 ? ctx.rb[k] : 0.0);
      p.tstatus[(size_t)mymodel * kMaxTargets + tid] = ok ? 1 : 0;
    }
  }

#undef BH_MODEL
  // ---- counters (one atomic per warp) ----
  for (int d = 16; d > 0; d >>= 1) {
    evaluated += __shfl_down_sync(0xffffffffu, evaluated, d);
    consumed += __shfl_down_sync(0xffffffffu, consumed, d);
  }
  if (lane == 0 && p.done) atomicAdd(p.done, 1);       // retired warps: releases the RF stream's gate (engine.cu)
  if (lane == 0 && p.counters) {
    atomicAdd(&p.counters[0], consumed);
    atomicAdd(&p.counters[1], evaluated);
    atomicAdd(&p.counters[2 + 2 * (p.counter_base + curve)], (unsigned long long)rounds);
    atomicMax(&p.counters[3 + 2 * (p.counter_base + curve)], (unsigned long long)rounds);
  }
}

}  // namespace

// One thread that waits until `threshold` dispersion warps have retired (bounded: max_ns), so that
// the work queued behind it in its stream starts in the dispersion kernel's tail whatever the order
// in which the two streams happened to reach the device.
__global__ void swd_gate_kernel(const int* done, int threshold, long long max_ns) {
  long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if (*(volatile const int*)done >= threshold) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > max_ns) break;
    __nanosleep(2000);
  }
}

void launch_swd_gate(const int* done, int threshold, cudaStream_t st) {
  swd_gate_kernel<<<1, 1, 0, st>>>(done, threshold, 100LL * 1000 * 1000);
}

int swd_warp_count(const SwdLaunch& p) {
  int warps = 0;
  for (int c = 0; c < p.ncurves; ++c) warps += (p.B + p.spw[c] - 1) / p.spw[c];
  return warps;
}

size_t swd_smem_bytes(int lcap, int S) {
  return (size_t)SWD_REC_FIELDS * lcap * S * sizeof(double) + sizeof(WarpShared);
}

template <bool kDirect, int kWave, int kMinBlocks>
static void launch_inst(const SwdLaunch& p, int warps, size_t smem, cudaStream_t st) {
  static KernelAttrs attrs;
  bh_configure_kernel(swd_kernel<kDirect, kWave, kMinBlocks>, smem, attrs);
  static bool reported = false;
  if (!reported && getenv("BH_DEBUG")) {
    reported = true;
    int nb = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, swd_kernel<kDirect, kWave, kMinBlocks>, 32, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, swd_kernel<kDirect, kWave, kMinBlocks>);
    fprintf(stderr, "[bh] swd_kernel<%d,%d,%d>: %d warps, smem %zu B, regs %d, local %zu B, carveout %d, max resident CTAs/SM %d\n",
            (int)kDirect, kWave, kMinBlocks, warps, smem, fa.numRegs, fa.localSizeBytes, fa.preferredShmemCarveout, nb);
  }
  swd_kernel<kDirect, kWave, kMinBlocks><<<warps, 32, smem, st>>>(p);
}

template <bool kDirect, int kWave>
static void launch_one(const SwdLaunch& p, int warps, size_t smem, cudaStream_t st) {
  if (kWave == 1) { launch_inst<kDirect, 1, BH_SWD_MIN_BLOCKS_LOVE>(p, warps, smem, st); return; }
  static int nsm = 0;
  if (nsm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm < 1) nsm = 148;
  }
  // the roomy build holds 12 warps per SM (147 registers are allocated as 160 per thread); a larger grid
  // takes the 128-register build (measured cliff between 1776 and 1792 warps on 148 SMs)
  if (warps > 12 * nsm) launch_inst<kDirect, kWave, BH_SWD_MIN_BLOCKS_RAYLEIGH_DENSE>(p, warps, smem, st);
  else launch_inst<kDirect, kWave, BH_SWD_MIN_BLOCKS_RAYLEIGH>(p, warps, smem, st);
}

// Curves of one wave type run the specialised instantiation; a mixed set of curves
// runs the generic one (one launch, Rayleigh and Love warps side by side).
void launch_swd(SwdLaunch& p, cudaStream_t st) {
  if (p.ncurves <= 0 || p.B <= 0) return;
  int warps = 0, smax = 1;
  for (int c = 0; c < p.ncurves; ++c) {
    p.warp_begin[c] = warps;
    warps += (p.B + p.spw[c] - 1) / p.spw[c];
    if (p.spw[c] > smax) smax = p.spw[c];
  }
  p.warp_begin[p.ncurves] = warps;
  {
    int c = 0;
    while (c < p.ncurves && p.wave[c] == 2) ++c;      // Rayleigh curves come first
    p.type_begin[0] = 0; p.type_begin[1] = p.warp_begin[c]; p.type_begin[2] = warps;
  }
  const size_t smem = swd_smem_bytes(p.lcap, smax);
  int kind = p.wave[0];
  for (int c = 1; c < p.ncurves; ++c) if (p.wave[c] != kind) kind = 0;
  if (p.direct) {
    if (kind == 1) launch_one<true, 1>(p, warps, smem, st);
    else if (kind == 2) launch_one<true, 2>(p, warps, smem, st);
    else launch_one<true, 0>(p, warps, smem, st);
  } else {
    if (kind == 1) launch_one<false, 1>(p, warps, smem, st);
    else if (kind == 2) launch_one<false, 2>(p, warps, smem, st);
    else launch_one<false, 0>(p, warps, smem, st);
  }
}

}  // namespace bh
