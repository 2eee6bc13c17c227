// swd_lockstep.cu -- the dispersion search for full batches: every lane owns a chain.
//
// Why a second kernel.  One joint evaluation of 8192 models is ~49 000 independent chains of root
// searches, each ~700-900 secular-function evaluations long and strictly serial (SURF96's
// bracket walk and bisection / Neville refinement, surfdisp96.f:426-481, :582-673).  A B200 has
// room for ~56 000 resident lanes at the kernel's register budget, so every chain is resident
// from the first cycle to the last and the kernel's duration is
//     rounds of the longest chain  x  (latency of one evaluation + latency of the bookkeeping between two)
// -- a latency chain, not a throughput problem.  Measured on swd_kernel (profiles/r02_phase_cycles.txt):
// evaluation 5400 cycles, dealing the lanes 650, consuming the values 1800 per round; the lone-warp
// evaluation is 3280 (Rayleigh) / 2330 (Love) cycles and reorganising it for instruction-level
// parallelism moves it by 2-10 % (tools/micro/secular_bench.cu), so what is left to remove is the
// bookkeeping: 31 % of the round, 35 % of the executed instructions at 11 active lanes.
//
// This kernel keeps swd_kernel's layout (one warp per CTA, fp64 layer records field-major in shared
// memory, a group curve as a first-root chain on lane m and a second-root chain on lane 16 + m) and
// its exact candidate sequence, and strips the bookkeeping down:
//   * a lane evaluates ITS OWN chain's next candidate: no dealing pass, no candidate mailbox, no
//     per-round shared-memory traffic.  The one form of speculation kept is pairwise and costs
//     five shuffles: a lane without work (second-root chain waiting for its first root, finished
//     chain, lane without a chain) evaluates the second bracket candidate of lane (i + 16) mod 32
//     when that chain is walking its bracket.
//   * the search state lives in registers with the pending candidate precomputed (the bracket
//     step is applied when the candidate is issued, not re-derived when its value arrives);
//   * first roots reach the second-root chain through a shared-memory column, not global memory;
//   * the refinement step is one short block: 74 % of SURF96's refinement steps are plain
//     bisections, 17 % a Neville step on a two-point tableau (measured over 18 000 roots).
// Results are bit-identical to swd_kernel's (tests/test_gpu_parity.py::test_engine_device_tensors_and_tunables).
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

#ifndef BH_SWD_LS_MIN_BLOCKS
#define BH_SWD_LS_MIN_BLOCKS 12
#endif

namespace bh {

namespace {

constexpr int LS_TAB_ROWS = 22;

// shared memory of one warp, after the layer records
struct LockstepShared {
  double omA[SWD_MAX_PERIODS], omB[SWD_MAX_PERIODS];   // angular frequencies of the first / second root of each period
  double tab[LS_TAB_ROWS * 32];                        // Neville tableaus x(1..11), y(1..11), one column per lane
  double del1st[16];                                   // getsol's SAVEd del1st of each group model (:415,430)
  // first roots c(k) of the warp's group models follow: [kmax][16]
};

struct Chain {
  double c1, c2, del1, del2;      // bracket ends and their secular values
  double cpend;                   // candidate whose secular value is awaited (c1, c2 or nevill's c3)
  double omega, clow, cprev, del1st;
  double cc, betmx;               // start value cm = cc, dble(REAL*4 max vs)
  int stage, idir, k, nev, m, nctrl;
};

__device__ __forceinline__ double ls_dc() { return fabs((double)0.005f); }

// compares of non-negative doubles (no NaNs): the bit patterns order like the values
__device__ __forceinline__ bool lt_pos(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }
__device__ __forceinline__ bool le_pos(double a, double b) { return __double_as_longlong(a) <= __double_as_longlong(b); }

// getsol's bracket step (:448-458): (c1, idir) -> c2, with the floor at clow
__device__ __forceinline__ double ls_next(double& c1, int& idir, double clow) {
  const double dc = ls_dc();
  double c2 = (idir > 0) ? c1 + dc : c1 - dc;
  if (c2 <= clow) {
    idir = +1;
    c1 = clow;
    c2 = c1 + dc;
  }
  return c2;
}

// The second candidate a chain offers each round: the one SURF96 would most likely ask for next.
//   bracket walk (getsol :448-470): the step after the pending one -- needed unless the pending one closes the
//     bracket; at the first evaluation of a period the walk is taken to go upwards (:432-438: it does unless
//     the secular function changed sign below the start value);
//   refinement (nevill): 74 % of the steps bisect, so the midpoint of the half bracket that the linear
//     interpolant of the bracket ends puts the root in (62 % of the refinement guesses are used).
// A guess is used only if the search then asks for exactly that velocity (bit for bit), so the sequence of
// consumed candidates stays the reference's whatever is guessed.
__device__ __forceinline__ double ls_guess(const Chain& s) {
  if (s.stage <= ST_BR_STEP) {
    double c1 = s.cpend;
    int idir = (s.stage == ST_BR_FIRST) ? +1 : s.idir;
    return ls_next(c1, idir, s.clow);
  }
  const double c3 = s.cpend;
  const double g = fma(s.del1, s.c2 - c3, s.del2 * (c3 - s.c1));          // interpolant at c3, times (c2 - c1)
  const int sg = __double2hiint(g) ^ __double2hiint(s.c2 - s.c1) ^ __double2hiint(s.del1);
  return (sg < 0) ? 0.5 * (s.c1 + c3) : 0.5 * (c3 + s.c2);                // root between c1 and c3 : between c3 and c2
}

// kWave: 1 Love, 2 Rayleigh, 0 both (wave type per curve) -- one launch carries all curves of an evaluation
template <int kWave, int kMinBlocks>
__global__ void __launch_bounds__(32, kMinBlocks)
swd_lockstep_kernel(SwdLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int gw = blockIdx.x;
  int curve = 0;
  while (curve + 1 < p.ncurves && gw >= p.warp_begin[curve + 1]) ++curve;
  const int S = p.spw[curve];                         // models per warp: <= 32 phase, <= 16 group
  const int b0 = (gw - p.warp_begin[curve]) * S;
  const int nsearch = min(S, p.B - b0);
  const int wave = kWave ? kWave : p.wave[curve];
  const int igr = p.igr[curve], kmax = p.kmax[curve];
  const bool kGroup = igr > 0;
  const double* __restrict__ periods = p.periods[curve];
  const int tid = p.target_id[curve];
  const int stride = p.row_stride;
  const int lcap = p.lcap;
  const int* __restrict__ perm = p.perm;
#define BH_MODEL(j) (perm ? perm[b0 + (j)] : b0 + (j))

  double* rec = reinterpret_cast<double*>(smem_raw);
  const int fs = lcap * S;
  LockstepShared* ws = reinterpret_cast<LockstepShared*>(rec + (size_t)SWD_REC_FIELDS * fs);
  double* rootsA = reinterpret_cast<double*>(ws + 1);          // [kmax][16], group curves only

  for (int k = lane; k < kmax; k += 32) swd_period_omegas(igr, periods[k], &ws->omA[k], &ws->omB[k]);

  // Lane layout.  Phase curve: chain m on lane m < S.  Group curve: first-root chain (role A) of model m on lane m,
  // its second-root chain (role B) on lane m + gsh, gsh = 8 when the warp holds at most 8 models (so that lanes
  // 16..31 are free to evaluate guesses), else 16.  The lane 16 places away is every lane's partner: it evaluates
  // this lane's guess whenever it has no candidate of its own.
  const int gsh = (kGroup && S <= 8) ? 8 : 16;
  const int role = (kGroup && (lane & gsh) && lane < 2 * gsh) ? 1 : 0;
  const int sidx = kGroup ? (lane & (gsh - 1)) : lane;
  bool owner = sidx < nsearch && (kGroup ? lane < 2 * gsh : lane < S);
  if (owner) {
    const int n = p.nlay[BH_MODEL(sidx)];
    owner = n > p.nlay_lo && n <= p.nlay_hi;
  }
  if (!__any_sync(0xffffffffu, owner)) {
    if (lane == 0 && p.done) atomicAdd(p.done, 1);
    return;
  }
  const int mymodel = owner ? BH_MODEL(sidx) : 0;
  double* __restrict__ ra = p.roots + ((size_t)mymodel * p.curve_stride + p.curve_off[curve]) * 2;
  double* __restrict__ rb = ra + kmax;
  double* const tabx = ws->tab + lane;
  double* const taby = ws->tab + 11 * 32 + lane;

  Chain s;
  int myL = 0;
  {
    Search t;
    bool ok = false;
    if (owner) {
      myL = min(p.nlay[mymodel], lcap);
      ok = search_setup(t, p.rows + (size_t)mymodel * stride, 1, myL, kmax, role, nullptr, 32);
    }
    s.cc = ok ? t.cc : 0.0; s.betmx = ok ? t.betmx : 0.0;
    s.c1 = s.cc; s.c2 = s.del1 = s.del2 = 0.0; s.cprev = 0.0; s.del1st = 0.0;
    s.idir = 1; s.k = 0; s.nev = 0; s.m = 0; s.nctrl = 0;
    s.clow = s.cc; s.omega = 0.0; s.cpend = s.cc;
    if (!owner) s.stage = ST_DONE;                      // a lane without a chain: evaluates its partner's guesses
    else if (!ok) s.stage = ST_FAILED;
    else if (role) s.stage = ST_WAIT;
    else s.stage = ST_BR_FIRST;                         // :253-256: c1 = clow = cc, ifirst = 1
  }
  __syncwarp();
  if (owner && role == 0 && s.stage == ST_BR_FIRST) s.omega = ws->omA[0];

  // ---- fp64 layer records of this warp's models ----
  for (int t = lane; t < lcap * S; t += 32) {
    const int m = t % S, l = t / S;
    if (m < nsearch) {
      const int n = p.nlay[BH_MODEL(m)];
      const int L = (n > p.nlay_lo && n <= p.nlay_hi) ? min(n, lcap) : 0;
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
  long long cycA = 0, cycB = 0, cycC = 0, nact = 0;
#endif
  const double dc = ls_dc();
  const int partner = lane ^ 16;                       // evaluates my guess when it has nothing of its own
  const int link = lane ^ gsh;                         // group curves: the other chain of my model
  const int colL = sidx | (myL << 8);                  // record column and row count, for the partner

  for (;;) {
    // ---- second-root chains start a period as soon as its first root exists (:282-287) ----
    if (kGroup) {
      const int pk = __shfl_sync(0xffffffffu, s.k, link);
      const int pstage = __shfl_sync(0xffffffffu, s.stage, link);
      if (s.stage == ST_WAIT) {
        if (s.k < pk) {
          s.cprev = rootsA[s.k * 16 + sidx];
          s.del1st = ws->del1st[sidx];
          s.clow = dmul(1.0e-2, dc);                   // cb(k) + one*dc, cb(k) = 0 in the fundamental mode
          s.c1 = dadd(s.cprev, -dmul(1.5, dc));
          s.omega = ws->omB[s.k];
          s.cpend = s.c1;
          s.stage = ST_BR_FIRST;
        } else if (pstage == ST_FAILED) {
          s.stage = ST_FAILED;
        }
      }
    }
    const bool run = s.stage < ST_WAIT;
    if (!__any_sync(0xffffffffu, run)) {
      if (!kGroup || !__any_sync(0xffffffffu, s.stage == ST_WAIT)) break;
      continue;
    }
    ++rounds;
#ifdef BH_SWD_TIMING
    const long long tA = clock64();
#endif
    // ---- every chain offers a guess; a lane without a candidate of its own evaluates its partner's ----
    const double guess = ls_guess(s);
    const int p_run = __shfl_sync(0xffffffffu, (int)run, partner);
    const double p_c = __shfl_sync(0xffffffffu, guess, partner);
    const double p_om = __shfl_sync(0xffffffffu, s.omega, partner);
    const int p_colL = __shfl_sync(0xffffffffu, colL, partner);
    const bool lend = !run && p_run;
    double v = 0.0;
#ifdef BH_SWD_TIMING
    const long long tB = clock64();
#endif
    if (run || lend) {
      const double c = run ? s.cpend : p_c;
      const double om = run ? s.omega : p_om;
      const int cl = run ? colL : p_colL;
      v = secular_rec(wave, rec + (cl & 0xff), fs, S, cl >> 8, fm::div(om, c), om);
      evaluated += 1;
    }
#ifdef BH_SWD_TIMING
    const long long tC = clock64();
#endif
    const double vg = __shfl_sync(0xffffffffu, v, partner);
    const bool hasg = __shfl_sync(0xffffffffu, (int)lend, partner) != 0;

    // ---- consume (getsol :429-470, nevill :587-670): the value of the pending candidate, then the guess's
    //      if the search asks for exactly that velocity next ----
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
    if (pass) {
      if (!(hasg && s.stage < ST_WAIT && __double_as_longlong(s.cpend) == __double_as_longlong(guess))) break;
      v = vg;
    } else if (!run) break;
    int ended = 0;              // 1: root found (s.cpend), 2: no root for this period
    consumed += 1;
    if (s.stage == ST_BR_FIRST) {
      s.del1 = v;
      if (role == 0 && s.k == 0) {                     // ifirst = 1 (:430)
        s.del1st = v;
        if (kGroup) ws->del1st[sidx] = v;
        s.idir = +1;
      } else {
        s.idir = sign_differs(s.del1st, s.del1) ? -1 : +1;     // :432-438
      }
      s.stage = ST_BR_STEP;
    } else if (s.stage == ST_BR_STEP) {
      s.c2 = s.cpend;
      s.del2 = v;
      if (sign_differs(s.del1, s.del2)) {              // :462 -> nevill (:583)
        s.cpend = 0.5 * (s.c1 + s.c2);
        s.nev = 1; s.nctrl = 1;
        s.stage = ST_RF_TOP;
      } else {
        s.c1 = s.c2; s.del1 = s.del2;
        if (lt_pos(s.c1, s.cc) || le_pos(s.betmx + dc, s.c1)) ended = 2;      // :468-469 (cm = cc)
      }
    } else {
      // nevill from label 100 (:587) for the value of c3 = s.cpend.  Everything a step can need is formed
      // side by side from the values at entry -- the range test, the bracket update, the convergence test,
      // the bisection point and the two-point Neville (secant) estimate -- and the outcome is selected at
      // the end: the step's latency is its longest chain, not the sum of its branches.  Compares of
      // non-negative doubles are integer compares (they leave the fp64 pipe, which the evaluations of the
      // other warps keep busy).
      const double c3 = s.cpend;
      const double e_c2 = s.c2, e_d1 = s.del1, e_d2 = s.del2;
      const bool top = s.stage == ST_RF_TOP;
      s.nctrl += top ? 1 : 0;
      const bool maxit = top && s.nctrl >= 100;                                          // :589
      const bool c1lo = lt_pos(s.c1, s.c2);
      const double lo = c1lo ? s.c1 : s.c2, hi = c1lo ? s.c2 : s.c1;
      const bool oor = top && (lt_pos(c3, lo) || lt_pos(hi, c3));                         // :594-598
      const double mid_old = 0.5 * (s.c1 + s.c2);
      const bool opp = sign_differs(v, s.del1);                                          // :604-610
      const double nc1 = opp ? s.c1 : c3, nd1 = opp ? s.del1 : v;
      const double nc2 = opp ? c3 : s.c2, nd2 = opp ? v : s.del2;
      const double s13 = s.del1 - v, s32 = v - s.del2;
      const bool conv = le_pos(fabs(nc1 - nc2), 1.0e-6 * nc1);                            // :614
      const int nev = sign_differs(s13, s32) ? 0 : s.nev;                                // :619
      const double a1 = fabs(nd1), a2 = fabs(nd2);
      const bool ratio = lt_pos(a2, (double)0.01f * a1) || lt_pos(a1, (double)0.01f * a2);   // :625-628 (REAL*4 literal)
      const double mid_new = 0.5 * (nc1 + nc2);
      const double denom = nd2 - nd1;                                                    // fresh tableau: x = (c1, c2), y = (del1, del2)
      const bool bad = lt_pos(fabs(denom), 1.0e-10 * fabs(nd2));
      const double xs = neville_step(nd1, nc2, nd2, nc1, denom);
      if (maxit) {
        ended = (c3 > s.betmx) ? 2 : 1;                                                  // :475-476
      } else if (oor) {
        s.nev = 0; s.cpend = mid_old; s.stage = ST_RF_POST;
      } else {
        s.c1 = nc1; s.del1 = nd1; s.c2 = nc2; s.del2 = nd2;
        s.stage = ST_RF_TOP;
        if (conv) {
          ended = (c3 > s.betmx) ? 2 : 1;
        } else if (ratio || nev == 0) {                                                  // :628-632
          s.cpend = mid_new; s.nev = 1; s.m = 1;
        } else if (nev != 2) {                                                           // two-point tableau (97 % of the Neville steps)
          if (bad) { s.cpend = mid_new; s.nev = 1; s.m = 1; }                            // :663-667
          else { s.cpend = xs; s.nev = 2; s.m = 2; }                                     // :655-661
        } else {
          // the tableau grows (:634-654): it moves to shared memory when its third point arrives -- what the
          // two-point step left in x(1..2), y(1..2) are this step's bracket values at entry
          if (s.m == 2) { tabx[0] = c3; taby[0] = e_d1; tabx[32] = e_c2; taby[32] = e_d2; }
          tabx[s.m * 32] = c3; taby[s.m * 32] = v;
          bool bad2 = false;
          const double ym = v;
          double xn = c3;
          for (int kk = 1; kk <= s.m; ++kk) {
            const int j = s.m - kk;
            const double yj = taby[j * 32];
            const double dn = ym - yj;
            if (lt_pos(fabs(dn), 1.0e-10 * fabs(ym))) { bad2 = true; break; }
            xn = neville_step(yj, xn, ym, tabx[j * 32], dn);
            tabx[j * 32] = xn;
          }
          if (bad2) { s.cpend = mid_new; s.nev = 1; s.m = 1; }
          else { s.cpend = xn; s.nev = 2; s.m = min(s.m + 1, 10); }
        }
      }
    }
    // ---- a search ended: store the root, start the next period ----
    if (ended) {
      const double root = s.cpend;
      if (role == 0) {
        if (ended == 2) {
          s.stage = ST_FAILED;                         // :277 -> err = 1
        } else {
          ra[s.k] = root;                              // c(k) = c1
          if (kGroup) rootsA[s.k * 16 + sidx] = root;
          s.cprev = root;
          s.k += 1;
          if (s.k >= kmax) s.stage = ST_DONE;
          else {                                       // :268-271
            s.c1 = dadd(s.cprev, -dmul(1.5, dc));
            s.clow = s.cc;
            s.omega = ws->omA[s.k];
            s.cpend = s.c1;
            s.stage = ST_BR_FIRST;
          }
        }
      } else {
        rb[s.k] = (ended == 1) ? root : s.cprev;       // :291-293: no second root -> c1 = c(k)
        s.k += 1;
        s.stage = (s.k >= kmax) ? ST_DONE : ST_WAIT;
      }
    } else if (s.stage == ST_BR_STEP) {
      s.cpend = ls_next(s.c1, s.idir, s.clow);          // the next bracket candidate, bracket state advanced
    }
    }
    __syncwarp();              // first roots / del1st in shared memory before the second-root chains poll
#ifdef BH_SWD_TIMING
    { const long long tD = clock64(); cycA += tB - tA; cycB += tC - tB; cycC += tD - tC; nact += __popc(__ballot_sync(0xffffffffu, run || lend)); }
#endif
  }
#ifdef BH_SWD_TIMING
  if (lane == 0 && ((gw - p.warp_begin[curve]) % 61) == 5)
    printf("ls timing curve %d wave %d igr %d warp %d: rounds %u  lend %lld  eval %lld  consume %lld cycles per round, lanes %.1f\n", curve, wave, igr, gw,
           rounds, cycA / max(rounds, 1u), cycB / max(rounds, 1u), cycC / max(rounds, 1u), (double)nact / max(rounds, 1u));
#endif

  // ---- curve values from the stored roots; validity flag ----
  {
    const bool done = s.stage == ST_DONE;
    const unsigned bdone = __ballot_sync(0xffffffffu, done);
    if (owner && role == 0) {
      bool ok = done;
      if (kGroup) ok = ok && ((bdone >> (lane + gsh)) & 1u);
      double* __restrict__ my_curve = p.curves + (size_t)mymodel * p.curve_stride + p.curve_off[curve];
      if (ok)
        for (int k = 0; k < kmax; ++k)
          my_curve[k] = swd_curve_value(igr, periods[k], ra[k], kGroup ? rb[k] : 0.0);
      p.tstatus[(size_t)mymodel * kMaxTargets + tid] = ok ? 1 : 0;
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
    atomicAdd(&p.counters[2 + 2 * (p.counter_base + curve)], (unsigned long long)rounds);
    atomicMax(&p.counters[3 + 2 * (p.counter_base + curve)], (unsigned long long)rounds);
  }
}

size_t lockstep_smem_bytes(int lcap, int S, int igr, int kmax) {
  return (size_t)SWD_REC_FIELDS * lcap * S * sizeof(double) + sizeof(LockstepShared) +
         (igr ? (size_t)kmax * 16 * sizeof(double) : 0);
}

template <int kWave, int kMinBlocks>
void launch_inst(const SwdLaunch& p, int warps, size_t smem, cudaStream_t st) {
  static KernelAttrs attrs;
  bh_configure_kernel(swd_lockstep_kernel<kWave, kMinBlocks>, smem, attrs);
  static bool reported = false;
  if (!reported && getenv("BH_DEBUG")) {
    reported = true;
    int nb = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, swd_lockstep_kernel<kWave, kMinBlocks>, 32, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, swd_lockstep_kernel<kWave, kMinBlocks>);
    fprintf(stderr, "[bh] swd_lockstep_kernel<%d,%d>: %d warps, smem %zu B, regs %d, local %zu B, max resident CTAs/SM %d\n",
            kWave, kMinBlocks, warps, smem, fa.numRegs, fa.localSizeBytes, nb);
  }
  swd_lockstep_kernel<kWave, kMinBlocks><<<warps, 32, smem, st>>>(p);
}

}  // namespace

// One launch for all curves of `p` (Rayleigh first, group before phase: the engine's order).
void launch_swd_lockstep(SwdLaunch& p, cudaStream_t st) {
  if (p.ncurves <= 0 || p.B <= 0) return;
  int warps = 0;
  size_t smem = 0;
  for (int c = 0; c < p.ncurves; ++c) {
    if (p.igr[c] && p.spw[c] > 16) p.spw[c] = 16;
    if (p.spw[c] > 32) p.spw[c] = 32;
    p.warp_begin[c] = warps;
    warps += (p.B + p.spw[c] - 1) / p.spw[c];
    const size_t b = lockstep_smem_bytes(p.lcap, p.spw[c], p.igr[c], p.kmax[c]);
    if (b > smem) smem = b;
  }
  p.warp_begin[p.ncurves] = warps;
  int kind = p.wave[0];
  for (int c = 1; c < p.ncurves; ++c) if (p.wave[c] != kind) kind = 0;
  static int nsm = 0;
  if (nsm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm < 1) nsm = 148;
  }
  // register budget by grid size, so that the whole grid is resident at once: 12 warps per SM at 143 registers,
  // 16 at 124, 20 at 96 (a 56-byte spill of search state, outside the layer loops)
  const int per_sm = (warps + nsm - 1) / nsm;
  if (kind == 1) {
    if (per_sm <= 16) launch_inst<1, 16>(p, warps, smem, st); else launch_inst<1, 20>(p, warps, smem, st);
  } else if (kind == 2) {
    if (per_sm <= 12) launch_inst<2, 12>(p, warps, smem, st);
    else if (per_sm <= 16) launch_inst<2, 16>(p, warps, smem, st);
    else launch_inst<2, 20>(p, warps, smem, st);
  } else {
    if (per_sm <= 12) launch_inst<0, 12>(p, warps, smem, st);
    else if (per_sm <= 16) launch_inst<0, 16>(p, warps, smem, st);
    else launch_inst<0, 20>(p, warps, smem, st);
  }
}

}  // namespace bh
