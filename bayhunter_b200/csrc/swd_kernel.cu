// swd_kernel.cu -- batched SURF96-equivalent dispersion search on sm_100a.
//
// Mapping: a warp owns up to S searches (one search = one dispersion curve of
// one model; S = searches_per_warp).  Lane s < S keeps search s's state machine
// in registers.  Every round the 32 lanes of the warp are dealt out to the
// pending secular-function candidates of those searches: one lane for a search
// that is refining a root (the reference's Neville/bisection sequence is
// strictly serial), and the spare lanes to the searches that are still walking
// their bracket (candidates c1+dc, c1+2dc, ... are independent of earlier
// secular values, so evaluating them ahead is pure speculation that cannot
// change the result).  All lanes then run the secular function together -- the
// only expensive, and fully convergent, part of the kernel.
//
// Model rows are staged once per warp into shared memory as REAL*4 float4 rows
// (exactly the precision SURF96 sees); the row stride is odd so that the
// float4 reads of 8 consecutive lanes hit 8 distinct 16-byte bank groups.
#include "kernels.h"

namespace bh {

namespace {

constexpr int kWarpsPerBlock = 4;

struct WarpShared {
  double c[32];       // pending candidate base (c1 or c3) per owner
  double clow[32];
  double omega[32];
  double del[32];     // secular values per lane
  int stage[32];
  int idir[32];
  int nlay[32];       // rows of the owner's model (constant per warp)
};

struct CurveEmit {
  double* dst;
  __device__ __forceinline__ void operator()(int k, double v) const { dst[k] = v; }
};

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
swd_kernel(SwdLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int S = p.searches_per_warp;
  const int warps_per_curve = (p.B + S - 1) / S;
  const int gw = blockIdx.x * kWarpsPerBlock + wib;   // global warp id
  const int curve = gw / warps_per_curve;
  if (curve >= p.ncurves) return;                     // whole warp exits together
  const int b0 = (gw - curve * warps_per_curve) * S;  // first model of this warp
  const int nsearch = min(S, p.B - b0);

  // ---- shared-memory carve-up: per warp [rows S*stride][WarpShared] ----
  const int stride = p.row_stride;
  const size_t rows_bytes = ((size_t)S * stride * sizeof(LayerRow) + 15) & ~size_t(15);
  const size_t per_warp = rows_bytes + sizeof(WarpShared);
  unsigned char* base = smem_raw + per_warp * wib;
  LayerRow* rows = reinterpret_cast<LayerRow*>(base);
  WarpShared* ws = reinterpret_cast<WarpShared*>(base + rows_bytes);

  // ---- stage the model rows of this warp's models (coalesced 16 B loads) ----
  {
    const LayerRow* src = p.rows + (size_t)b0 * stride;
    const int total = nsearch * stride;
    for (int i = lane; i < total; i += 32) rows[i] = src[i];
  }
  __syncwarp();

  const int wave = p.wave[curve], igr = p.igr[curve], kmax = p.kmax[curve];
  const double* __restrict__ periods = p.periods[curve];
  const int tid = p.target_id[curve];

  // ---- owner state ----
  Search s;
  int myL = 0;
  bool owner = lane < nsearch;
  if (owner) {
    myL = p.nlay[b0 + lane];
    if (myL > stride) myL = stride;
    if (search_setup(s, rows + lane * stride, 1, myL, wave, igr, kmax))
      search_begin_period(s, periods[0]);
  } else {
    s.stage = ST_DONE;
  }
  ws->nlay[lane] = myL;
  __syncwarp();
  double* __restrict__ my_curve =
      p.curves + (size_t)(b0 + (owner ? lane : 0)) * p.curve_stride + p.curve_off[curve];

  unsigned long long consumed = 0, evaluated = 0;
  const int max_spec = p.max_spec;

  for (;;) {
    // ---- phase A: owners publish, warp deals lanes ----
    int want = owner ? search_nwant(s, 32) : 0;
    unsigned active = __ballot_sync(0xffffffffu, want > 0);
    if (active == 0) break;
    unsigned bracket = __ballot_sync(0xffffffffu, want > 1);
    int nact = __popc(active), nbr = __popc(bracket);
    int extra = 32 - nact;
    int cnt = 0;
    if (want > 0) {
      cnt = 1;
      if (want > 1) {
        int rank = __popc(bracket & ((1u << lane) - 1u));
        int add = extra / nbr + (rank < (extra % nbr) ? 1 : 0);
        cnt += add;
        if (cnt > max_spec) cnt = max_spec;
      }
      ws->c[lane] = search_pending_c(s);
      ws->clow[lane] = s.clow;
      ws->omega[lane] = s.omega;
      ws->stage[lane] = s.stage;
      ws->idir[lane] = s.idir;
    }
    unsigned incl = warp_incl_scan((unsigned)cnt, lane);
    unsigned excl = incl - cnt;
    unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned startmask = __reduce_or_sync(0xffffffffu, cnt > 0 ? (1u << excl) : 0u);
    __syncwarp();

    // ---- phase B: every dealt lane evaluates one candidate ----
    if ((unsigned)lane < total) {
      unsigned below = startmask & (0xffffffffu >> (31 - lane));
      int start = 31 - __clz(below);
      int i = lane - start;
      int rank = __popc(below) - 1;
      int own = __fns(active, 0, rank + 1);
      double omega = ws->omega[own];
      double c = candidate_from(ws->stage[own], ws->c[own], ws->idir[own], ws->clow[own],
                                fabs((double)0.005f), i);
      int L = ws->nlay[own];
      double wvno = omega / c;
      ws->del[lane] = secular(wave, rows + own * stride, 1, L, wvno, omega);
      evaluated += 1;
    }
    __syncwarp();

    // ---- phase C: owners consume their values in reference order ----
    if (cnt > 0) {
      CurveEmit emit{my_curve};
      consumed += search_consume(s, &ws->del[excl], cnt, periods, emit);
    }
    __syncwarp();
  }

  if (owner) p.tstatus[(size_t)(b0 + lane) * kMaxTargets + tid] = (s.stage == ST_DONE) ? 1 : 0;

  // ---- counters (one atomic per warp) ----
  for (int d = 16; d > 0; d >>= 1) {
    evaluated += __shfl_down_sync(0xffffffffu, evaluated, d);
    consumed += __shfl_down_sync(0xffffffffu, consumed, d);
  }
  if (lane == 0 && p.counters) {
    atomicAdd(&p.counters[0], consumed);
    atomicAdd(&p.counters[1], evaluated);
  }
}

}  // namespace

void launch_swd(const SwdLaunch& p, cudaStream_t st) {
  if (p.ncurves <= 0 || p.B <= 0) return;
  const int S = p.searches_per_warp;
  const int warps_per_curve = (p.B + S - 1) / S;
  const int warps = warps_per_curve * p.ncurves;
  const int blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const size_t rows_bytes = ((size_t)S * p.row_stride * sizeof(LayerRow) + 15) & ~size_t(15);
  const size_t smem = (rows_bytes + sizeof(WarpShared)) * kWarpsPerBlock;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(swd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  swd_kernel<<<blocks, kWarpsPerBlock * 32, smem, st>>>(p);
}

}  // namespace bh
