// swd_general.cu -- the branches of SURF96 outside BayHunter's default settings:
// higher modes (mode > 1), the earth-flattening transform (flsph = 1) and a water
// layer on top of the stack.
//
// Behavioural reference: src/extensions/surfdisp96.f of BayHunter
//   mode loop / start values of higher modes   :223-271, failure bookkeeping :313-355
//   sphere                                      :486-553
//   water layer                                 :134-135 (llw), :145-149 (jsol), :850-867
//
// The fast kernel (swd_kernel.cu) covers mode 1 / flat earth / solid stacks, which is
// what a BayHunter inversion runs by default.  Everything else goes through this
// kernel: one thread per (model, curve) walks the reference's loop nest
// (mode -> period -> first root [-> second root]) with the same search state machine
// as the fast kernel (swd_core.cuh), one candidate per step, and the secular
// functions in the reference's operation order (secular_*_reforder) on device libm.
// Lanes of a warp stay converged on the secular evaluation, the expensive part.
// Rare settings, so no further tuning: the model rows live in local memory.
#include "kernels.h"
#include "swd_general_core.cuh"

namespace bh {

namespace {

__global__ void __launch_bounds__(64)
swd_general_kernel(SwdGeneralLaunch p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * p.ncurves) return;
  const int curve = idx / p.B, b = idx - curve * p.B;    // a warp = 32 models of one curve
  int L = p.nlay[b];
  if (L > p.lcap) L = p.lcap;
  if (L < 1) return;
  unsigned long long nsec = 0;
  const int err = swd_general_curve(p.rows + (size_t)b * p.row_stride, 1, L, p.wave[curve], p.igr[curve],
                                    p.kmax[curve], p.mode[curve], p.flsph[curve], p.periods[curve],
                                    p.curves + (size_t)b * p.curve_stride + p.curve_off[curve], &nsec);
  p.tstatus[(size_t)b * kMaxTargets + p.target_id[curve]] = err ? 0 : 1;
  if (p.counters) {
    atomicAdd(&p.counters[0], nsec);
    atomicAdd(&p.counters[1], nsec);
  }
}

}  // namespace

void launch_swd_general(const SwdGeneralLaunch& p, cudaStream_t st) {
  if (p.ncurves <= 0 || p.B <= 0) return;
  const int threads = 64;
  const long long total = (long long)p.B * p.ncurves;
  static KernelAttrs attrs;
  bh_configure_kernel(swd_general_kernel, 0, attrs);
  swd_general_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(p);
}

}  // namespace bh
