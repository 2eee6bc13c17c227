// Microbenchmark (developer tool): secular-function formulations on their own, one warp per CTA as in
// swd_kernel, for several resident-warp counts.  Reports ns and SM cycles per warp evaluation and per
// sub-partition, and checks that every formulation returns the same bits as secular_*_rec.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o secular_bench tools/micro/secular_bench.cu
//   ./secular_bench [L=6] [rounds=400]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../bayhunter_b200/csrc/swd_eval.cuh"

using namespace bh;

constexpr int S = 16;   // model columns per warp

template <int WAVE, int V>
__device__ __forceinline__ double eval_one(const double* rec, int fs, int L, double wvno, double omega, double* ht) {
  if (WAVE == 2) {
    if (V == 0) return secular_rayleigh_rec(rec, fs, S, L, wvno, omega);
    if (V == 1) return secular_rayleigh_2pass<2>(rec, fs, S, L, wvno, omega, ht, 32);
    if (V == 2) return secular_rayleigh_2pass<1>(rec, fs, S, L, wvno, omega, ht, 32);
    if (V == 3) return secular_rayleigh_2pass<3>(rec, fs, S, L, wvno, omega, ht, 32);
    if (V == 4) return secular_rayleigh_rot(rec, fs, S, L, wvno, omega);
    if (V == 6) { const double wv[2] = {wvno, wvno * 0.9993}; double o[2]; secular_rayleigh_rec2(rec, fs, S, L, wv, omega, o); return o[0] + 1e-3 * o[1]; }
    return secular_rayleigh_unrolled<6>(rec, fs, S, wvno, omega);
  } else {
    if (V == 0) return secular_love_rec(rec, fs, S, L, wvno, omega);
    if (V == 1) return secular_love_2pass<2>(rec, fs, S, L, wvno, omega, ht, 32);
    if (V == 2) return secular_love_2pass<3>(rec, fs, S, L, wvno, omega, ht, 32);
    if (V == 3) return secular_love_2pass<6>(rec, fs, S, L, wvno, omega, ht, 32);
    if (V == 4) return secular_love_grp<2>(rec, fs, S, L, wvno, omega);
    if (V == 6) { const double wv[2] = {wvno, wvno * 0.9993}; double o[2]; secular_love_rec2(rec, fs, S, L, wv, omega, o); return o[0] + 1e-3 * o[1]; }
    return secular_love_grp<3>(rec, fs, S, L, wvno, omega);
  }
}

template <int WAVE, int V>
__global__ void __launch_bounds__(32) kbench(const LayerRow* rows, int nmodels, int L, int rounds, double* out,
                                              long long* cyc) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const int fs = L * S;
  double* rec = sm;
  double* ht = rec + SWD_REC_FIELDS * fs + lane;
  for (int t = lane; t < L * S; t += 32) {
    const int m = t % S, l = t / S;
    const LayerRow r = rows[(size_t)((blockIdx.x * S + m) % nmodels) * L + l];
    swd_make_rec(WAVE, r, l == L - 1, rec + l * S + m, fs);
  }
  __syncwarp();
  const int col = lane % S;
  const LayerRow top = rows[(size_t)((blockIdx.x * S + col) % nmodels) * L];
  const LayerRow bot = rows[(size_t)((blockIdx.x * S + col) % nmodels) * L + L - 1];
  const double T = 1.0 + 39.0 * ((lane * 7 + blockIdx.x) % 32) / 31.0;
  const double omega = 6.283185307179586 / T;
  const double clo = 0.8 * top.z, chi = bot.z;
  double c = clo + 0.001 * lane;
  double acc = 0.0;
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < rounds; ++r) {
    c += 0.005;
    if (c > chi) c = clo;
    const double v = eval_one<WAVE, V>(rec + col, fs, L, fm::div(omega, c), omega, ht);
    acc += v;
    c += 1e-9 * v;      // the next candidate depends on this value, as in the search
  }
  long long t1 = clock64();
  out[(size_t)blockIdx.x * 32 + lane] = acc;
  if (blockIdx.x == 0 && lane == 0 && cyc) *cyc = t1 - t0;
}

template <int WAVE, int V>
static void run(const LayerRow* d_rows, int nmodels, int L, int rounds, std::vector<double>* ref, const char* name) {
  const size_t smem = (size_t)(SWD_REC_FIELDS * L * S + SWD_HT_SLOTS * L * 32) * sizeof(double);
  cudaFuncSetAttribute(kbench<WAVE, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kbench<WAVE, V>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kbench<WAVE, V>);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kbench<WAVE, V>, 32, smem);
  printf("%-28s regs %3d smem %6zu B occupancy %2d CTAs/SM\n", name, fa.numRegs, smem, occ);
  const int grids_all[] = {1, 148, 148 * 4, 148 * 8, 148 * 10, 148 * 12, 148 * 16};
  std::vector<int> grids(grids_all, grids_all + 7);
  if (getenv("MB_ONLY")) {          // "wave,variant,grid": one configuration only (profiling)
    int w = 0, v = 0, g = 0;
    sscanf(getenv("MB_ONLY"), "%d,%d,%d", &w, &v, &g);
    if (w != WAVE || v != V) return;
    grids.assign(1, g);
  }
  double* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, sizeof(double) * 32 * 148 * 16);
  cudaMalloc(&d_cyc, sizeof(long long));
  for (int g : grids) {
    if (g > 148 * occ && g > 148) continue;
    if (getenv("MB_ONLY")) printf("  (only this configuration)\n");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kbench<WAVE, V><<<g, 32, smem>>>(d_rows, nmodels, L, rounds / 4 + 1, d_out, nullptr);
    cudaEventRecord(e0);
    kbench<WAVE, V><<<g, 32, smem>>>(d_rows, nmodels, L, rounds, d_out, d_cyc);
    cudaEventRecord(e1);
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(ce)); exit(1); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double wps = g / 592.0;    // warps per sub-partition
    const double cyc_per = (double)cyc / rounds;
    printf("  grid %5d (%.2f warps/SMSP)  %8.3f ms  %7.1f cycles per warp-eval (CTA 0)  %7.1f cycles per warp-eval per SMSP\n",
           g, wps, ms, cyc_per, wps >= 1 ? cyc_per / wps : cyc_per);
    if (g == 148) {
      std::vector<double> h(32 * 148);
      cudaMemcpy(h.data(), d_out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost);
      if (ref->empty()) *ref = h;
      else {
        size_t bad = 0;
        double worst = 0;
        for (size_t i = 0; i < h.size(); ++i)
          if (h[i] != (*ref)[i]) { ++bad; worst = fmax(worst, fabs(h[i] - (*ref)[i])); }
        printf("  vs secular_rec: %zu of %zu sums differ (max |diff| %.3g)\n", bad, h.size(), worst);
      }
    }
  }
  cudaFree(d_out); cudaFree(d_cyc);
}

int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 6;
  const int rounds = argc > 2 ? atoi(argv[2]) : 400;
  const int nmodels = 4096;
  std::vector<LayerRow> rows((size_t)nmodels * L);
  srand(12345);
  for (int m = 0; m < nmodels; ++m) {
    std::vector<double> vs(L);
    for (int l = 0; l < L; ++l) vs[l] = 2.0 + 3.0 * rand() / RAND_MAX;
    for (int i = 0; i < L; ++i) for (int j = i + 1; j < L; ++j) if (vs[j] < vs[i]) std::swap(vs[i], vs[j]);
    const double vpvs = 1.4 + 0.7 * rand() / RAND_MAX;
    for (int l = 0; l < L; ++l) {
      LayerRow r;
      r.x = l == L - 1 ? 0.f : (float)(0.5 + 60.0 / L * rand() / RAND_MAX);
      r.z = (float)vs[l];
      r.y = (float)(vs[l] * vpvs);
      r.w = (float)(0.32 * vs[l] * vpvs + 0.77);
      rows[(size_t)m * L + l] = r;
    }
  }
  LayerRow* d_rows;
  cudaMalloc(&d_rows, sizeof(LayerRow) * rows.size());
  cudaMemcpy(d_rows, rows.data(), sizeof(LayerRow) * rows.size(), cudaMemcpyHostToDevice);
  printf("# L = %d rows, %d evaluations per lane, one warp per CTA, %d model columns per warp\n", L, rounds, S);
  std::vector<double> ref;
  run<2, 0>(d_rows, nmodels, L, rounds, &ref, "rayleigh secular_rec");
  run<2, 6>(d_rows, nmodels, L, rounds, &ref, "rayleigh two candidates per lane");
  run<2, 4>(d_rows, nmodels, L, rounds, &ref, "rayleigh rotated");
  if (L == 6) run<2, 5>(d_rows, nmodels, L, rounds, &ref, "rayleigh unrolled<6>");
  if (argc > 3 || getenv("MB_ONLY")) {
  run<2, 2>(d_rows, nmodels, L, rounds, &ref, "rayleigh 2pass 1 layer");
  run<2, 1>(d_rows, nmodels, L, rounds, &ref, "rayleigh 2pass 2 layers");
  run<2, 3>(d_rows, nmodels, L, rounds, &ref, "rayleigh 2pass 3 layers");
  }
  ref.clear();
  run<1, 0>(d_rows, nmodels, L, rounds, &ref, "love secular_rec");
  run<1, 6>(d_rows, nmodels, L, rounds, &ref, "love two candidates per lane");
  run<1, 4>(d_rows, nmodels, L, rounds, &ref, "love 2 per iteration");
  run<1, 5>(d_rows, nmodels, L, rounds, &ref, "love 3 per iteration");
  if (argc > 3 || getenv("MB_ONLY")) {
  run<1, 1>(d_rows, nmodels, L, rounds, &ref, "love 2pass 2 layers");
  run<1, 2>(d_rows, nmodels, L, rounds, &ref, "love 2pass 3 layers");
  run<1, 3>(d_rows, nmodels, L, rounds, &ref, "love 2pass 6 layers");
  }
  return 0;
}
