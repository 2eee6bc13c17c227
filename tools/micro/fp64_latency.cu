// Microbenchmark (developer tool): dependent-chain latency and multi-chain throughput of the fp64
// instructions the dispersion kernel is made of, one warp per SM sub-partition at a time.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, double a, double b, int iters, long long* cyc) {
  double x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = a + i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], b, a);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dadd_chain(double* out, double a, int iters, long long* cyc) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = x + a;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void rsq_chain(double* out, double a, int iters, long long* cyc) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      double y;
      asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      x = y;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void sel_chain(double* out, double a, int iters, long long* cyc) {
  double x = a + threadIdx.x, y = a * 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) { double z = (x < y) ? y + 1.0 : x; y = x; x = z; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lds_chain(double* out, int iters, long long* cyc) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = nxt[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void imad_chain(double* out, int a, int iters, long long* cyc) {
  int x = a + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = x * a + 7;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void ffma_chain(double* out, float a, int iters, long long* cyc) {
  float x = a + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = fmaf(x, a, 0.5f);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000; const double n = iters * 16.0;
#define RUN(name, launch, per) launch; cudaDeviceSynchronize(); launch; cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-44s %8.2f cycles per step (%s)\n", name, h / n, per);
  RUN("DFMA dependent, 1 warp, 1 chain", (dfma_chain<1><<<1, 32>>>(out, 1.0, 0.999, iters, cyc)), "latency");
  RUN("DFMA 1 warp, 2 chains", (dfma_chain<2><<<1, 32>>>(out, 1.0, 0.999, iters, cyc)), "per 2 DFMA");
  RUN("DFMA 1 warp, 4 chains", (dfma_chain<4><<<1, 32>>>(out, 1.0, 0.999, iters, cyc)), "per 4 DFMA");
  RUN("DFMA 1 warp, 8 chains", (dfma_chain<8><<<1, 32>>>(out, 1.0, 0.999, iters, cyc)), "per 8 DFMA");
  RUN("DFMA 4 warps (1/SMSP), 8 chains", (dfma_chain<8><<<1, 128>>>(out, 1.0, 0.999, iters, cyc)), "per 8 DFMA");
  RUN("DFMA 8 warps (2/SMSP), 8 chains", (dfma_chain<8><<<1, 256>>>(out, 1.0, 0.999, iters, cyc)), "per 8 DFMA, 2 warps share");
  RUN("DFMA 16 warps (4/SMSP), 1 chain", (dfma_chain<1><<<1, 512>>>(out, 1.0, 0.999, iters, cyc)), "per DFMA, 4 warps share");
  RUN("DADD dependent", (dadd_chain<<<1, 32>>>(out, 1.0, iters, cyc)), "latency");
  RUN("MUFU.RSQ64H dependent", (rsq_chain<<<1, 32>>>(out, 1.5, iters, cyc)), "latency incl. moves");
  RUN("DSETP+FSELx2+DADD dependent", (sel_chain<<<1, 32>>>(out, 1.0, iters, cyc)), "latency");
  RUN("LDS dependent (pointer chase)", (lds_chain<<<1, 32>>>(out, iters, cyc)), "latency");
  RUN("IMAD dependent", (imad_chain<<<1, 32>>>(out, 3, iters, cyc)), "latency");
  RUN("FFMA dependent", (ffma_chain<<<1, 32>>>(out, 0.999f, iters, cyc)), "latency");
  return 0;
}
