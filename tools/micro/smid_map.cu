// Developer probe: how does the CTA scheduler map blockIdx.x -> SM for a fully resident grid of
// 1-warp CTAs?  (Is smid a function of blockIdx.x mod 148?)
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void probe(int* smid, long long spin) {
  extern __shared__ double sm[];
  unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  if (threadIdx.x == 0) smid[blockIdx.x] = (int)id;
  long long t0 = clock64();
  while (clock64() - t0 < spin) { sm[threadIdx.x] += 1.0; }
}
int main() {
  for (int grid : {1536, 1776, 2048}) {
    int* d; cudaMalloc(&d, grid * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    probe<<<grid, 32, 17728>>>(d, 2000000);
    cudaDeviceSynchronize();
    std::vector<int> h(grid); cudaMemcpy(h.data(), d, grid * sizeof(int), cudaMemcpyDeviceToHost);
    int consistent = 0, total = 0, maxsm = 0;
    std::vector<int> cnt(256, 0);
    for (int i = 0; i < grid; ++i) { cnt[h[i]]++; if (h[i] > maxsm) maxsm = h[i]; if (i >= 148) { total++; consistent += (h[i] == h[i - 148]); } }
    int mn = 1 << 30, mx = 0; for (int s = 0; s <= maxsm; ++s) { if (cnt[s] < mn) mn = cnt[s]; if (cnt[s] > mx) mx = cnt[s]; }
    printf("grid %d: max smid %d, CTAs per SM min %d max %d, smid[i]==smid[i-148] for %d of %d\n", grid, maxsm, mn, mx, consistent, total);
    printf("  first 20: "); for (int i = 0; i < 20; ++i) printf("%d ", h[i]); printf("\n  148..167: "); for (int i = 148; i < 168; ++i) printf("%d ", h[i]); printf("\n");
    cudaFree(d);
  }
  return 0;
}
