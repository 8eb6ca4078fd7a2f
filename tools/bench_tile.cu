// Developer micro-benchmark: latency anatomy of ONE top-level solve tile (16 warps: 32 strided panel loads per
// lane from cold HBM, three ld.cg gathers of data written just before by other SMs, shared-memory staging,
// 32 dependent FMAs).  nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_tile tools/bench_tile.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void gbar(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned long long v;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < target);
  }
  __syncthreads();
}

// mode bit 0: panel loads, bit 1: gathers, bit 2: gathers read data written by the same CTA instead of a neighbour
__global__ void __launch_bounds__(512, 1)
k(const double* __restrict__ P, size_t pstride_cta, int ld, double* w, unsigned long long* ctr, long long* out, int mode, int nrep, int nactive) {
  __shared__ double stage[16][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long target = 0;
  double acc = 0.0;
  for (int rep = 0; rep < nrep; ++rep) {
    // "previous phase": every CTA writes its slice of w
    w[(size_t)rep * gridDim.x * 512 + (size_t)blockIdx.x * 512 + threadIdx.x] = (double)(rep + 1);
    target += gridDim.x;
    gbar(ctr, target);
    long long t0 = clock64();
    const double* base = P + (size_t)rep * 40000003ull % (pstride_cta / 2) + (size_t)blockIdx.x * pstride_cta + (size_t)warp * 32 * ld + lane;
    double m[32];
    if ((mode & 1) && (int)blockIdx.x < nactive) {
#pragma unroll
      for (int t = 0; t < 32; ++t) m[t] = __ldg(base + (size_t)t * ld);
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) m[t] = 1.0;
    }
    long long t1 = clock64();
    double v = 0.0;
    if (mode & 2) {
      int src = (mode & 4) ? blockIdx.x : (blockIdx.x + 37) % gridDim.x;
      const double* wp = w + (size_t)rep * gridDim.x * 512 + (size_t)src * 512 + warp * 32 + lane;
      v = __ldcg(wp) + __ldcg(wp + 16 * 0) * 0.5 + __ldcg(w + (size_t)rep * gridDim.x * 512 + (size_t)((src + 11) % gridDim.x) * 512 + warp * 32 + lane);
    }
    stage[warp][lane] = v;
    __syncwarp();
    long long t2 = clock64();
#pragma unroll
    for (int t = 0; t < 32; ++t) acc = fma(m[t], stage[warp][t], acc);
    if (acc == 1.23456e-300) out[0] = 1;
    long long t3 = clock64();
    __syncthreads();
    long long t4 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      long long* o = out + 8 + rep * 8;
      o[0] = t1 - t0; o[1] = t2 - t1; o[2] = t3 - t2; o[3] = t4 - t3; o[4] = t4 - t0;
    }
  }
}

int main() {
  int ld = 751, grid = 148, nrep = 24;
  size_t pstride = (size_t)16 * 32 * ld + 4 * 1024 * 1024;          // doubles per CTA region
  double* P; cudaMalloc(&P, pstride * grid * 8 + (1u << 20));
  cudaMemset(P, 0, pstride * grid * 8);
  double* w; cudaMalloc(&w, (size_t)nrep * grid * 512 * 8);
  unsigned long long* ctr; cudaMalloc(&ctr, 256);
  long long* out; cudaMalloc(&out, (8 + nrep * 8) * 8);
  printf("panel buffer %.0f MB\n", pstride * grid * 8 / 1e6);
  const char* names[] = {"no loads", "panel only", "gather only (neighbour data)", "panel + gather (neighbour)", "", "", "gather only (own data)", "panel + gather (own)"};
  for (int nactive : {1, 8, 32, 148})
  for (int mode : {1, 3}) {
    cudaMemset(ctr, 0, 256);
    void* args[] = {&P, &pstride, &ld, &w, &ctr, &out, &mode, &nrep, &nactive};
    cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(512), args, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(8 + nrep * 8);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    double s[5] = {0, 0, 0, 0, 0};
    for (int r = 4; r < nrep; ++r) for (int q = 0; q < 5; ++q) s[q] += h[8 + r * 8 + q] / double(nrep - 4);
    printf("active %3d  %-32s issue %6.0f  gather+stage %6.0f  fma(+panel wait) %6.0f  sync %6.0f  total %6.0f cycles\n", nactive, names[mode], s[0], s[1], s[2], s[3], s[4]);
  }
  return 0;
}
