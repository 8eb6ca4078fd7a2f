// Developer micro-benchmark: cost of a software grid barrier on B200 (cooperative launch).
// nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_barrier tools/bench_barrier.cu
// (build on the GPU box into /tmp or gpurun_out/: binaries are never kept in the tree)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void barrier_fence(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ULL);
    unsigned long long v;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < target);
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void barrier_release(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned long long v;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < target);
  }
  __syncthreads();
}
// relaxed polling (volatile) + one acquire fence at the end
__device__ __forceinline__ void barrier_relaxed(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned long long v;
    do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < target);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

template <int MODE>
__global__ void k(unsigned long long* ctr, unsigned long long base, int nb, double* data, int work) {
  unsigned long long target = base;
  for (int p = 0; p < nb; ++p) {
    if (work) {   // a dependent store + load like a phase would do
      data[(size_t)(p & 7) * 296 * 512 + (size_t)blockIdx.x * blockDim.x + threadIdx.x] = (double)p;
    }
    target += gridDim.x;
    if (MODE == 0) barrier_fence(ctr, target);
    else if (MODE == 1) barrier_release(ctr, target);
    else barrier_relaxed(ctr, target);
    if (work) {
      size_t j = ((size_t)(blockIdx.x + 1) % gridDim.x) * blockDim.x + threadIdx.x;
      double v = __ldcg(data + (size_t)(p & 7) * 296 * 512 + j);
      if (v != (double)p) atomicAdd(ctr + 8, 1ULL);
    }
  }
}

template <int MODE>
float run(int grid, int block, int nb, int work, unsigned long long* ctr, unsigned long long& base, double* data) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    void* args[] = {&ctr, &base, &nb, &data, &work};
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(grid), dim3(block), args, 0, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    base += (unsigned long long)nb * grid;
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return best;
}

int main() {
  unsigned long long* ctr; cudaMalloc(&ctr, 256); cudaMemset(ctr, 0, 256);
  double* data; cudaMalloc(&data, (size_t)8 * 296 * 512 * 8);
  unsigned long long base = 0;
  int nb = 64;
  for (int work = 0; work < 2; ++work)
    for (int grid : {148, 296})
      for (int block : {128, 512}) {
        float a = run<0>(grid, block, nb, work, ctr, base, data);
        float b = run<1>(grid, block, nb, work, ctr, base, data);
        float c = run<2>(grid, block, nb, work, ctr, base, data);
        float z = run<0>(grid, block, 0, work, ctr, base, data);
        printf("work %d grid %3d block %3d : per barrier  fence+atomic %.2f us   red.release %.2f us   relaxed-poll %.2f us   (empty launch %.1f us)\n",
               work, grid, block, (a - z) * 1e3 / nb, (b - z) * 1e3 / nb, (c - z) * 1e3 / nb, z * 1e3);
      }
  unsigned long long bad; cudaMemcpy(&bad, ctr + 8, 8, cudaMemcpyDeviceToHost); printf("mismatches %llu\n", bad);
  return 0;
}
