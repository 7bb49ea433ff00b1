// Micro-benchmark: cost of one grid-wide barrier with G co-resident CTAs (cooperative launch), three implementations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gbar tools/micro/grid_barrier_bench.cu && /tmp/gbar
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void bar_atomic(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}
// red.release + ld.acquire polling with backoff
__device__ __forceinline__ void bar_relacq(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
// flags: every CTA publishes its epoch in its own slot (stride 32 words), CTA 0's warp 0 polls all slots and publishes the
// release epoch that everyone else polls
__device__ __forceinline__ void bar_flags(unsigned* flags, unsigned* release, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (blockIdx.x == 0) {
    if (threadIdx.x < 32) {
      for (unsigned c = 1 + threadIdx.x; c < gridDim.x; c += 32) {
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32 * c) : "memory"); } while (v < epoch);
      }
      __syncwarp();
      if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(release), "r"(epoch) : "memory");
    }
  } else if (threadIdx.x == 0) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32 * blockIdx.x), "r"(epoch) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(release) : "memory"); } while (v < epoch);
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(unsigned* mem, int iters, float* sink) {
  unsigned t = 0;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) bar_atomic(mem, t);
    else if (MODE == 1) bar_relacq(mem, t);
    else bar_flags(mem + 64, mem + 32, t);
    acc += (float)i;
  }
  if (threadIdx.x == 0) sink[blockIdx.x] = acc;
}

template <int MODE>
void run(int G, const char* name) {
  unsigned* mem; float* sink;
  cudaMalloc(&mem, 64 * 1024); cudaMalloc(&sink, 4096);
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(mem, 0, 64 * 1024);
    void* args[] = {&mem, (void*)&iters, &sink};
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(G), dim3(256), args, 100 * 1024, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("%-28s G=%3d: %.2f us per barrier (%s)\n", name, G, ms * 1e3 / iters, cudaGetErrorString(err));
  }
  cudaFree(mem); cudaFree(sink);
}

int main() {
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int G : {32, 64, 128}) {
    run<0>(G, "atomicAdd + volatile poll");
    run<1>(G, "red.release + ld.acquire");
    run<2>(G, "flags + master");
  }
  return 0;
}
