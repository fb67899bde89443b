// L2 microbenchmarks SURVEY.md §8d asks for: the memory-side yardsticks of an L2-resident working set on this GPU.
//   1. streaming read of a buffer that fits L2 (16 B loads, every SM, repeated passes) -> GB/s
//   2. red.global.add.{u32,u64} to pseudo-random cells of a 64 MB buffer (one RED per lane, all lanes different cells)
//      and the same with all 32 lanes of a warp on ONE cell (what an unaggregated ray step looks like) -> G RED/s
//   3. the same streaming read on a 4 GB buffer (HBM) for comparison
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l2_microbench tools/l2_microbench.cu
// Prints one JSON object.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                \
  do                                                                                         \
  {                                                                                          \
    cudaError_t e = (x);                                                                     \
    if (e != cudaSuccess)                                                                    \
    {                                                                                        \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                               \
    }                                                                                        \
  } while (0)

__global__ void __launch_bounds__(256) k_read(const uint4* __restrict__ p, const size_t n16, const int passes, unsigned* __restrict__ sink)
{
  unsigned acc = 0;
  for (int r = 0; r < passes; r++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u)
    *sink = acc;
}

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// MODE 0: every lane its own random cell; MODE 1: the 32 lanes of a warp share one random cell
template <class T, int MODE>
__global__ void __launch_bounds__(256) k_red(T* __restrict__ p, const uint32_t n_cells_mask, const int per_thread)
{
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s = MODE == 0 ? tid : (tid >> 5);
  for (int k = 0; k < per_thread; k++)
  {
    s = mix(s + 0x9e3779b9u * (uint32_t)(k + 1));
    T* a = p + (s & n_cells_mask);
    if (sizeof(T) == 4)
      asm volatile("red.global.add.u32 [%0], %1;" ::"l"(a), "r"(1u) : "memory");
    else
      asm volatile("red.global.add.u64 [%0], %1;" ::"l"(a), "l"(1ull) : "memory");
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b)
{
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  unsigned* sink;
  CK(cudaMalloc(&sink, 4));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d", prop.name, sms, prop.l2CacheSize);

  // ---- 1/3: streaming reads
  const size_t sizes_mb[] = {16, 32, 64, 96, 4096};
  for (size_t smb : sizes_mb)
  {
    const size_t bytes = smb << 20;
    uint4* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    const int passes = smb >= 1024 ? 2 : 40;
    const int grid = sms * 8;
    k_read<<<grid, 256>>>(buf, bytes / 16, 2, sink);  // warm-up: brings the buffer into L2
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++)
    {
      CK(cudaEventRecord(e0));
      k_read<<<grid, 256>>>(buf, bytes / 16, passes, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      const float ms = time_ms(e0, e1);
      if (ms < best) best = ms;
    }
    printf(", \"read_%zuMB_GBs\": %.1f", smb, (double)bytes * passes / (best * 1e-3) / 1e9);
    CK(cudaFree(buf));
  }

  // ---- 2: REDs into a 64 MB buffer
  {
    const size_t bytes = 64ull << 20;
    void* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    const int grid = sms * 16, per_thread = 256;
    const double n_red = (double)grid * 256 * per_thread;
    struct { const char* name; int kind; } cases[] = {{"red_u32_random_lane", 0}, {"red_u64_random_lane", 1}, {"red_u32_warp_same_cell", 2}, {"red_u64_warp_same_cell", 3}};
    for (auto& c : cases)
    {
      float best = 1e30f;
      for (int rep = 0; rep < 6; rep++)
      {
        CK(cudaEventRecord(e0));
        if (c.kind == 0) k_red<uint32_t, 0><<<grid, 256>>>((uint32_t*)buf, (uint32_t)(bytes / 4 - 1), per_thread);
        if (c.kind == 1) k_red<unsigned long long, 0><<<grid, 256>>>((unsigned long long*)buf, (uint32_t)(bytes / 8 - 1), per_thread);
        if (c.kind == 2) k_red<uint32_t, 1><<<grid, 256>>>((uint32_t*)buf, (uint32_t)(bytes / 4 - 1), per_thread);
        if (c.kind == 3) k_red<unsigned long long, 1><<<grid, 256>>>((unsigned long long*)buf, (uint32_t)(bytes / 8 - 1), per_thread);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const float ms = time_ms(e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      printf(", \"%s_Gred_s\": %.2f", c.name, n_red / (best * 1e-3) / 1e9);
    }
    CK(cudaFree(buf));
  }
  printf("}\n");
  return 0;
}
