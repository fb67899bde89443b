// Throughput of the warp collectives the raycast aggregation is built from, per SM, with every warp slot busy:
// cycles per warp-instruction per SM = elapsed cycles / (iterations x warps per SM).  g = number of distinct keys in the warp.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/wcm tools/warp_collective_microbench.cu && /tmp/wcm
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(1024) k(int g, int* out)
{
  const unsigned lane = threadIdx.x & 31;
  int key = (int)(lane * (unsigned)g / 32u) * 977 + 13;  // g groups of contiguous lanes
  int q = (int)threadIdx.x;
  int accv = 0;
  const unsigned own_group = __match_any_sync(0xffffffffu, key);
#pragma unroll 1
  for (int it = 0; it < ITERS; it++)
  {
    if (OP == 0)
    {  // loop overhead + the dependent adds every variant carries
      accv += key;
    } else if (OP == 1)
    {
      unsigned m;
      asm volatile("match.any.sync.b32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(key));
      accv += (int)m;
    } else if (OP == 2)
    {
      unsigned m;
      int p;
      asm volatile("{ .reg .pred p; match.all.sync.b32 %0|p, %2, 0xffffffff; selp.s32 %1, 1, 0, p; }" : "=r"(m), "=r"(p) : "r"(key));
      accv += (int)m + p;
    } else if (OP == 3)
    {
      int s;
      asm volatile("redux.sync.add.s32 %0, %1, 0xffffffff;" : "=r"(s) : "r"(q));
      accv += s;
    } else if (OP == 4)
    {  // redux over the lanes of the own group (the compiler's path for non-uniform masks)
      accv += __reduce_add_sync(own_group, q);
    } else if (OP == 5)
    {
      accv += __shfl_sync(0xffffffffu, q, (it + lane) & 31);
    } else if (OP == 6)
    {
      accv += (int)__ballot_sync(0xffffffffu, (q + it) & 1);
    } else if (OP == 7)
    {  // match.any + redux per group = what the raycast loop does per step
      const unsigned m = __match_any_sync(0xffffffffu, key);
      accv += __reduce_add_sync(m, q);
    } else if (OP == 8)
    {  // group loop with full-mask collectives only: shfl + ballot + redux per distinct key
      unsigned rem = 0xffffffffu;
      while (rem)
      {
        const int src = __ffs(rem) - 1;
        const int k0 = __shfl_sync(0xffffffffu, key, src);
        const bool in = key == k0;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        int s;
        const int v = in ? q : 0;
        asm volatile("redux.sync.add.s32 %0, %1, 0xffffffff;" : "=r"(s) : "r"(v));
        if (in)
          accv += s;
        rem &= ~m;
      }
    } else if (OP == 9)
    {  // segmented sum by an inclusive scan (5 shfl.up) + two indexed shuffles; groups are lane-contiguous
      const int up = __shfl_up_sync(0xffffffffu, key, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || up != key);
      int s = q;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= (unsigned)o)
          s += t;
      }
      const unsigned below = heads & ((2u << lane) - 1u);
      const int first = 31 - __clz(below);
      const unsigned above = heads & ~((2u << lane) - 1u);
      const int last = above ? __ffs(above) - 2 : 31;
      const int hi = __shfl_sync(0xffffffffu, s, last);
      const int lo = __shfl_sync(0xffffffffu, s, first ? first - 1 : 0);
      accv += hi - (first ? lo : 0);
    }
    key += accv & 0;  // keeps the chain formally dependent without changing the keys
  }
  if (accv == 0x7fffffff)
    out[0] = accv;
}

template <int OP>
static float run(int g, int* d_out, int blocks)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<OP><<<blocks, 1024>>>(g, d_out);
  cudaEventRecord(a);
  k<OP><<<blocks, 1024>>>(g, d_out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main()
{
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  int* d_out;
  cudaMalloc(&d_out, 64);
  const int blocks = pr.multiProcessorCount * 2;  // 64 warps per SM
  const double warps_per_sm = 64.0;
  const char* names[10] = {"loop_overhead", "match_any", "match_all", "redux_full", "redux_group_masks", "shfl_idx", "ballot", "match_any+redux_group", "group_loop_fullmask",
                           "segmented_scan"};
  printf("{\"gpu\": \"%s\", \"sm_clock_khz\": %d, \"unit\": \"cycles per warp-iteration per SM (64 warps per SM resident), loop overhead included\"", pr.name, clk_khz);
  const int gs[6] = {1, 2, 3, 4, 8, 32};
  for (int gi = 0; gi < 6; gi++)
  {
    const int g = gs[gi];
    float ms[10];
    ms[0] = run<0>(g, d_out, blocks);
    ms[1] = run<1>(g, d_out, blocks);
    ms[2] = run<2>(g, d_out, blocks);
    ms[3] = run<3>(g, d_out, blocks);
    ms[4] = run<4>(g, d_out, blocks);
    ms[5] = run<5>(g, d_out, blocks);
    ms[6] = run<6>(g, d_out, blocks);
    ms[7] = run<7>(g, d_out, blocks);
    ms[8] = run<8>(g, d_out, blocks);
    ms[9] = run<9>(g, d_out, blocks);
    printf(",\n \"g%d\": {", g);
    for (int o = 0; o < 10; o++)
      printf("%s\"%s\": %.2f", o ? ", " : "", names[o], (double)ms[o] * 1e-3 * (double)clk_khz * 1e3 / ((double)ITERS * warps_per_sm));
    printf("}");
  }
  printf("}\n");
  return 0;
}
