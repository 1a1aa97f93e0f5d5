// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) issue/throughput.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
// PACK: 0 scalar FFMA, 1 FFMA2.  KI: independent integer ALU ops per FMA pair.  KL: 1 = one LDS per 2 pairs
template <int PACK, int KI, int KL>
__global__ void __launch_bounds__(256) k(float* out, float s, int extra) {
  __shared__ float sm[256];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  float2 a[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); x[i] = threadIdx.x + i; }
  const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f, 0.125f);
  const unsigned y = extra;
  float l = 0.f;
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACK == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
      else a[i] = __ffma2_rn(a[i], m, c);
      if (KI >= 1) x[i] = (x[i] ^ y) + 0x9e37u;      // LOP3 + IADD
      if (KI >= 2) x[i] = (x[i] & 0xffffffu) + y;
      if (KL && (i & 1)) l += sm[(x[i] + it) & 255];
    }
  }
  float r = l;
  unsigned xs = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { r += a[i].x + a[i].y; xs += x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)xs;
}
template <int PACK, int KI, int KL> void run(const char* name, float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8;
  k<PACK, KI, KL><<<grid, 256>>>(d, 1.0001f, 3);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<PACK, KI, KL><<<grid, 256>>>(d, 1.0001f, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double fma = (double)grid * 256 * N_IT * 16;
  printf("%-34s %.3f ms  %.1f FMA/clk/SM  (%.1f SMSP-cycles per 16 FMA warp-instr)\n", name, ms,
         fma / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (8.0 * 256 / 32 / 4 * N_IT));
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0, 0, 0>("FFMA", d);
  run<1, 0, 0>("FFMA2", d);
  run<0, 1, 0>("FFMA  + 2 int per pair", d);
  run<1, 1, 0>("FFMA2 + 2 int per pair", d);
  run<0, 2, 0>("FFMA  + 4 int per pair", d);
  run<1, 2, 0>("FFMA2 + 4 int per pair", d);
  run<0, 1, 1>("FFMA  + 2 int per pair + LDS/2", d);
  run<1, 1, 1>("FFMA2 + 2 int per pair + LDS/2", d);
  return 0;
}
