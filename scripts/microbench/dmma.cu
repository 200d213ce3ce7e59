// fp64 tensor-core probe on B200 (sm_100a): mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 (SASS: DMMA)
// dependent-issue latency and throughput per SM, next to DFMA (scripts/microbench/dfma.cu) — the numbers that
// decide whether the reduced pose solve's 6x6 block products belong on the tensor pipe (north_star names it;
// VERDICT r1 item 1a).  One CTA on one SM, clock64() around unrolled chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// K independent accumulator pairs per warp; every DMMA of a chain depends on the previous one through D
template <int K>
__global__ void k_dmma(double *out, long long *cyc, double a, double b) {
  double d0[K], d1[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { d0[i] = a + i + threadIdx.x; d1[i] = b - i; }
  const double fa = a * (1 + (threadIdx.x & 3)), fb = b * (1 + (threadIdx.x >> 2));
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < K; ++i) dmma(d0[i], d1[i], fa, fb);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += d0[i] + d1[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// the shape the solver would use: a 6x6x6 block product padded to 8x8x8 = two dependent DMMAs, operands from
// shared memory (one LDS.64 per operand and lane), K independent products per warp in flight
template <int K>
__global__ void k_block_product(double *out, long long *cyc, const double *src) {
  __shared__ double sA[64 * 8], sB[64 * 8];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) { sA[i] = src[i]; sB[i] = src[512 + i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, row = lane >> 2, kk = lane & 3;
  double d0[K], d1[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { d0[i] = 0.0; d1[i] = 0.0; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const double *A = sA + 64 * ((i + it) & 7), *B = sB + 64 * ((i + 3 * it) & 7);
      dmma(d0[i], d1[i], A[8 * row + kk], B[8 * row + kk]);
      dmma(d0[i], d1[i], A[8 * row + 4 + kk], B[8 * row + 4 + kk]);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += d0[i] + d1[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int K> void run(double *out, long long *cyc) {
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) { k_dmma<K><<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); }
    const double per = cyc[0] / (64.0 * 4 * K);
    // one m8n8k4 DMMA = 8 * 8 * 4 = 256 fp64 FMAs
    printf("DMMA  K=%2d chains  warps=%2d : %6.2f cycles per DMMA per warp -> %6.1f fp64 FMA/clk/SM (DFMA peak: 64)\n", K, warps, per, 256.0 * warps / per);
  }
}
template <int K> void run_block(double *out, long long *cyc, const double *src) {
  for (int warps : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { k_block_product<K><<<1, 32 * warps>>>(out, cyc, src); cudaDeviceSynchronize(); }
    const double per = cyc[0] / (64.0 * K);
    printf("6x6x6 block product as 2 DMMA from shared memory, K=%d in flight, warps=%2d : %6.1f cycles per product per warp -> %6.1f useful FMA/clk/SM (216 per product)\n",
           K, warps, per, 216.0 * warps / per);
  }
}

int main() {
  double *out, *src; long long *cyc;
  cudaMalloc(&out, 2048 * 8); cudaMallocManaged(&cyc, 64); cudaMallocManaged(&src, 1024 * 8);
  for (int i = 0; i < 1024; ++i) src[i] = 1.0 / (1 + i);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1e3);
  run<1>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc);
  run_block<1>(out, cyc, src); run_block<4>(out, cyc, src);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
