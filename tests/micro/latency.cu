// Micro-benchmark (GPU box only): single-warp latency (cycles per operation in a dependent chain)
// and full-chip throughput of the field / point primitives. Build:
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr latency.cu -o latency
#include <cstdio>
#include <cuda_runtime.h>
#include "../../blitzar_b200/csrc/curve.cuh"
using namespace b200;

template <int OP> __global__ void k_lat(Fe<8>* io, long long* cycles, int iters) {
  typedef F25519 F;
  Fe<8> a = io[threadIdx.x & 3], b = io[4 + (threadIdx.x & 3)];
  Ed25519::Point P;
  P.X = a; P.Y = b; P.Z = io[1]; P.T = io[2];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) F::mul(a, a, b);
    if (OP == 1) F::mul_lat(a, a, b);
    if (OP == 2) F::mul_ref(a, a, b);
    if (OP == 3) F::add(a, a, b);
    if (OP == 4) F::sub(a, a, b);
    if (OP == 5) Ed25519::dbl<SeqExec>(P, P);
    if (OP == 6) Ed25519::dbl<QuadExec>(P, P);
    if (OP == 7) Ed25519::add<SeqExec>(P, P, P);
    if (OP == 8) Ed25519::add<QuadExec>(P, P, P);
    if (OP == 9) Ed25519::dbl<QuadExecConv>(P, P);
    if (OP == 10) Ed25519::add<QuadExecConv>(P, P, P);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  io[8 + (threadIdx.x & 3)] = (OP >= 5) ? P.X : a;
}
template <int OP> __global__ void k_thr(Fe<8>* io, int iters) {
  typedef F25519 F;
  Fe<8> a = io[threadIdx.x & 3], b = io[4 + (threadIdx.x & 3)];
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) F::mul(a, a, b);
    if (OP == 1) F::mul_lat(a, a, b);
    if (OP == 2) F::mul_ref(a, a, b);
  }
  if (a.l[0] == 0x12345678u) io[8] = a;
}
template <int OP> void lat(const char* name, Fe<8>* d, long long* dc, int threads) {
  const int iters = 200;
  k_lat<OP><<<1, threads>>>(d, dc, iters);
  k_lat<OP><<<1, threads>>>(d, dc, iters);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("latency  %-28s %4d thr: %8.1f cycles/op\n", name, threads, (double)c / iters);
}
template <int OP> void thr(const char* name, Fe<8>* d) {
  const int iters = 2000, blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_thr<OP><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e0);
  k_thr<OP><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * threads * iters;
  printf("throughput %-26s: %.2f G mul/s (%.1f SM-cycles per warp-mul at 1.9 GHz)\n", name, muls / ms * 1e-6,
         1.9e9 * 148 / (muls / 32 / (ms * 1e-3)));
}
int main() {
  Fe<8>* d; long long* dc;
  cudaMalloc(&d, 64 * sizeof(Fe<8>)); cudaMalloc(&dc, 8);
  Fe<8> h[16];
  for (int i = 0; i < 16; ++i) for (int k = 0; k < 8; ++k) h[i].l[k] = 0x9e3779b9u * (i * 8 + k + 1);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  lat<0>("F25519::mul (chains)", d, dc, 32);
  lat<1>("F25519::mul_lat (columns)", d, dc, 32);
  lat<2>("F25519::mul_ref (plain C)", d, dc, 32);
  lat<3>("F25519::add", d, dc, 32);
  lat<4>("F25519::sub", d, dc, 32);
  lat<5>("Ed25519::dbl seq", d, dc, 32);
  lat<6>("Ed25519::dbl quad", d, dc, 32);
  lat<7>("Ed25519::add seq", d, dc, 32);
  lat<8>("Ed25519::add quad", d, dc, 32);
  lat<9>("Ed25519::dbl quad fullmask", d, dc, 32);
  lat<10>("Ed25519::add quad fullmask", d, dc, 32);
  thr<0>("F25519::mul (chains)", d);
  thr<1>("F25519::mul_lat (columns)", d);
  thr<2>("F25519::mul_ref (plain C)", d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
