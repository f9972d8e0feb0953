// tools/ubench/fp32x2_probe.cu -- issue rate of the packed single-precision instructions of sm_100a
// (FADD2 / FMUL2 / FFMA2) against their scalar forms, alone and mixed with integer work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o variants/fp32x2_probe tools/ubench/fp32x2_probe.cu   (variants/ ships to the GPU box)
// Prints warp-instructions per clock per SM sub-partition for each loop body.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, long long* clk, float seed) {
  float2 a[8];
  unsigned u[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  const float2 c = make_float2(seed * 0.5f, seed * 0.25f);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 16 scalar FADD
        a[i].x += c.x;
        a[i].y += c.y;
      } else if (MODE == 1) {  // 8 FADD2
        a[i] = __fadd2_rn(a[i], c);
      } else if (MODE == 2) {  // 16 scalar FFMA
        a[i].x = fmaf(a[i].x, c.x, c.y);
        a[i].y = fmaf(a[i].y, c.y, c.x);
      } else if (MODE == 3) {  // 8 FFMA2
        a[i] = __ffma2_rn(a[i], c, c);
      } else if (MODE == 4) {  // 16 scalar FADD + 8 integer LOP3/IADD
        a[i].x += c.x;
        a[i].y += c.y;
        u[i & 3] = (u[i & 3] ^ 0x9e3779b9u) + (u[(i + 1) & 3] >> 3);
      } else if (MODE == 5) {  // 8 FADD2 + 8 integer
        a[i] = __fadd2_rn(a[i], c);
        u[i & 3] = (u[i & 3] ^ 0x9e3779b9u) + (u[(i + 1) & 3] >> 3);
      } else if (MODE == 6) {  // 8 FADD2 + 8 scalar FFMA (mixed butterflies)
        a[i] = __fadd2_rn(a[i], c);
        a[(i + 4) & 7].x = fmaf(a[(i + 4) & 7].x, c.x, c.y);
      } else if (MODE == 7) {  // 16 scalar FADD + 8 scalar FFMA (the scalar equivalent of 6)
        a[i].x += c.x;
        a[i].y += c.y;
        a[(i + 4) & 7].x = fmaf(a[(i + 4) & 7].x, c.x, c.y);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] + u[1] + u[2] + u[3]);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int ctas_per_sm) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * ctas_per_sm;
  float* out;
  long long* clk;
  cudaMalloc(&out, sizeof(float) * grid * 256);
  cudaMalloc(&clk, sizeof(long long) * grid);
  probe<MODE><<<grid, 256>>>(out, clk, 1.0f);
  probe<MODE><<<grid, 256>>>(out, clk, 1.0f);
  cudaDeviceSynchronize();
  long long* h = new long long[grid];
  cudaMemcpy(h, clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += (double)h[i];
  avg /= grid;
  // warps per SM sub-partition: ctas_per_sm * 8 warps / 4
  const double winstr = (double)ITERS * instr_per_iter * (ctas_per_sm * 8 / 4.0);
  printf("%-44s ctas/SM %d: %7.0f clk, %5.3f warp-instr/clk/SMSP\n", name, ctas_per_sm, avg, winstr / avg);
  cudaFree(out);
  cudaFree(clk);
  delete[] h;
}

int main() {
  for (int c : {2, 4}) {
    run<0>("16 FADD", 16, c);
    run<1>("8 FADD2 (same flops)", 8, c);
    run<2>("16 FFMA", 16, c);
    run<3>("8 FFMA2 (same flops)", 8, c);
    run<4>("16 FADD + 16 int (LOP3, IADD/SHF)", 32, c);
    run<5>("8 FADD2 + 16 int", 24, c);
    run<7>("16 FADD + 8 FFMA", 24, c);
    run<6>("8 FADD2 + 8 FFMA", 16, c);
  }
  return 0;
}
