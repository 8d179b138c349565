// FP64 pipe micro-benchmark for sm_100a: how much of the DFMA peak do dependent chains reach at the occupancies the
// ERI kernels run at (4 and 6 warps per scheduler), with register operands and a DFMA/DMUL/DADD mix?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <int ILP, int MIX>
__global__ void __launch_bounds__(256) k(int iters, const double* __restrict__ in, double* __restrict__ out) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = in[threadIdx.x + 256 * i]; y[i] = in[threadIdx.x + 256 * (i + 8)]; z[i] = in[threadIdx.x + 256 * (i + 16)]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MIX == 0) x[i] = fma(x[i], y[i], z[i]);
        else if (MIX == 1) { if (r & 1) x[i] = x[i] * y[i]; else x[i] = x[i] + z[i]; }
        else if (MIX == 3) x[i] = fma(x[i], y[i], 1.0000000001);                 // two register operands + immediate
        else if (MIX == 4) x[i] = fma(x[i], 1.0000000001, 1e-12);                // one register operand
        else if (MIX == 5) x[i] = fma(y[0], z[0], x[i]);                         // two operands shared by consecutive instructions (.reuse)
        else { if (r == 0) x[i] = fma(x[i], y[i], z[i]); else if (r == 1) x[i] = x[i] * y[i]; else if (r == 2) x[i] = x[i] + z[i]; else x[i] = fma(y[i], z[i], x[i]); }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP, int MIX>
static void run(int ctas_per_sm, int sms, double* in, double* out) {
  const int iters = 20000;
  // occupancy control: dynamic shared memory so that exactly ctas_per_sm CTAs of 256 threads fit
  const int smem = (220 * 1024) / ctas_per_sm - 1024;
  cudaFuncSetAttribute(k<ILP, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int blocks = sms * ctas_per_sm;
  k<ILP, MIX><<<blocks, 256, smem>>>(100, in, out);
  cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<ILP, MIX><<<blocks, 256, smem>>>(iters, in, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double inst = (double)blocks * 8 /*warps*/ * iters * 4.0 * ILP;   // FP64 warp instructions
  const double peak = sms * 4.0 * 0.5 * 1.965e9;                            // 1 warp instruction / 2 cycles / scheduler
  printf("ILP %d mix %d warps/scheduler %2d : %6.2f G warp-inst/s = %5.1f %% of the FP64 issue peak\n", ILP, MIX, ctas_per_sm * 2,
         inst / (ms * 1e-3) * 1e-9, 100.0 * inst / (ms * 1e-3) / peak);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  std::vector<double> h(256 * 24, 1.0000001);
  double *in, *out; cudaMalloc(&in, h.size() * 8); cudaMalloc(&out, 8);
  cudaMemcpy(in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  for (int c : {2, 8}) {
    run<1, 0>(c, p.multiProcessorCount, in, out); run<2, 0>(c, p.multiProcessorCount, in, out);
    run<4, 0>(c, p.multiProcessorCount, in, out); run<8, 0>(c, p.multiProcessorCount, in, out);
    run<1, 2>(c, p.multiProcessorCount, in, out); run<2, 2>(c, p.multiProcessorCount, in, out); run<4, 2>(c, p.multiProcessorCount, in, out);
    run<2, 1>(c, p.multiProcessorCount, in, out);
    run<4, 3>(c, p.multiProcessorCount, in, out); run<4, 4>(c, p.multiProcessorCount, in, out); run<4, 5>(c, p.multiProcessorCount, in, out);
  }
  return 0;
}
