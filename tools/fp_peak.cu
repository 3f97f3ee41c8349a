// fp_peak.cu -- measures the FP64 / FP32 FMA issue peaks of the GPU (the roofline
// denominators of the interpreter kernel; MEASURED_PEAKS.json only has HBM and bf16).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_peak tools/fp_peak.cu
// Run on the GPU box: ./fp_peak > profiles/fp_peaks.json
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void fma_chain(T* out, int iters, T a, T b) {
  T x0 = a + threadIdx.x, x1 = a * 2, x2 = a * 3, x3 = a * 4, x4 = a * 5, x5 = a * 6, x6 = a * 7, x7 = a * 8;
  for (int i = 0; i < iters; ++i) {
    x0 = x0 * b + a; x1 = x1 * b + a; x2 = x2 * b + a; x3 = x3 * b + a;
    x4 = x4 * b + a; x5 = x5 * b + a; x6 = x6 * b + a; x7 = x7 * b + a;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <typename T>
double measure(int sms) {
  const int threads = 512, blocks = sms * 4, iters = 1 << 16;
  T* out;
  cudaMalloc(&out, sizeof(T) * threads * blocks);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    fma_chain<T><<<blocks, threads>>>(out, iters, (T)1.0001, (T)0.9999);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * 8 * (double)iters * threads * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaFree(out);
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double f64 = measure<double>(p.multiProcessorCount);
  const double f32 = measure<float>(p.multiProcessorCount);
  std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_tflops\": %.3f, \"fp32_tflops\": %.3f, "
              "\"how\": \"8 independent FMA chains per thread, 512 threads x 4 CTAs per SM, best of 4 timed "
              "launches, CUDA events (tools/fp_peak.cu)\"}\n",
              p.name, p.multiProcessorCount, f64, f32);
  return 0;
}
