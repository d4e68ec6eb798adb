// ffma_forms.cu -- developer micro-benchmark: FP32 FMA issue rate on sm_100a by operand form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_forms tools/ffma_forms.cu && /tmp/ffma_forms
// A: x = fma(x, c, c)          one register operand (constant/uniform multiplicand and addend)
// B: x_k = fma(x_k, y_k, z_k)  three distinct register operands per instruction
// C: t_k = fma(s, y_k, t_k)    16 chains sharing one multiplicand (the pair-test pattern of edge.cu)
// D: like C plus one FMNMX per 2 FFMA
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__global__ void __launch_bounds__(256) kA(float* out, float a, float b)
{
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 1e-3f + k;
  for (int i = 0; i < ITERS; i++)
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fmaf(x[k], a, b);
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 1.2345f) out[0] = s;
}

__global__ void __launch_bounds__(256) kB(float* out, const float* in)
{
  float x[8], y[8], z[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
  {
    x[k] = in[threadIdx.x + k];
    y[k] = in[threadIdx.x + 8 + k];
    z[k] = in[threadIdx.x + 16 + k];
  }
  for (int i = 0; i < ITERS; i++)
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fmaf(x[k], y[k], z[k]);
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 1.2345f) out[0] = s;
}

__global__ void __launch_bounds__(128) kC(float* out, const float* in, const float4* obs, int nobs)
{
  __shared__ float4 so[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) so[i] = obs[i % nobs];
  __syncthreads();
  float cx[16], cy[16], cz[16], cr[16], B[16];
#pragma unroll
  for (int k = 0; k < 16; k++)
  {
    cx[k] = in[threadIdx.x + k];
    cy[k] = in[threadIdx.x + 16 + k];
    cz[k] = in[threadIdx.x + 32 + k];
    cr[k] = in[threadIdx.x + 48 + k];
    B[k] = in[threadIdx.x + 64 + k];
  }
  float acc = 0.f;
  for (int i = 0; i < ITERS / 8; i++)
    for (int o = 0; o < 32; o++)
    {
      const float4 ob = so[(i * 32 + o) & 1023];
#pragma unroll
      for (int k = 0; k < 16; k++)
      {
        float t = fmaf(ob.x, cx[k], B[k]);
        t = fmaf(ob.y, cy[k], t);
        t = fmaf(ob.z, cz[k], t);
        t = fmaf(ob.w, cr[k], t);
        acc += t;  // one extra FADD per 4 FFMA keeps the chains alive without min/compare
      }
    }
  if (acc == 1.2345f) out[0] = acc;
}

__global__ void __launch_bounds__(128) kD(float* out, const float* in, const float4* obs, int nobs)
{
  __shared__ float4 so[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) so[i] = obs[i % nobs];
  __syncthreads();
  float cx[16], cy[16], cz[16], cr[16], B[16];
#pragma unroll
  for (int k = 0; k < 16; k++)
  {
    cx[k] = in[threadIdx.x + k];
    cy[k] = in[threadIdx.x + 16 + k];
    cz[k] = in[threadIdx.x + 32 + k];
    cr[k] = in[threadIdx.x + 48 + k];
    B[k] = in[threadIdx.x + 64 + k];
  }
  bool hit = false;
  for (int i = 0; i < ITERS / 8; i++)
    for (int o = 0; o < 32; o++)
    {
      const float4 ob = so[(i * 32 + o) & 1023];
      float m = 3e38f;
#pragma unroll
      for (int k = 0; k < 16; k++)
      {
        float t = fmaf(ob.x, cx[k], B[k]);
        t = fmaf(ob.y, cy[k], t);
        t = fmaf(ob.z, cz[k], t);
        t = fmaf(ob.w, cr[k], t);
        m = fminf(m, t);
      }
      hit |= m < -1e30f;
    }
  if (hit) out[0] = 1.f;
}

// E: pattern D with packed FFMA2 (fma.rn.f32x2): two spheres per instruction
__global__ void __launch_bounds__(128) kE(float* out, const float* in, const float4* obs, int nobs)
{
  __shared__ float4 so[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) so[i] = obs[i % nobs];
  __syncthreads();
  float2 cx[8], cy[8], cz[8], cr[8], B[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
  {
    cx[k] = make_float2(in[threadIdx.x + k], in[threadIdx.x + 100 + k]);
    cy[k] = make_float2(in[threadIdx.x + 16 + k], in[threadIdx.x + 116 + k]);
    cz[k] = make_float2(in[threadIdx.x + 32 + k], in[threadIdx.x + 132 + k]);
    cr[k] = make_float2(in[threadIdx.x + 48 + k], in[threadIdx.x + 148 + k]);
    B[k] = make_float2(in[threadIdx.x + 64 + k], in[threadIdx.x + 164 + k]);
  }
  bool hit = false;
  for (int i = 0; i < ITERS / 8; i++)
    for (int o = 0; o < 32; o++)
    {
      const float4 ob = so[(i * 32 + o) & 1023];
      const float2 ox = make_float2(ob.x, ob.x), oy = make_float2(ob.y, ob.y), oz = make_float2(ob.z, ob.z),
                   ow = make_float2(ob.w, ob.w);
      float m = 3e38f;
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        float2 t = __ffma2_rn(ox, cx[k], B[k]);
        t = __ffma2_rn(oy, cy[k], t);
        t = __ffma2_rn(oz, cz[k], t);
        t = __ffma2_rn(ow, cr[k], t);
        m = fminf(m, fminf(t.x, t.y));
      }
      hit |= m < -1e30f;
    }
  if (hit) out[0] = 1.f;
}

int main()
{
  float *out, *in;
  float4* obs;
  cudaMalloc(&out, 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&obs, 1024 * 16);
  cudaMemset(in, 0, 4096 * 4);
  cudaMemset(obs, 0, 1024 * 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int variant = 0; variant < 5; variant++)
    for (int wps : { 1, 2, 4, 8 })  // warps per scheduler
    {
      const int threads = (variant < 2) ? 256 : 128;
      const int blocks = sms * (wps * 4 * 32 / threads);
      if (blocks < sms) continue;
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++)
      {
        cudaEventRecord(e0);
        if (variant == 0) kA<<<blocks, threads>>>(out, 1.0000001f, 1e-7f);
        if (variant == 1) kB<<<blocks, threads>>>(out, in);
        if (variant == 2) kC<<<blocks, threads>>>(out, in, obs, 1000);
        if (variant == 3) kD<<<blocks, threads>>>(out, in, obs, 1000);
        if (variant == 4) kE<<<blocks, threads>>>(out, in, obs, 1000);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      double fmas;
      if (variant < 2) fmas = (double)blocks * threads * ITERS * 32.0;
      else fmas = (double)blocks * threads * (ITERS / 8) * 32.0 * 64.0;
      printf("variant %c  warps/scheduler %d  blocks %d  %.3f ms  %.2f TFLOP/s (FMA only)  %.3f FMA-inst/clk/SMSP @1.9GHz\n",
             "ABCDE"[variant], wps, blocks, best, 2.0 * fmas / best / 1e9, fmas / 32.0 / (best * 1e-3) / (sms * 4.0) / 1.9e9);
    }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
