// ffma_forms4.cu -- developer micro-benchmark #4: operand reuse and packed FFMA2 in the pair-test loop.
// All variants do the real inner loop's work (16 spheres x 4 fused multiply-adds + min tree + compare per obstacle,
// obstacles from shared memory) and differ in how the work is ordered / packed:
//   X0 scalar FFMA, one obstacle at a time                      (the round-1 kernel)
//   X1 FFMA2 (two spheres per register pair), one obstacle at a time
//   X2 scalar FFMA, two obstacles interleaved per sphere        (the sphere constants are reused by consecutive FFMAs)
//   X3 FFMA2, two obstacles interleaved per sphere pair
//   X4 scalar FFMA, four obstacles interleaved per sphere
//   X5 FFMA2, obstacle pair packed (two OBSTACLES per register pair, constants broadcast): 16 spheres x 4 FFMA2 per 2 obstacles
#include <cstdio>
#include <cuda_runtime.h>
#define OBS 1024
#define REPS 64

template <int V, int OCC>
__global__ void __launch_bounds__(128, OCC) kern(float* out, const float* in, const float4* obs)
{
  __shared__ float4 so[2 * OBS];  // (x,x,y,y) (z,z,r,r)
  __shared__ float sna[OBS];
  for (int i = threadIdx.x; i < OBS; i += blockDim.x)
  {
    const float4 o = obs[i];
    so[2 * i] = make_float4(o.x, o.x, o.y, o.y);
    so[2 * i + 1] = make_float4(o.z, o.z, o.w, o.w);
    sna[i] = o.w * o.w - (o.x * o.x + o.y * o.y + o.z * o.z);
  }
  __syncthreads();
  const int t = threadIdx.x;
  float cx[16], cy[16], cz[16], cr[16], cb[16];
#pragma unroll
  for (int k = 0; k < 16; k++)
  {
    cx[k] = in[t + k]; cy[k] = in[t + 16 + k]; cz[k] = in[t + 32 + k]; cr[k] = in[t + 48 + k]; cb[k] = in[t + 64 + k];
  }
  float2 px[8], py[8], pz[8], pr[8], pb[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
  {
    px[k] = make_float2(cx[2 * k], cx[2 * k + 1]); py[k] = make_float2(cy[2 * k], cy[2 * k + 1]);
    pz[k] = make_float2(cz[2 * k], cz[2 * k + 1]); pr[k] = make_float2(cr[2 * k], cr[2 * k + 1]);
    pb[k] = make_float2(cb[2 * k], cb[2 * k + 1]);
  }
  bool hit = false;
  for (int rep = 0; rep < REPS; rep++)
  {
    if (V == 0 || V == 1)
    {
#pragma unroll 2
      for (int o = 0; o < OBS; o++)
      {
        const float4 a = so[2 * o], b = so[2 * o + 1];
        const float na = sna[o];
        float m = 3e38f;
        if (V == 0)
        {
#pragma unroll
          for (int k = 0; k < 16; k++)
          {
            float v = fmaf(a.x, cx[k], cb[k]);
            v = fmaf(a.z, cy[k], v); v = fmaf(b.x, cz[k], v); v = fmaf(b.z, cr[k], v);
            m = fminf(m, v);
          }
        }
        else
        {
          const float2 ox = make_float2(a.x, a.y), oy = make_float2(a.z, a.w), oz = make_float2(b.x, b.y), ow = make_float2(b.z, b.w);
#pragma unroll
          for (int k = 0; k < 8; k++)
          {
            float2 v = __ffma2_rn(ox, px[k], pb[k]);
            v = __ffma2_rn(oy, py[k], v); v = __ffma2_rn(oz, pz[k], v); v = __ffma2_rn(ow, pr[k], v);
            m = fminf(m, fminf(v.x, v.y));
          }
        }
        hit |= (m < na);
      }
    }
    else if (V == 2 || V == 3)
    {
      for (int o = 0; o < OBS; o += 2)
      {
        const float4 a0 = so[2 * o], b0 = so[2 * o + 1], a1 = so[2 * o + 2], b1 = so[2 * o + 3];
        const float na0 = sna[o], na1 = sna[o + 1];
        float m0 = 3e38f, m1 = 3e38f;
        if (V == 2)
        {
#pragma unroll
          for (int k = 0; k < 16; k++)
          {
            float v0 = fmaf(a0.x, cx[k], cb[k]);
            float v1 = fmaf(a1.x, cx[k], cb[k]);
            v0 = fmaf(a0.z, cy[k], v0); v1 = fmaf(a1.z, cy[k], v1);
            v0 = fmaf(b0.x, cz[k], v0); v1 = fmaf(b1.x, cz[k], v1);
            v0 = fmaf(b0.z, cr[k], v0); v1 = fmaf(b1.z, cr[k], v1);
            m0 = fminf(m0, v0); m1 = fminf(m1, v1);
          }
        }
        else
        {
          const float2 ox0 = make_float2(a0.x, a0.y), oy0 = make_float2(a0.z, a0.w), oz0 = make_float2(b0.x, b0.y), ow0 = make_float2(b0.z, b0.w);
          const float2 ox1 = make_float2(a1.x, a1.y), oy1 = make_float2(a1.z, a1.w), oz1 = make_float2(b1.x, b1.y), ow1 = make_float2(b1.z, b1.w);
#pragma unroll
          for (int k = 0; k < 8; k++)
          {
            float2 v0 = __ffma2_rn(ox0, px[k], pb[k]);
            float2 v1 = __ffma2_rn(ox1, px[k], pb[k]);
            v0 = __ffma2_rn(oy0, py[k], v0); v1 = __ffma2_rn(oy1, py[k], v1);
            v0 = __ffma2_rn(oz0, pz[k], v0); v1 = __ffma2_rn(oz1, pz[k], v1);
            v0 = __ffma2_rn(ow0, pr[k], v0); v1 = __ffma2_rn(ow1, pr[k], v1);
            m0 = fminf(m0, fminf(v0.x, v0.y)); m1 = fminf(m1, fminf(v1.x, v1.y));
          }
        }
        hit |= (m0 < na0) | (m1 < na1);
      }
    }
    else if (V == 4)
    {
      for (int o = 0; o < OBS; o += 4)
      {
        float4 a[4], b[4];
        float na[4], m[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { a[j] = so[2 * (o + j)]; b[j] = so[2 * (o + j) + 1]; na[j] = sna[o + j]; m[j] = 3e38f; }
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; j++) v[j] = fmaf(a[j].x, cx[k], cb[k]);
#pragma unroll
          for (int j = 0; j < 4; j++) v[j] = fmaf(a[j].z, cy[k], v[j]);
#pragma unroll
          for (int j = 0; j < 4; j++) v[j] = fmaf(b[j].x, cz[k], v[j]);
#pragma unroll
          for (int j = 0; j < 4; j++) v[j] = fmaf(b[j].z, cr[k], v[j]);
#pragma unroll
          for (int j = 0; j < 4; j++) m[j] = fminf(m[j], v[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) hit |= (m[j] < na[j]);
      }
    }
    else if (V == 6)  // X5 with two obstacle pairs (four obstacles) per sphere constant
    {
      for (int o = 0; o < OBS; o += 4)
      {
        const float4 a0 = so[2 * o], b0 = so[2 * o + 1], a1 = so[2 * o + 2], b1 = so[2 * o + 3];
        const float4 a2 = so[2 * o + 4], b2 = so[2 * o + 5], a3 = so[2 * o + 6], b3 = so[2 * o + 7];
        const float2 oxA = make_float2(a0.x, a1.x), oyA = make_float2(a0.z, a1.z), ozA = make_float2(b0.x, b1.x), owA = make_float2(b0.z, b1.z);
        const float2 oxB = make_float2(a2.x, a3.x), oyB = make_float2(a2.z, a3.z), ozB = make_float2(b2.x, b3.x), owB = make_float2(b2.z, b3.z);
        float2 mA = make_float2(3e38f, 3e38f), mB = mA;
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
          const float2 c0 = make_float2(cx[k], cx[k]), c1 = make_float2(cy[k], cy[k]), c2 = make_float2(cz[k], cz[k]),
                       c3 = make_float2(cr[k], cr[k]), c4 = make_float2(cb[k], cb[k]);
          float2 vA = __ffma2_rn(oxA, c0, c4);
          float2 vB = __ffma2_rn(oxB, c0, c4);
          vA = __ffma2_rn(oyA, c1, vA); vB = __ffma2_rn(oyB, c1, vB);
          vA = __ffma2_rn(ozA, c2, vA); vB = __ffma2_rn(ozB, c2, vB);
          vA = __ffma2_rn(owA, c3, vA); vB = __ffma2_rn(owB, c3, vB);
          mA.x = fminf(mA.x, vA.x); mA.y = fminf(mA.y, vA.y); mB.x = fminf(mB.x, vB.x); mB.y = fminf(mB.y, vB.y);
        }
        hit |= (mA.x < sna[o]) | (mA.y < sna[o + 1]) | (mB.x < sna[o + 2]) | (mB.y < sna[o + 3]);
      }
    }
    else if (V == 7)  // scalar FFMA, 8 spheres per lane (a second lane of the group would take the other 8)
    {
#pragma unroll 2
      for (int o = 0; o < OBS; o++)
      {
        const float4 a = so[2 * o], b = so[2 * o + 1];
        const float na = sna[o];
        float m = 3e38f;
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
          float v = fmaf(a.x, cx[k], cb[k]);
          v = fmaf(a.z, cy[k], v); v = fmaf(b.x, cz[k], v); v = fmaf(b.z, cr[k], v);
          m = fminf(m, v);
        }
        hit |= (m < na);
      }
    }
    else  // V == 5: two obstacles per register pair, sphere constants broadcast (both halves equal)
    {
      for (int o = 0; o < OBS; o += 2)
      {
        const float4 a0 = so[2 * o], b0 = so[2 * o + 1], a1 = so[2 * o + 2], b1 = so[2 * o + 3];
        const float2 ox = make_float2(a0.x, a1.x), oy = make_float2(a0.z, a1.z), oz = make_float2(b0.x, b1.x), ow = make_float2(b0.z, b1.z);
        const float2 na = make_float2(sna[o], sna[o + 1]);
        float2 m = make_float2(3e38f, 3e38f);
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
          float2 v = __ffma2_rn(ox, make_float2(cx[k], cx[k]), make_float2(cb[k], cb[k]));
          v = __ffma2_rn(oy, make_float2(cy[k], cy[k]), v);
          v = __ffma2_rn(oz, make_float2(cz[k], cz[k]), v);
          v = __ffma2_rn(ow, make_float2(cr[k], cr[k]), v);
          m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y);
        }
        hit |= (m.x < na.x) | (m.y < na.y);
      }
    }
  }
  if (hit) out[0] = 1.0f;
}


template <int V, int OCC>
void run(float* out, float* in, float4* obs, int sms, cudaEvent_t e0, cudaEvent_t e1)
{
  for (int wps : { OCC })
  {
    const int blocks = sms * wps;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++)
    {
      cudaEventRecord(e0);
      kern<V, OCC><<<blocks, 128>>>(out, in, obs);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double fmas = (double)blocks * 128 * REPS * OBS * (V == 7 ? 32.0 : 64.0);
    printf("X%d warps/scheduler %d: %.3f ms  %.2f TFLOP/s (FMA)  %.3f FMA/clk/lane @1.9GHz  pairs/s %.3e\n", V, wps, best,
           2.0 * fmas / best / 1e9, fmas / 32.0 / (best * 1e-3) / (sms * 4.0) / 1.9e9, fmas / 4.0 / (best * 1e-3));
  }
}

int main()
{
  float *out, *in;
  float4* obs;
  cudaMalloc(&out, 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&obs, OBS * 16);
  cudaMemset(in, 0, 4096 * 4);
  cudaMemset(obs, 0, OBS * 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  run<0, 4>(out, in, obs, sms, e0, e1);
  run<7, 4>(out, in, obs, sms, e0, e1);
  run<7, 5>(out, in, obs, sms, e0, e1);
  run<7, 6>(out, in, obs, sms, e0, e1);
  run<7, 8>(out, in, obs, sms, e0, e1);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
