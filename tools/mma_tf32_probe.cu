// Feasibility probe for the next step of the sphere-arm pair test ("what comes next"): the expanded
// pair test  t = [ox oy oz ro Ko 1]·[-2cx -2cy -2cz -2r 1 (|c|²-r²)] < 0  is a K=6 inner product, i.e. a GEMM of the
// obstacle cloud (M) against the robot spheres of the states in flight (N).  This tool measures, on the GPU,
//   (1) the issue rate of the legacy tensor path for TF32 (mma.sync.m16n8k8) in MMA per clock per SM,
//   (2) a pre-filter loop shaped like the real kernel: A fragments (obstacles, static per world) from shared memory,
//       B fragments (sphere constants of 8 columns) in registers, FMNMX3 reduction of the 4 accumulators,
//       in single-pass TF32 and in the 3-term split (hi·hi + hi·lo + lo·hi), in sphere pairs per second,
//   (3) the error of both against fp64 on a synthetic cloud, i.e. the margin under which a pair has to be re-decided
//       with the spec'd fp32 sequence to keep the verdict bit-exact, and the fraction of 16x8 tiles that margin flags.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_probe tools/mma_tf32_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                                                          \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (x);                                                                                              \
    if (e_ != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                                  \
      exit(1);                                                                                                         \
    }                                                                                                                  \
  } while (0)

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2], const float (&c)[4])
{
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// (1) raw issue rate: ACC independent accumulator chains per warp
template <int ACC>
__global__ void rate_kernel(float* out, int iters)
{
  float acc[ACC][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; i++)
    a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  b[0] = __float_as_uint(0.5f);
  b[1] = __float_as_uint(0.25f);
  for (int j = 0; j < ACC; j++)
    for (int i = 0; i < 4; i++)
      acc[j][i] = 0.f;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int j = 0; j < ACC; j++)
      mma_tf32(acc[j], a, b, acc[j]);
  }
  float s = 0;
  for (int j = 0; j < ACC; j++)
    for (int i = 0; i < 4; i++)
      s += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (2) pre-filter loop.  Obstacle tiles in shared memory already in A-fragment order: tile = 16 obstacles x 8 k,
// 32 lanes x 4 words (one LDS.128 per lane per tile and term).  NB column groups of 8 spheres per warp.
// TERMS = 1: single TF32 pass; TERMS = 3: a_hi·b_hi + a_hi·b_lo + a_lo·b_hi (two A fragment sets, two B sets).
template <int NB, int TERMS>
__global__ void __launch_bounds__(128) filter_kernel(const uint4* __restrict__ tiles_hi, const uint4* __restrict__ tiles_lo,
                                                     int ntiles, const float* __restrict__ cols, float* out, int reps)
{
  extern __shared__ uint4 sm[];
  uint4* s_hi = sm;
  uint4* s_lo = sm + (size_t)ntiles * 32;
  for (int i = threadIdx.x; i < ntiles * 32; i += blockDim.x)
  {
    s_hi[i] = tiles_hi[i];
    if (TERMS == 3)
      s_lo[i] = tiles_lo[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
  float vmin = 3.4e38f;
  for (int r = 0; r < reps; r++)
  {
    // B fragments of this warp's NB column groups (a new "state" per rep: different columns)
    uint32_t bh[NB][2], bl[NB][2];
#pragma unroll
    for (int j = 0; j < NB; j++)
    {
      const float* c = cols + ((size_t)((warp * reps + r) % (8192 / NB)) * NB + j) * 64;  // 8 k x 8 n, k-major per column
      const int g = lane >> 2, t = lane & 3;
      float v0 = c[g * 8 + t], v1 = c[g * 8 + t + 4];
      uint32_t h0 = __float_as_uint(v0) & 0xffffe000u, h1 = __float_as_uint(v1) & 0xffffe000u;
      bh[j][0] = h0;
      bh[j][1] = h1;
      bl[j][0] = __float_as_uint(v0 - __uint_as_float(h0));
      bl[j][1] = __float_as_uint(v1 - __uint_as_float(h1));
    }
#pragma unroll 2
    for (int tl = 0; tl < ntiles; tl++)
    {
      const uint4 ah4 = s_hi[tl * 32 + lane];
      const uint32_t ah[4] = { ah4.x, ah4.y, ah4.z, ah4.w };
      uint32_t al[4];
      if (TERMS == 3)
      {
        const uint4 al4 = s_lo[tl * 32 + lane];
        al[0] = al4.x;
        al[1] = al4.y;
        al[2] = al4.z;
        al[3] = al4.w;
      }
#pragma unroll
      for (int j = 0; j < NB; j++)
      {
        float d[4];
        mma_tf32(d, ah, bh[j], zero);
        if (TERMS == 3)
        {
          mma_tf32(d, ah, bl[j], d);
          mma_tf32(d, al, bh[j], d);
        }
        vmin = fminf(vmin, fminf(d[0], d[1]));
        vmin = fminf(vmin, fminf(d[2], d[3]));
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = vmin;
}

// (3) values out, for the error study: one warp per column group, all tiles
template <int TERMS>
__global__ void values_kernel(const uint4* __restrict__ tiles_hi, const uint4* __restrict__ tiles_lo, int ntiles,
                              const float* __restrict__ cols, int ngroups, float* out)
{
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= ngroups)
    return;
  const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
  const int g = lane >> 2, t = lane & 3;
  const float* c = cols + (size_t)warp * 64;
  float v0 = c[g * 8 + t], v1 = c[g * 8 + t + 4];
  uint32_t bh[2], bl[2];
  bh[0] = __float_as_uint(v0) & 0xffffe000u;
  bh[1] = __float_as_uint(v1) & 0xffffe000u;
  bl[0] = __float_as_uint(v0 - __uint_as_float(bh[0]));
  bl[1] = __float_as_uint(v1 - __uint_as_float(bh[1]));
  for (int tl = 0; tl < ntiles; tl++)
  {
    const uint4 ah4 = tiles_hi[tl * 32 + lane], al4 = tiles_lo[tl * 32 + lane];
    const uint32_t ah[4] = { ah4.x, ah4.y, ah4.z, ah4.w }, al[4] = { al4.x, al4.y, al4.z, al4.w };
    float d[4];
    mma_tf32(d, ah, bh, zero);
    if (TERMS == 3)
    {
      mma_tf32(d, ah, bl, d);
      mma_tf32(d, al, bh, d);
    }
    // D layout: d0 (row g, col 2t), d1 (g, 2t+1), d2 (g+8, 2t), d3 (g+8, 2t+1)
    float* o = out + ((size_t)warp * ntiles + tl) * 128;
    o[g * 8 + 2 * t] = d[0];
    o[g * 8 + 2 * t + 1] = d[1];
    o[(g + 8) * 8 + 2 * t] = d[2];
    o[(g + 8) * 8 + 2 * t + 1] = d[3];
  }
}

static double urand()
{
  return (double)rand() / RAND_MAX;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float* d_out;
  CK(cudaMalloc(&d_out, sizeof(float) * sms * 64 * 1024));

  // (1)
  for (int warps = 4; warps <= 16; warps *= 2)
  {
    const int iters = 20000;
    auto run = [&](auto kern, int acc) {
      kern<<<sms, warps * 32>>>(d_out, 100);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      kern<<<sms, warps * 32>>>(d_out, iters);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double mmas = (double)sms * warps * iters * acc;
      printf("rate: %2d warps/SM, %d chains: %.3f ms, %.3f MMA/clk/SM at %.0f MHz, %.1f TFLOP/s (m16n8k8 = 2048 flop)\n", warps,
             acc, ms, mmas / sms / (ms * 1e-3 * clk_khz * 1e3), clk_khz / 1e3, mmas * 2048 / (ms * 1e-3) / 1e12);
    };
    run(rate_kernel<4>, 4);
    run(rate_kernel<8>, 8);
  }

  // synthetic world: obstacles in [-1.5,1.5]^3 with radius 0.03..0.10; robot spheres in [-1,1]^3 radius 0.04..0.12
  const int nobs = 1024, ntiles = nobs / 16, ngroups = 4096 * 2;  // 8 columns per group
  std::vector<double> ob(nobs * 4), sp((size_t)ngroups * 8 * 4);
  srand(7);
  for (int i = 0; i < nobs; i++)
  {
    for (int d = 0; d < 3; d++)
      ob[i * 4 + d] = -1.5 + 3.0 * urand();
    ob[i * 4 + 3] = 0.03 + 0.07 * urand();
  }
  for (size_t i = 0; i < (size_t)ngroups * 8; i++)
  {
    for (int d = 0; d < 3; d++)
      sp[i * 4 + d] = -1.0 + 2.0 * urand();
    sp[i * 4 + 3] = 0.04 + 0.08 * urand();
  }
  // A rows (fp32): [ox oy oz ro Ko 1 0 0]; B columns: [-2cx -2cy -2cz -2r 1 e 0 0], e = |c|²-r²
  std::vector<float> A(nobs * 8, 0.f), B((size_t)ngroups * 64, 0.f);
  for (int i = 0; i < nobs; i++)
  {
    const float x = (float)ob[i * 4], y = (float)ob[i * 4 + 1], z = (float)ob[i * 4 + 2], r = (float)ob[i * 4 + 3];
    A[i * 8 + 0] = x;
    A[i * 8 + 1] = y;
    A[i * 8 + 2] = z;
    A[i * 8 + 3] = r;
    A[i * 8 + 4] = x * x + y * y + z * z - r * r;
    A[i * 8 + 5] = 1.f;
  }
  for (size_t i = 0; i < (size_t)ngroups * 8; i++)
  {
    const float x = (float)sp[i * 4], y = (float)sp[i * 4 + 1], z = (float)sp[i * 4 + 2], r = (float)sp[i * 4 + 3];
    float* c = &B[(i / 8) * 64 + (i % 8) * 8];  // column n = i%8: k contiguous
    c[0] = -2.f * x;
    c[1] = -2.f * y;
    c[2] = -2.f * z;
    c[3] = -2.f * r;
    c[4] = 1.f;
    c[5] = x * x + y * y + z * z - r * r;
  }
  // A fragments: a0 (row g, k t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4)
  std::vector<uint32_t> th((size_t)ntiles * 32 * 4), tlo((size_t)ntiles * 32 * 4);
  auto split = [](float v, uint32_t& hi, uint32_t& lo) {
    uint32_t u;
    memcpy(&u, &v, 4);
    hi = u & 0xffffe000u;
    float h;
    memcpy(&h, &hi, 4);
    float l = v - h;
    memcpy(&lo, &l, 4);
  };
  for (int tl = 0; tl < ntiles; tl++)
    for (int lane = 0; lane < 32; lane++)
    {
      const int g = lane >> 2, t = lane & 3;
      const int rows[4] = { g, g + 8, g, g + 8 }, ks[4] = { t, t, t + 4, t + 4 };
      for (int i = 0; i < 4; i++)
        split(A[(tl * 16 + rows[i]) * 8 + ks[i]], th[((size_t)tl * 32 + lane) * 4 + i], tlo[((size_t)tl * 32 + lane) * 4 + i]);
    }
  uint4 *d_th, *d_tl;
  float* d_cols;
  CK(cudaMalloc(&d_th, th.size() * 4));
  CK(cudaMalloc(&d_tl, tlo.size() * 4));
  CK(cudaMalloc(&d_cols, B.size() * 4));
  CK(cudaMemcpy(d_th, th.data(), th.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tl, tlo.data(), tlo.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cols, B.data(), B.size() * 4, cudaMemcpyHostToDevice));

  // (2) a warp covers NB*8 spheres (NB=2: one 16-sphere arm state) against the whole cloud per rep
  auto filt = [&](auto kern, int nb, int terms, const char* name) {
    const size_t smem = (size_t)ntiles * 32 * 16 * 2;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int reps = 64, grid = sms * 4;
    kern<<<grid, 128, smem>>>(d_th, d_tl, ntiles, d_cols, d_out, 2);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<grid, 128, smem>>>(d_th, d_tl, ntiles, d_cols, d_out, reps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double pairs = (double)grid * 4 * reps * nb * 8 * nobs;
    printf("filter %-22s NB=%d terms=%d: %.3f ms, %.3e pairs/s (arm_state_kernel today: 3.8e12 executed pairs/s)\n", name, nb, terms,
           ms, pairs / (ms * 1e-3));
  };
  filt(filter_kernel<2, 1>, 2, 1, "tf32x1");
  filt(filter_kernel<4, 1>, 4, 1, "tf32x1");
  filt(filter_kernel<2, 3>, 2, 3, "tf32x3");
  filt(filter_kernel<4, 3>, 4, 3, "tf32x3");

  // (3)
  const int eg = 512;  // column groups in the error study
  float* d_vals;
  CK(cudaMalloc(&d_vals, sizeof(float) * (size_t)eg * ntiles * 128));
  std::vector<float> v1((size_t)eg * ntiles * 128), v3(v1.size());
  values_kernel<1><<<(eg * 32 + 127) / 128, 128>>>(d_th, d_tl, ntiles, d_cols, eg, d_vals);
  CK(cudaMemcpy(v1.data(), d_vals, v1.size() * 4, cudaMemcpyDeviceToHost));
  values_kernel<3><<<(eg * 32 + 127) / 128, 128>>>(d_th, d_tl, ntiles, d_cols, eg, d_vals);
  CK(cudaMemcpy(v3.data(), d_vals, v3.size() * 4, cudaMemcpyDeviceToHost));
  double e1max = 0, e3max = 0, e32max = 0;
  std::vector<double> exact(v1.size());
  std::vector<float> f32v(v1.size());
  for (int w = 0; w < eg; w++)
    for (int tl = 0; tl < ntiles; tl++)
      for (int row = 0; row < 16; row++)
        for (int n = 0; n < 8; n++)
        {
          const float* a = &A[(tl * 16 + row) * 8];
          const float* b = &B[(size_t)w * 64 + n * 8];
          double s = 0;
          for (int k = 0; k < 6; k++)
            s += (double)a[k] * (double)b[k];
          // the fp32 sequence of the spec'd pair test (fma chain), for the scale of ITS deviation from the exact value
          float t = fmaf(a[0], b[0], a[4]);
          t = fmaf(a[1], b[1], t);
          t = fmaf(a[2], b[2], t);
          t = fmaf(a[3], b[3], t);
          t += b[5];
          const size_t idx = ((size_t)w * ntiles + tl) * 128 + row * 8 + n;
          exact[idx] = s;
          f32v[idx] = t;
          e1max = fmax(e1max, fabs(v1[idx] - s));
          e3max = fmax(e3max, fabs(v3[idx] - s));
          e32max = fmax(e32max, fabs(t - s));
        }
  printf("error vs fp64: tf32x1 max %.3e, tf32x3 max %.3e, fp32 fma chain max %.3e\n", e1max, e3max, e32max);
  // margin = max error of the filter + max error of the fp32 sequence; tiles flagged = any value < margin
  for (int terms = 1; terms <= 3; terms += 2)
  {
    const std::vector<float>& v = terms == 1 ? v1 : v3;
    const double margin = 1.5 * ((terms == 1 ? e1max : e3max) + e32max);
    size_t flagged = 0, colliding = 0, missed = 0, ntl = (size_t)eg * ntiles;
    for (size_t tI = 0; tI < ntl; tI++)
    {
      bool fl = false, col = false;
      for (int i = 0; i < 128; i++)
      {
        fl |= v[tI * 128 + i] < margin;
        col |= f32v[tI * 128 + i] < 0.f;
      }
      flagged += fl;
      colliding += col;
      missed += (col && !fl);
    }
    printf("terms=%d margin %.3e: tiles flagged %.4f, tiles truly colliding %.4f, missed %zu\n", terms, margin,
           (double)flagged / ntl, (double)colliding / ntl, missed);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
