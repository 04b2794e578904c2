// K1: batched forward complex FFT (fp32) for OFDM symbols -- 32K / 16K (and any n = 256 * 2^m >= 1024),
// unnormalised, with the half-swap (fftshift) of fast_fourier_transform::execute folded into the store.
//
// Reference semantics (src/DSP/fast_fourier_transform.h:54-70): out = fftshift(DFT_forward(in)), FFTW
// sign convention (e^{-j 2 pi n k / N}), no 1/N.  The reference calls FFTW 3.3.8 (binary only), so
// there is no bit-exact target: the contract is <= 1e-5 * max|X| against a float64 DFT.
//
// B200 design: four-step decomposition N = N1 x 256 in two streaming passes.  A 32K symbol (256 KiB) does
// not fit one SM's shared memory, so pass A transforms 16 columns at a time (length N1, stride 256) and
// applies the inter-pass twiddle, pass B transforms 16 rows at a time (length 256) and writes the shifted
// spectrum; every global access is a full 128-byte line and the intermediate stays L2-resident because
// the batch is walked in chunks smaller than L2.  Inside a tile the 16 transforms run side by side as
// radix-4 Stockham stages (auto-sorting: no bit reversal) ping-ponging between two shared-memory images
// laid out [element][17] so that stage accesses and tile loads / stores are bank-conflict free.
#include "ctx.h"
#include <cmath>
#include <vector>

struct FftPlan {
  int n = 0, n1 = 0;
  float2* d_wn = nullptr;      // W_n^m, m < n          (inter-pass twiddles)
  float2* d_w256 = nullptr;    // W_256^m, m < 256      (stage twiddles)
};

namespace {

constexpr int TILE = 16;       // transforms per CTA
constexpr int PITCH = 17;      // shared-memory pitch per element (float2 units)
constexpr int N2 = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One radix-R Stockham stage over TILE transforms of length L: x -> y (both [L][PITCH]).
// wl = W_L^m table (L entries, built from W_256 by stride).  Ns = product of the radices already done.
template <int R>
__device__ __forceinline__ void stockham_stage(const float2* __restrict__ x, float2* __restrict__ y,
                                               const float2* __restrict__ w256, int L, int Ns)
{
  const int T = L / R;
  const int wstep = (N2 / L) * (L / (Ns * R));       // exp(-2 pi i r k / (Ns R)) = W_256^(r k wstep)
  for (int t = threadIdx.x; t < T * TILE; t += blockDim.x) {
    const int c = t % TILE, j = t / TILE;
    const int k = j & (Ns - 1);
    float2 u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      u[r] = x[(j + r * T) * PITCH + c];
      if (r) u[r] = cmul(u[r], w256[(r * k * wstep) & (N2 - 1)]);
    }
    const int j0 = (j - k) * R + k;
    if (R == 2) {
      y[j0 * PITCH + c] = make_float2(u[0].x + u[1].x, u[0].y + u[1].y);
      y[(j0 + Ns) * PITCH + c] = make_float2(u[0].x - u[1].x, u[0].y - u[1].y);
    } else {
      const float2 a0 = make_float2(u[0].x + u[2].x, u[0].y + u[2].y), a1 = make_float2(u[0].x - u[2].x, u[0].y - u[2].y);
      const float2 a2 = make_float2(u[1].x + u[3].x, u[1].y + u[3].y);
      const float2 d = make_float2(u[1].x - u[3].x, u[1].y - u[3].y);
      const float2 a3 = make_float2(d.y, -d.x);                                   // (u1 - u3) * (-j)
      y[j0 * PITCH + c] = make_float2(a0.x + a2.x, a0.y + a2.y);
      y[(j0 + Ns) * PITCH + c] = make_float2(a1.x + a3.x, a1.y + a3.y);
      y[(j0 + 2 * Ns) * PITCH + c] = make_float2(a0.x - a2.x, a0.y - a2.y);
      y[(j0 + 3 * Ns) * PITCH + c] = make_float2(a1.x - a3.x, a1.y - a3.y);
    }
  }
}

// all stages of the length-L transforms held in buf0; returns the buffer holding the result
__device__ __forceinline__ float2* stockham_all(float2* buf0, float2* buf1, const float2* w256, int L)
{
  float2 *x = buf0, *y = buf1;
  int Ns = 1;
  while (Ns < L) {
    if (L / Ns >= 4) { stockham_stage<4>(x, y, w256, L, Ns); Ns *= 4; }
    else { stockham_stage<2>(x, y, w256, L, Ns); Ns *= 2; }
    __syncthreads();
    float2* t = x; x = y; y = t;
  }
  return x;
}

// Pass A: for 16 consecutive columns n2: Y[k1][n2] = W_n^(k1 n2) * sum_n1 x[n1*256 + n2] W_n1^(n1 k1)
__global__ void __launch_bounds__(256) fft_pass_a(const float2* __restrict__ in, float2* __restrict__ tmp,
                                                   const float2* __restrict__ wn, const float2* __restrict__ w256g,
                                                   int n, int n1)
{
  extern __shared__ __align__(16) float2 sm[];
  float2* w256 = sm;
  float2* buf0 = sm + N2;
  float2* buf1 = buf0 + n1 * PITCH;
  const int n2_0 = blockIdx.x * TILE;
  const float2* src = in + (size_t)blockIdx.y * n;
  float2* dst = tmp + (size_t)blockIdx.y * n;
  for (int i = threadIdx.x; i < N2; i += blockDim.x) w256[i] = __ldg(w256g + i);
  for (int t = threadIdx.x; t < n1 * TILE; t += blockDim.x) {
    const int c = t % TILE, e = t / TILE;
    buf0[e * PITCH + c] = __ldg(src + (size_t)e * N2 + n2_0 + c);
  }
  __syncthreads();
  const float2* res = stockham_all(buf0, buf1, w256, n1);
  for (int t = threadIdx.x; t < n1 * TILE; t += blockDim.x) {
    const int c = t % TILE, k1 = t / TILE;
    const float2 w = __ldg(wn + k1 * (n2_0 + c));
    dst[(size_t)k1 * N2 + n2_0 + c] = cmul(res[k1 * PITCH + c], w);
  }
}

// Pass B: for 16 consecutive rows k1: X[k1 + n1*k2] = sum_n2 Y[k1][n2] W_256^(n2 k2), stored half-swapped
__global__ void __launch_bounds__(256) fft_pass_b(const float2* __restrict__ tmp, float2* __restrict__ out,
                                                   const float2* __restrict__ w256g, int n, int n1)
{
  extern __shared__ __align__(16) float2 sm[];
  float2* w256 = sm;
  float2* buf0 = sm + N2;
  float2* buf1 = buf0 + N2 * PITCH;
  const int k1_0 = blockIdx.x * TILE;
  const float2* src = tmp + (size_t)blockIdx.y * n;
  float2* dst = out + (size_t)blockIdx.y * n;
  for (int i = threadIdx.x; i < N2; i += blockDim.x) w256[i] = __ldg(w256g + i);
  for (int t = threadIdx.x; t < N2 * TILE; t += blockDim.x) {
    const int e = t % N2, c = t / N2;                       // coalesced along the row
    buf0[e * PITCH + c] = __ldg(src + (size_t)(k1_0 + c) * N2 + e);
  }
  __syncthreads();
  const float2* res = stockham_all(buf0, buf1, w256, N2);
  const int half = n >> 1;
  for (int t = threadIdx.x; t < N2 * TILE; t += blockDim.x) {
    const int c = t % TILE, k2 = t / TILE;
    const int k = k1_0 + c + n1 * k2;
    dst[(k + half) & (n - 1)] = res[k2 * PITCH + c];          // fast_fourier_transform.h:67-68
  }
}

}  // namespace

void t2_fft_free(t2b200_ctx* ctx)
{
  for (auto& kv : ctx->fft) { cudaFree(kv.second->d_wn); cudaFree(kv.second->d_w256); delete kv.second; }
  ctx->fft.clear();
}

static int get_plan(t2b200_ctx* ctx, int n, FftPlan** out)
{
  auto it = ctx->fft.find(n);
  if (it != ctx->fft.end()) { *out = it->second; return T2B200_OK; }
  if (n < 1024 || n > 32768 || (n & (n - 1))) { ctx->err = "t2b200_fft: n must be a power of two in [1024, 32768]"; return T2B200_ERR_ARG; }
  FftPlan* p = new FftPlan();
  p->n = n; p->n1 = n / N2;
  std::vector<float2> wn(n), w256(N2);
  const double tau = 6.283185307179586476925286766559;
  for (int m = 0; m < n; ++m) wn[m] = make_float2((float)cos(tau * m / n), (float)-sin(tau * m / n));
  for (int m = 0; m < N2; ++m) w256[m] = make_float2((float)cos(tau * m / N2), (float)-sin(tau * m / N2));
  T2_CUDA(ctx, cudaMalloc(&p->d_wn, n * sizeof(float2)));
  T2_CUDA(ctx, cudaMalloc(&p->d_w256, N2 * sizeof(float2)));
  T2_CUDA(ctx, cudaMemcpy(p->d_wn, wn.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
  T2_CUDA(ctx, cudaMemcpy(p->d_w256, w256.data(), N2 * sizeof(float2), cudaMemcpyHostToDevice));
  ctx->fft[n] = p;
  *out = p;
  return T2B200_OK;
}

// device-resident batch: in, out, tmp all on the device
int t2_fft_device(t2b200_ctx* ctx, int n, const float2* d_in, int batch, float2* d_out, float2* d_tmp)
{
  FftPlan* p; int rc;
  if ((rc = get_plan(ctx, n, &p))) return rc;
  const size_t smem_a = (size_t)(N2 + 2 * p->n1 * PITCH) * sizeof(float2);
  const size_t smem_b = (size_t)(N2 + 2 * N2 * PITCH) * sizeof(float2);
  T2_CUDA(ctx, cudaFuncSetAttribute(fft_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  T2_CUDA(ctx, cudaFuncSetAttribute(fft_pass_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  // walk the batch in chunks whose input + intermediate + output stay well inside the 126 MB L2
  const int chunk = std::max(1, (int)((48u << 20) / ((size_t)n * sizeof(float2) * 2)));
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = std::min(chunk, batch - b0);
    fft_pass_a<<<dim3(N2 / TILE, nb), 256, smem_a, ctx->stream>>>(d_in + (size_t)b0 * n, d_tmp + (size_t)(b0 % chunk) * n, p->d_wn, p->d_w256, n, p->n1);
    T2_CUDA(ctx, cudaGetLastError());
    fft_pass_b<<<dim3(p->n1 / TILE, nb), 256, smem_b, ctx->stream>>>(d_tmp + (size_t)(b0 % chunk) * n, d_out + (size_t)b0 * n, p->d_w256, n, p->n1);
    T2_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
  }
  return T2B200_OK;
}

extern "C" int t2b200_fft(t2b200_ctx* ctx, int n, const float* in, int batch, float* out)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (!in || !out || batch < 0) { ctx->err = "t2b200_fft: bad argument"; return T2B200_ERR_ARG; }
  if (n < 4096 || n > 32768 || (n & (n - 1))) { ctx->err = "t2b200_fft: n must be 4096, 8192, 16384 or 32768"; return T2B200_ERR_ARG; }
  if (batch == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc; const void* din; void *dout, *dtmp;
  const size_t bytes = (size_t)batch * n * sizeof(float2);
  if ((rc = t2_to_device(ctx, 0, in, bytes, &din))) return rc;
  if ((rc = t2_out_device(ctx, 1, out, bytes, &dout))) return rc;
  const int chunk = std::max(1, (int)((48u << 20) / ((size_t)n * sizeof(float2) * 2)));
  if ((rc = t2_dev_scratch(ctx, 4, (size_t)std::min(chunk, batch) * n * sizeof(float2), &dtmp))) return rc;
  if ((rc = t2_fft_device(ctx, n, (const float2*)din, batch, (float2*)dout, (float2*)dtmp))) return rc;
  return t2_finish_out(ctx, out, dout, bytes);
}
