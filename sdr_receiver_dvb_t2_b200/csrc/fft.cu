// K1: batched forward complex FFT (fp32) for OFDM symbols -- 32K / 16K (and 8K / 4K: n = 256 * 16 * A, A = 1, 2, 4, 8),
// unnormalised, with the half-swap (fftshift) of fast_fourier_transform::execute folded into the store.
//
// Reference semantics (src/DSP/fast_fourier_transform.h:54-70): out = fftshift(DFT_forward(in)), FFTW
// sign convention (e^{-j 2 pi n k / N}), no 1/N.  The reference calls FFTW 3.3.8 (binary only), so
// there is no bit-exact target: the contract is <= 1e-5 * max|X| against a float64 DFT.
//
// B200 design: four-step decomposition N = N1 x 256 in two streaming passes.  A 32K symbol (256 KiB) does
// not fit one SM's shared memory, so pass A transforms 32 columns at a time (length N1 = 16 A, stride 256) and
// applies the inter-pass twiddle, pass B transforms 16 rows at a time (length 256) and writes the shifted
// spectrum; every global access is a full 128- or 256-byte run and the intermediate stays L2-resident because
// the batch is walked in chunks smaller than L2.  The butterflies run in REGISTERS: every thread loads 16
// samples straight from global memory, does a radix-16 DFT (two radix-4 levels, constants folded), and one
// shared-memory exchange hands the data to the second register stage (radix-A in pass A, radix-16 in pass B),
// whose results are stored straight to global memory.
#include "stages.h"
#include <cmath>
#include <cstdlib>
#include <vector>

struct FftPlan {
  int n = 0, n1 = 0;
  float2* d_wn = nullptr;      // W_n^m, m < n          (inter-pass twiddles)
  float2* d_w256 = nullptr;    // W_256^m, m < 256      (stage twiddles)
};

namespace {

constexpr int N2 = 256;
constexpr int kColsA = 32;     // columns per pass-A CTA (one warp = 32 consecutive columns = 256 contiguous bytes)
constexpr int kRowsB = 16;     // rows per pass-B CTA
constexpr int kPitchB = 17;    // shared-memory pitch (float2) of the pass-B exchange

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// 4-point DFT in place, natural order
__device__ __forceinline__ void fft4(float2& x0, float2& x1, float2& x2, float2& x3)
{
  const float2 a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3), d = csub(x1, x3);
  const float2 a3 = make_float2(d.y, -d.x);                      // (x1 - x3) * (-j)
  x0 = cadd(a0, a2); x1 = cadd(a1, a3); x2 = csub(a0, a2); x3 = csub(a1, a3);
}
// multiply by W_16^e, e a compile-time constant
template <int E> __device__ __forceinline__ float2 mul_w16(float2 v)
{
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
  constexpr int e = E & 15;
  if (e == 0) return v;
  if (e == 4) return make_float2(v.y, -v.x);
  if (e == 8) return make_float2(-v.x, -v.y);
  if (e == 12) return make_float2(-v.y, v.x);
  if (e == 2) return make_float2(H * (v.x + v.y), H * (v.y - v.x));
  if (e == 6) return make_float2(H * (v.y - v.x), -H * (v.x + v.y));
  constexpr float wr = e == 1 ? C1 : e == 3 ? S1 : e == 9 ? -C1 : e == 5 ? -S1 : e == 7 ? -C1 : 0.0f;
  constexpr float wi = e == 1 ? -S1 : e == 3 ? -C1 : e == 9 ? S1 : e == 5 ? -C1 : e == 7 ? -S1 : 0.0f;
  return make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
}
// 16-point DFT of x[n] in place (n = 4*n1 + n2).  Output X[K] is left in x[4*(K & 3) + (K >> 2)].
__device__ __forceinline__ void fft16(float2 (&x)[16])
{
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);     // k1 now at x[4*k1 + n2]
  x[5] = mul_w16<1>(x[5]);  x[6] = mul_w16<2>(x[6]);   x[7] = mul_w16<3>(x[7]);       // W_16^(n2*k1)
  x[9] = mul_w16<2>(x[9]);  x[10] = mul_w16<4>(x[10]); x[11] = mul_w16<6>(x[11]);
  x[13] = mul_w16<3>(x[13]); x[14] = mul_w16<6>(x[14]); x[15] = mul_w16<9>(x[15]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
}
__device__ __forceinline__ constexpr int pos16(int K) { return 4 * (K & 3) + (K >> 2); }

// A-point DFT (A = 1, 2, 4, 8) of x[0..A) in place, natural order
template <int A> __device__ __forceinline__ void fft_small(float2 (&x)[8])
{
  if (A == 2) { const float2 t = x[0]; x[0] = cadd(t, x[1]); x[1] = csub(t, x[1]); }
  if (A == 4) fft4(x[0], x[1], x[2], x[3]);
  if (A == 8) {
    // n = 2*n1 + n2: two 4-point DFTs over n1, twiddle W_8^(n2*k1), then the 2-point stage
    fft4(x[0], x[2], x[4], x[6]);
    fft4(x[1], x[3], x[5], x[7]);
    const float2 y1 = mul_w16<2>(x[3]), y2 = mul_w16<4>(x[5]), y3 = mul_w16<6>(x[7]);
    const float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1];
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, y1); x[5] = csub(e1, y1);
    x[2] = cadd(e2, y2); x[6] = csub(e2, y2);
    x[3] = cadd(e3, y3); x[7] = csub(e3, y3);
  }
}

// Pass A: columns n2 of the N1 x 256 view: Y[k1][n2] = W_n^(k1 n2) * sum_n1 x[n1*256 + n2] W_N1^(n1 k1), N1 = 16*A.
// Thread (a = warp, c = lane) loads the 16 samples n1 = A*r + a of column n2_0 + c straight from global memory (a warp
// reads 256 contiguous bytes per instruction), runs the radix-16 butterflies in registers, passes the results through one
// shared-memory exchange to the radix-A stage (again in registers) and stores Y, inter-pass twiddle applied, as 256-byte rows.
// I16: the samples are the int16 I/Q pairs of a device front-end, converted on load with the reference's scale
// (dvbt2_demodulator.cpp:182-186: real = I * short_to_float) -- 4 bytes per sample over HBM and PCIe instead of 8.
template <int A, bool I16>
__global__ void __launch_bounds__(32 * A) fft_pass_a(const void* __restrict__ in_any, float2* __restrict__ tmp,
                                                      const float2* __restrict__ wn, int n, float scale)
{
  constexpr int N1 = 16 * A;
  __shared__ float2 sm[16 * A * kColsA];
  const int c = threadIdx.x & 31, a = threadIdx.x >> 5;
  const int n2 = blockIdx.x * kColsA + c;
  float2* dst = tmp + (size_t)blockIdx.y * n + n2;
  float2 x[16];
  if (I16) {
    const short2* src = static_cast<const short2*>(in_any) + (size_t)blockIdx.y * n + n2;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const short2 v = __ldg(src + (size_t)(A * r + a) * N2);
      x[r] = make_float2((float)v.x * scale, (float)v.y * scale);
    }
  } else {
    const float2* src = static_cast<const float2*>(in_any) + (size_t)blockIdx.y * n + n2;
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = __ldg(src + (size_t)(A * r + a) * N2);
  }
  fft16(x);
  const int wstep1 = n / N1;                                       // W_N1^e = W_n^(e * n / N1)
#pragma unroll
  for (int kr = 0; kr < 16; ++kr) {
    float2 v = x[pos16(kr)];
    if (A > 1 && kr) v = cmul(v, __ldg(wn + ((a * kr * wstep1) & (n - 1))));    // warp-uniform address
    sm[(kr * A + a) * kColsA + c] = v;
  }
  __syncthreads();
  const float2 wstep = __ldg(wn + ((16 * n2) & (n - 1)));          // W_n^(16 n2): k1 -> k1 + 16
#pragma unroll
  for (int j = 0; j < 16 / A; ++j) {
    const int kr = a + A * j;
    float2 y[8];
#pragma unroll
    for (int aa = 0; aa < A; ++aa) y[aa] = sm[(kr * A + aa) * kColsA + c];
    fft_small<A>(y);
    float2 w = __ldg(wn + ((kr * n2) & (n - 1)));
#pragma unroll
    for (int ka = 0; ka < A; ++ka) {
      dst[(size_t)(kr + 16 * ka) * N2] = cmul(y[ka], w);
      w = cmul(w, wstep);
    }
  }
}

// Pass B: rows k1 of Y: X[k1 + N1*k2] = sum_n2 Y[k1][n2] W_256^(n2 k2), stored half-swapped.  16 rows per CTA.  Stage 1:
// thread (a = lane & 15, row) takes n2 = 16*r + a (a half-warp reads 128 contiguous bytes); stage 2: thread (kr, row) with
// the ROW index fastest across lanes, so that the 16 results X[k1_0 .. k1_0+15 + N1*k2] of a half-warp are contiguous.
__global__ void __launch_bounds__(256) fft_pass_b(const float2* __restrict__ tmp, float2* __restrict__ out,
                                                   const float2* __restrict__ w256g, int n, int n1)
{
  __shared__ float2 sm[256 * kPitchB];
  const int k1_0 = blockIdx.x * kRowsB;
  const float2* src = tmp + (size_t)blockIdx.y * n;
  float2* dst = out + (size_t)blockIdx.y * n;
  float2 x[16];
  {
    const int a = threadIdx.x & 15, c = threadIdx.x >> 4;
    const float2* row = src + (size_t)(k1_0 + c) * N2 + a;
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = __ldg(row + 16 * r);
    fft16(x);
#pragma unroll
    for (int kr = 0; kr < 16; ++kr) {
      float2 v = x[pos16(kr)];
      if (kr) v = cmul(v, __ldg(w256g + ((a * kr) & (N2 - 1))));
      sm[(kr * 16 + a) * kPitchB + c] = v;
    }
  }
  __syncthreads();
  {
    const int c = threadIdx.x & 15, kr = threadIdx.x >> 4;
#pragma unroll
    for (int a = 0; a < 16; ++a) x[a] = sm[(kr * 16 + a) * kPitchB + c];
    fft16(x);
    const int half = n >> 1;
#pragma unroll
    for (int ka = 0; ka < 16; ++ka) {
      const int k = k1_0 + c + n1 * (kr + 16 * ka);
      dst[(k + half) & (n - 1)] = x[pos16(ka)];                    // fast_fourier_transform.h:67-68
    }
  }
}

// Small transforms (n = 256 .. 2048: the 1K FFT of the P1 symbol, p1_symbol.cpp:34-35,114): one CTA per transform, the
// whole symbol in shared memory, radix-2 Stockham stages with twiddles W_n^m from the plan's table, half-swap on store.
// One such transform per T2 frame is all the receiver does, so nothing here is tuned.
__global__ void __launch_bounds__(256) fft_small_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                        const float2* __restrict__ wn, int n, int log2n)
{
  extern __shared__ float2 sm_small[];
  float2* a = sm_small;
  float2* b = sm_small + n;
  const float2* src = in + (size_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = src[i];
  __syncthreads();
  // Stockham auto-sort, decimation in frequency: after stage s the sub-transform length is n >> (s + 1)
  int l = n >> 1, m = 1;
  for (int s = 0; s < log2n; ++s) {
    for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
      const int j = t / m, k = t - j * m;                          // j < l, k < m
      const float2 c0 = a[k + j * m], c1 = a[k + j * m + l * m];
      const float2 w = wn[(j * m) & (n - 1)];                      // W_n^(j * m) = W_(2l)^j
      b[k + 2 * j * m] = cadd(c0, c1);
      b[k + 2 * j * m + m] = cmul(csub(c0, c1), w);
    }
    __syncthreads();
    float2* t2 = a; a = b; b = t2;
    l >>= 1; m <<= 1;
  }
  float2* dst = out + (size_t)blockIdx.x * n;
  const int half = n >> 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[(i + half) & (n - 1)] = a[i];
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
  if (n < 256 || n > 32768 || (n & (n - 1))) { ctx->err = "t2b200_fft: n must be a power of two in [256, 32768]"; return T2B200_ERR_ARG; }
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

// device-resident batch: in, out, tmp all on the device.  d_in16 != nullptr: int16 I/Q input (scale applied on load)
int t2_fft_device(t2b200_ctx* ctx, int n, const float2* d_in, int batch, float2* d_out, float2* d_tmp, const short2* d_in16,
                  float scale)
{
  FftPlan* p; int rc;
  if ((rc = get_plan(ctx, n, &p))) return rc;
  if (n < 4096) {
    if (d_in16) { ctx->err = "t2b200_fft: int16 input needs n >= 4096"; return T2B200_ERR_ARG; }
    int log2n = 0;
    while ((1 << log2n) < n) ++log2n;
    fft_small_kernel<<<batch, 256, 2 * (size_t)n * sizeof(float2), ctx->stream>>>(d_in, d_out, p->d_wn, n, log2n);
    T2_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return T2B200_OK;
  }
  // walk the batch in chunks whose input + intermediate + output stay well inside the 126 MB L2
  const int chunk = std::max(1, (int)((48u << 20) / ((size_t)n * sizeof(float2) * 2)));
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = std::min(chunk, batch - b0);
    const void* cin = d_in16 ? (const void*)(d_in16 + (size_t)b0 * n) : (const void*)(d_in + (size_t)b0 * n);
    float2* ctmp = d_tmp;
    const dim3 ga(N2 / kColsA, nb);
#define T2_PASS_A(AA) \
    if (d_in16) fft_pass_a<AA, true><<<ga, 32 * AA, 0, ctx->stream>>>(cin, ctmp, p->d_wn, n, scale); \
    else fft_pass_a<AA, false><<<ga, 32 * AA, 0, ctx->stream>>>(cin, ctmp, p->d_wn, n, 1.0f)
    switch (p->n1 / 16) {
      case 1: T2_PASS_A(1); break;
      case 2: T2_PASS_A(2); break;
      case 4: T2_PASS_A(4); break;
      default: T2_PASS_A(8); break;
    }
#undef T2_PASS_A
    T2_CUDA(ctx, cudaGetLastError());
    fft_pass_b<<<dim3(p->n1 / kRowsB, nb), 256, 0, ctx->stream>>>(ctmp, d_out + (size_t)b0 * n, p->d_w256, n, p->n1);
    T2_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
  }
  return T2B200_OK;
}

extern "C" int t2b200_fft(t2b200_ctx* ctx, int n, const float* in, int batch, float* out)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (!in || !out || batch < 0) { ctx->err = "t2b200_fft: bad argument"; return T2B200_ERR_ARG; }
  if (n < 256 || n > 32768 || (n & (n - 1))) { ctx->err = "t2b200_fft: n must be a power of two from 256 to 32768"; return T2B200_ERR_ARG; }
  if (batch == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc; const void* din; void *dout, *dtmp;
  const size_t bytes = (size_t)batch * n * sizeof(float2);
  if ((rc = t2_to_device(ctx, 0, in, bytes, &din))) return rc;
  if ((rc = t2_out_device(ctx, 1, out, bytes, &dout))) return rc;
  const int chunk = std::max(1, (int)((48u << 20) / ((size_t)n * sizeof(float2) * 2)));
  if ((rc = t2_dev_scratch(ctx, 4, (size_t)std::min(chunk, batch) * n * sizeof(float2), &dtmp))) return rc;
  if ((rc = t2_fft_device(ctx, n, (const float2*)din, batch, (float2*)dout, (float2*)dtmp, nullptr, 1.0f))) return rc;
  return t2_finish_out(ctx, out, dout, bytes);
}
