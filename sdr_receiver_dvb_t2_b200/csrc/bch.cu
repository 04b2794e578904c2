// N3, opt-in: true BCH decoding of the DVB-T2 outer code on the GPU -- syndromes, Berlekamp-Massey, Chien search,
// correction of up to t bit errors per BBFRAME.  The reference stops at "TODO BCH decode" (bch_decoder.cpp:136): it strips
// the parity bits unread, so with this stage off (the default) the engine is bit-identical to it; with it on, the residual
// errors the LDPC decoder leaves behind (its error floor) are removed the way EN 302 755 6.1.1 intends.
//
// Code: shortened binary BCH over GF(2^16) (64 800-bit FECFRAMEs; t = 12, or t = 10 for rates 2/3 and 5/6) or GF(2^14)
// (16 200-bit FECFRAMEs; t = 12), N_bch = K_ldpc, generator = product of the minimal polynomials of alpha, alpha^3, ...,
// alpha^(2t-1) with alpha a root of 1 + x^2 + x^3 + x^5 + x^16 resp. 1 + x + x^3 + x^5 + x^14 (tables 6a / 6b).  Bit 0 of a
// word is the coefficient of x^(N_bch - 1).
//
// One CTA per BBFRAME.  The syndromes S_1, S_3, ..., S_(2t-1) are accumulated by all threads over the set bits (log /
// antilog tables, L2-resident), the even ones are squares; a zero syndrome vector -- the usual case -- ends the CTA after
// this single pass over the word.  Otherwise one thread runs Berlekamp-Massey (t <= 12: a few hundred field operations) and
// all threads evaluate the locator polynomial at every position of the shortened code.
#include "stages.h"
#include <vector>

struct BchField { int m = 0, n = 0; uint16_t* d_exp = nullptr; uint16_t* d_log = nullptr; };
struct BchState { BchField f[2]; };      // [0]: GF(2^16), [1]: GF(2^14)

namespace {

constexpr int kBchThreads = 256;
constexpr int kMaxT = 12;

__device__ __forceinline__ int mod_n(unsigned x, int m, int n)
{
  x = (x & (unsigned)n) + (x >> m);      // n = 2^m - 1
  x = (x & (unsigned)n) + (x >> m);
  return x >= (unsigned)n ? (int)(x - n) : (int)x;
}
__device__ __forceinline__ unsigned gf_mul(unsigned a, unsigned b, const uint16_t* __restrict__ ex, const uint16_t* __restrict__ lg)
{
  return (a && b) ? ex[lg[a] + lg[b]] : 0u;      // the antilog table has 2n entries
}

__global__ void __launch_bounds__(kBchThreads) bch_decode_kernel(uint8_t* __restrict__ bits, int row_stride, int n_bch, int t, int m,
                                                                  const uint16_t* __restrict__ ex, const uint16_t* __restrict__ lg,
                                                                  int32_t* __restrict__ corrected)
{
  const int n = (1 << m) - 1;
  uint8_t* word = bits + (size_t)blockIdx.x * row_stride;
  __shared__ unsigned s_syn[2 * kMaxT + 1];        // S_1 .. S_2t at [1 .. 2t]
  __shared__ unsigned s_red[kBchThreads / 32][kMaxT];
  __shared__ unsigned s_sigma_log[kMaxT + 1];      // log of the locator coefficients (0xffff: coefficient is zero)
  __shared__ int s_L, s_found, s_pos[kMaxT + 4];
  const int tid = threadIdx.x;
  // ---- odd syndromes over the set bits ----
  unsigned acc[kMaxT];
#pragma unroll
  for (int k = 0; k < kMaxT; ++k) acc[k] = 0;
  for (int i0 = 4 * tid; i0 < n_bch; i0 += 4 * kBchThreads) {                 // n_bch is a multiple of 8
    const uint32_t w = *reinterpret_cast<const uint32_t*>(word + i0);
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if ((w >> (8 * b)) & 1u) {
        const unsigned p = (unsigned)(n_bch - 1 - (i0 + b));
#pragma unroll
        for (int k = 0; k < kMaxT; ++k)
          if (k < t) acc[k] ^= ex[mod_n((2 * k + 1) * p, m, n)];
      }
  }
#pragma unroll
  for (int k = 0; k < kMaxT; ++k) {
    unsigned v = acc[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[tid >> 5][k] = v;
  }
  __syncthreads();
  if (tid < t) {
    unsigned v = 0;
    for (int w = 0; w < kBchThreads / 32; ++w) v ^= s_red[w][tid];
    s_syn[2 * tid + 1] = v;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned any = 0;
    for (int j = 2; j <= 2 * t; j += 2) { const unsigned h = s_syn[j >> 1]; s_syn[j] = gf_mul(h, h, ex, lg); }   // S_2j = S_j^2
    for (int j = 1; j <= 2 * t; ++j) any |= s_syn[j];
    s_L = any ? -2 : 0;
    s_found = 0;
    if (any) {
      // Berlekamp-Massey: sigma(x) = 1 + sigma_1 x + ... + sigma_L x^L
      unsigned Cc[2 * kMaxT + 2], Bb[2 * kMaxT + 2], Tt[2 * kMaxT + 2];
      for (int i = 0; i < 2 * kMaxT + 2; ++i) { Cc[i] = 0; Bb[i] = 0; }
      Cc[0] = 1; Bb[0] = 1;
      int L = 0, mm = 1; unsigned b = 1;
      for (int it = 0; it < 2 * t; ++it) {
        unsigned d = s_syn[it + 1];
        for (int i = 1; i <= L; ++i) d ^= gf_mul(Cc[i], s_syn[it + 1 - i], ex, lg);
        if (!d) { ++mm; continue; }
        const unsigned coef = ex[lg[d] + n - lg[b]];
        for (int i = 0; i < 2 * kMaxT + 2; ++i) Tt[i] = Cc[i];
        for (int i = 0; i + mm <= 2 * t; ++i) Cc[i + mm] ^= gf_mul(coef, Bb[i], ex, lg);
        if (2 * L <= it) { L = it + 1 - L; for (int i = 0; i < 2 * kMaxT + 2; ++i) Bb[i] = Tt[i]; b = d; mm = 1; } else ++mm;
      }
      if (L <= t) {
        s_L = L;
        for (int i = 0; i <= L; ++i) s_sigma_log[i] = Cc[i] ? lg[Cc[i]] : 0xffffu;
      } else s_L = -1;
    }
  }
  __syncthreads();
  const int L = s_L;
  if (L == 0) { if (tid == 0 && corrected) corrected[blockIdx.x] = 0; return; }
  if (L < 0) { if (tid == 0 && corrected) corrected[blockIdx.x] = -1; return; }
  // ---- Chien search over the positions of the shortened code: an error at power p <=> sigma(alpha^-p) = 0 ----
  for (int p = tid; p < n_bch; p += kBchThreads) {
    unsigned v = 1;
    const unsigned q = (unsigned)(n - mod_n((unsigned)p, m, n));
    for (int i = 1; i <= L; ++i) {
      const unsigned sl = s_sigma_log[i];
      if (sl != 0xffffu) v ^= ex[sl + mod_n((unsigned)i * q, m, n)];
    }
    if (!v) { const int k = atomicAdd(&s_found, 1); if (k < kMaxT + 4) s_pos[k] = p; }
  }
  __syncthreads();
  if (s_found == L) {
    if (tid < L) word[n_bch - 1 - s_pos[tid]] ^= 1;
    if (tid == 0 && corrected) corrected[blockIdx.x] = L;
  } else if (tid == 0 && corrected) corrected[blockIdx.x] = -1;      // more than t errors: the word is left as it is
}

}  // namespace

void t2_bch_free(t2b200_ctx* ctx)
{
  if (!ctx->bch) return;
  for (auto& f : ctx->bch->f) { cudaFree(f.d_exp); cudaFree(f.d_log); }
  delete ctx->bch;
  ctx->bch = nullptr;
}

static int bch_field(t2b200_ctx* ctx, bool short_frame, const BchField** out)
{
  if (!ctx->bch) ctx->bch = new BchState();
  BchField& f = ctx->bch->f[short_frame ? 1 : 0];
  if (!f.d_exp) {
    f.m = short_frame ? 14 : 16;
    f.n = (1 << f.m) - 1;
    const int prim = short_frame ? ((1 << 14) | (1 << 5) | (1 << 3) | (1 << 1) | 1) : ((1 << 16) | (1 << 5) | (1 << 3) | (1 << 2) | 1);
    std::vector<uint16_t> ex(2 * (size_t)f.n + 2), lg((size_t)f.n + 1, 0);
    int x = 1;
    for (int i = 0; i < f.n; ++i) {
      ex[i] = (uint16_t)x; lg[x] = (uint16_t)i;
      x <<= 1;
      if (x >> f.m) x ^= prim;
    }
    for (size_t i = f.n; i < ex.size(); ++i) ex[i] = ex[i - f.n];
    T2_CUDA(ctx, cudaMalloc(&f.d_exp, ex.size() * sizeof(uint16_t)));
    T2_CUDA(ctx, cudaMalloc(&f.d_log, lg.size() * sizeof(uint16_t)));
    T2_CUDA(ctx, cudaMemcpy(f.d_exp, ex.data(), ex.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    T2_CUDA(ctx, cudaMemcpy(f.d_log, lg.data(), lg.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  }
  *out = &f;
  return T2B200_OK;
}

extern "C" int t2b200_bch_t(int code) { return code < 0 || code > 11 ? 0 : (code == 2 || code == 5) ? 10 : 12; }

// device-level: bits uint8[n_words][row_stride] (one byte per bit, the first K_ldpc of each row are the BCH word), in place
int t2_bch_device(t2b200_ctx* ctx, int code, uint8_t* d_bits, int row_stride, int n_words, int32_t* d_corrected)
{
  const int t = t2b200_bch_t(code), n_bch = t2b200_ldpc_k(code);
  if (!t || !n_bch) { ctx->err = "code has no BCH geometry"; return T2B200_ERR_ARG; }
  if (n_words == 0) return T2B200_OK;
  const BchField* f; int rc;
  if ((rc = bch_field(ctx, code >= 6, &f))) return rc;
  bch_decode_kernel<<<n_words, kBchThreads, 0, ctx->stream>>>(d_bits, row_stride, n_bch, t, f->m, f->d_exp, f->d_log, d_corrected);
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return T2B200_OK;
}

extern "C" int t2b200_bch_decode(t2b200_ctx* ctx, int code, uint8_t* bits_inout, int n_words, int32_t* corrected)
{
  if (!ctx) return T2B200_ERR_ARG;
  const int k_ldpc = t2b200_ldpc_k(code);
  if (!bits_inout || n_words < 0 || !t2b200_bch_t(code)) { ctx->err = "t2b200_bch_decode: bad argument"; return T2B200_ERR_ARG; }
  if (n_words == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc; const void* din; void* dcor = nullptr;
  const size_t bytes = (size_t)n_words * k_ldpc;
  const bool host = !t2_is_device_ptr(bits_inout);
  if ((rc = t2_to_device(ctx, 0, bits_inout, bytes, &din))) return rc;
  if (corrected && (rc = t2_out_device(ctx, 2, corrected, 4 * (size_t)n_words, &dcor))) return rc;
  if ((rc = t2_bch_device(ctx, code, (uint8_t*)din, k_ldpc, n_words, (int32_t*)dcor))) return rc;
  if (host && (rc = t2_finish_out(ctx, bits_inout, din, bytes))) return rc;
  if (corrected && (rc = t2_finish_out(ctx, corrected, dcor, 4 * (size_t)n_words))) return rc;
  return T2B200_OK;
}
