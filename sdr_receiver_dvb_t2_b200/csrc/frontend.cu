// N2: the receiver front-end in front of the FFT -- int16 I/Q -> DC removal -> IQ-imbalance correction -> NCO
// derotation -> Farrow resampler -> half-band decimator -- for a batch of independent streams, one chunk (about one OFDM
// symbol) per stream per call, plus the guard-interval correlation.  Replaces the per-sample loop of
// dvbt2_demodulator::execute (dvbt2_demodulator.cpp:178-221), DSP/interpolator_farrow.hh:41-68, DSP/filter_decimator.h:72-131
// and dvbt2_demodulator.cpp:321-330.  The kernel bodies (and the description of how the reference's per-sample recurrences
// become parallel) are in frontend_kernels.h, which also compiles for the host: tests/test_frontend_emu.py runs it on the
// CPU against the oracle.
//
// Per call: fe_dc_partial_kernel (one CTA per 1 024 input samples: weighted tile sums of the DC average) ->
// fe_plan_kernel (one CTA per stream, the walk itself one thread: state at every tile boundary, output counts) -> fe_derotate_kernel (one CTA per
// 1 024 input samples) -> fe_resample_kernel (one CTA per 1 024 outputs + one per stream that commits the carried state).
// Algorithmic traffic per input sample at resample = 0.5: 4 B in, 8 B out; the derotated samples (8 B) make one round trip
// through L2 between the two passes.
#include "ctx.h"
#include "frontend_tables.h"
#include <cstring>

static_assert(sizeof(FeStream) == sizeof(t2b200_fe_state), "t2b200_fe_state mirrors FeStream");
static_assert(sizeof(FeChunk) == sizeof(t2b200_fe_chunk), "t2b200_fe_chunk mirrors FeChunk");
static_assert(sizeof(FeResult) == sizeof(t2b200_fe_result), "t2b200_fe_result mirrors FeResult");

struct FeState {
  int n_streams = 0, max_chunk = 0, cur = 0;
  FeStream* d_state[2] = {nullptr, nullptr};
  FeChunk* d_chunk = nullptr;
  FePlan* d_plan = nullptr;
  double2* d_dc_part = nullptr;
  double* d_theta_part = nullptr;
  float2* d_derot = nullptr;
  FeResult* d_result = nullptr;
  double* d_apow = nullptr; double* d_ainv = nullptr; float* d_h = nullptr;
  FeChunk* h_chunk = nullptr; FeResult* h_result = nullptr;     // pinned
};

__global__ void __launch_bounds__(FE_THREADS) fe_dc_partial_kernel(FeArgs A) { fe_dc_partial_body(A, blockIdx.y, blockIdx.x); }
__global__ void __launch_bounds__(64) fe_plan_kernel(FeArgs A) { fe_plan_body(A, blockIdx.x); }
__global__ void __launch_bounds__(FE_THREADS) fe_derotate_kernel(FeArgs A) { fe_derotate_body(A, blockIdx.y, blockIdx.x); }
__global__ void __launch_bounds__(FE_THREADS, 4) fe_resample_kernel(FeArgs A) { fe_resample_body(A, blockIdx.y, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(FE_THREADS) fe_cp_correlate_kernel(const float2* sym, long long stride, int fft_size, int guard, float* est)
{
  fe_cp_correlate_body(sym + blockIdx.x * stride, fft_size, guard, est + blockIdx.x);
}

__global__ void __launch_bounds__(FE_P1_THREADS) fe_p1_correlate_kernel(const float2* x, int n, const float2* hist, const float2* fq, int i0,
                                                                        double2* prefix, float* correlation, float2* out)
{
  fe_p1_correlate_body(x, n, hist, fq, i0, prefix, correlation, out);
}

void t2_fe_free(t2b200_ctx* ctx)
{
  if (ctx->d_p1_fq) { cudaFree(ctx->d_p1_fq); ctx->d_p1_fq = nullptr; }
  FeState* f = ctx->fe;
  if (!f) return;
  for (auto& p : f->d_state) if (p) cudaFree(p);
  if (f->d_chunk) cudaFree(f->d_chunk);
  if (f->d_plan) cudaFree(f->d_plan);
  if (f->d_dc_part) cudaFree(f->d_dc_part);
  if (f->d_theta_part) cudaFree(f->d_theta_part);
  if (f->d_derot) cudaFree(f->d_derot);
  if (f->d_result) cudaFree(f->d_result);
  if (f->d_apow) cudaFree(f->d_apow);
  if (f->d_ainv) cudaFree(f->d_ainv);
  if (f->d_h) cudaFree(f->d_h);
  if (f->h_chunk) cudaFreeHost(f->h_chunk);
  if (f->h_result) cudaFreeHost(f->h_result);
  delete f;
  ctx->fe = nullptr;
}

static void fresh(FeStream& s) { std::memset(&s, 0, sizeof(s)); s.x1 = -0.5f; }     // interpolator_farrow.hh:33-36

extern "C" int t2b200_frontend_configure(t2b200_ctx* ctx, int n_streams, int max_chunk_in)
{
  if (!ctx || n_streams < 1 || n_streams > 65535 || max_chunk_in < 1 || max_chunk_in > FE_MAX_TILES * FE_TILE_IN) {
    if (ctx) ctx->err = "t2b200_frontend_configure: 1 <= n_streams <= 65535, 1 <= max_chunk_in <= 262144";
    return T2B200_ERR_ARG;
  }
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t2_fe_free(ctx);
  int rc;
  if ((rc = t2_ensure_lut(ctx))) return rc;
  FeState* f = new FeState();
  ctx->fe = f;
  // a failed allocation leaves no half-built state behind
#define FE_ALLOC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); \
                                                                              cudaGetLastError(); t2_fe_free(ctx); return e__ == cudaErrorMemoryAllocation ? T2B200_ERR_NOMEM : T2B200_ERR_CUDA; } } while (0)
  f->n_streams = n_streams; f->max_chunk = max_chunk_in;
  const size_t S = (size_t)n_streams;
  for (auto& p : f->d_state) FE_ALLOC(cudaMalloc(&p, S * sizeof(FeStream)));
  FE_ALLOC(cudaMalloc(&f->d_chunk, S * sizeof(FeChunk)));
  FE_ALLOC(cudaMalloc(&f->d_plan, S * sizeof(FePlan)));
  FE_ALLOC(cudaMalloc(&f->d_dc_part, S * FE_MAX_TILES * sizeof(double2)));
  FE_ALLOC(cudaMalloc(&f->d_theta_part, S * FE_MAX_TILES * 3 * sizeof(double)));
  FE_ALLOC(cudaMalloc(&f->d_derot, S * ((size_t)max_chunk_in + 4) * sizeof(float2)));
  FE_ALLOC(cudaMalloc(&f->d_result, S * sizeof(FeResult)));
  FE_ALLOC(cudaMallocHost(&f->h_chunk, S * sizeof(FeChunk)));
  FE_ALLOC(cudaMallocHost(&f->h_result, S * sizeof(FeResult)));
  std::vector<double> apow, ainv; std::vector<float> lut, h;
  fe_make_tables(apow, ainv, lut, h);
  FE_ALLOC(cudaMalloc(&f->d_apow, apow.size() * sizeof(double)));
  FE_ALLOC(cudaMalloc(&f->d_ainv, ainv.size() * sizeof(double)));
  FE_ALLOC(cudaMalloc(&f->d_h, h.size() * sizeof(float)));
  FE_ALLOC(cudaMemcpy(f->d_apow, apow.data(), apow.size() * sizeof(double), cudaMemcpyHostToDevice));
  FE_ALLOC(cudaMemcpy(f->d_ainv, ainv.data(), ainv.size() * sizeof(double), cudaMemcpyHostToDevice));
  FE_ALLOC(cudaMemcpy(f->d_h, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  FE_ALLOC(cudaMemcpyToSymbol(fe_c_h, h.data(), h.size() * sizeof(float)));
#undef FE_ALLOC
  return t2b200_frontend_reset(ctx, -1);
}

extern "C" int t2b200_frontend_reset(t2b200_ctx* ctx, int stream)
{
  if (!ctx) return T2B200_ERR_ARG;
  FeState* f = ctx->fe;
  if (!f) { ctx->err = "front-end not configured"; return T2B200_ERR_STATE; }
  if (stream < -1 || stream >= f->n_streams) { ctx->err = "t2b200_frontend_reset: no such stream"; return T2B200_ERR_ARG; }
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  FeStream z; fresh(z);
  for (int s = (stream < 0 ? 0 : stream); s < (stream < 0 ? f->n_streams : stream + 1); ++s)
    T2_CUDA(ctx, cudaMemcpy(f->d_state[f->cur] + s, &z, sizeof(z), cudaMemcpyHostToDevice));
  return T2B200_OK;
}

extern "C" int t2b200_frontend_get_state(t2b200_ctx* ctx, int stream, t2b200_fe_state* state)
{
  if (!ctx || !state) return T2B200_ERR_ARG;
  FeState* f = ctx->fe;
  if (!f) { ctx->err = "front-end not configured"; return T2B200_ERR_STATE; }
  if (stream < 0 || stream >= f->n_streams) { ctx->err = "t2b200_frontend_get_state: no such stream"; return T2B200_ERR_ARG; }
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  T2_CUDA(ctx, cudaMemcpy(state, f->d_state[f->cur] + stream, sizeof(FeStream), cudaMemcpyDeviceToHost));
  return T2B200_OK;
}

extern "C" int t2b200_frontend_set_state(t2b200_ctx* ctx, int stream, const t2b200_fe_state* state)
{
  if (!ctx || !state) return T2B200_ERR_ARG;
  FeState* f = ctx->fe;
  if (!f) { ctx->err = "front-end not configured"; return T2B200_ERR_STATE; }
  if (stream < 0 || stream >= f->n_streams) { ctx->err = "t2b200_frontend_set_state: no such stream"; return T2B200_ERR_ARG; }
  // what the kernels rely on: the resampler phase the reference can leave behind (x1 in [-0.5, 0.5 + resample), resample <= 1:
  // the output rows are sized for it), the decimator phase 0 / 1, an NCO phase inside its wrap range
  if (!(state->x1 >= -0.5f && state->x1 < 1.5f) || (state->parity != 0 && state->parity != 1) ||
      !(state->frequency_nco >= -FE_TWO_PI_F && state->frequency_nco <= FE_TWO_PI_F) || !(state->dc_re == state->dc_re) ||
      !(state->dc_im == state->dc_im)) {
    ctx->err = "t2b200_frontend_set_state: x1 outside [-0.5, 1.5), parity not 0 / 1, or NCO phase outside +-2 pi";
    return T2B200_ERR_ARG;
  }
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  T2_CUDA(ctx, cudaMemcpy(f->d_state[f->cur] + stream, state, sizeof(FeStream), cudaMemcpyHostToDevice));
  return T2B200_OK;
}

extern "C" int t2b200_frontend_execute(t2b200_ctx* ctx, const int16_t* i_in, const int16_t* q_in, long long stream_stride,
                                       int sample_step, const t2b200_fe_chunk* chunk, float* out, long long out_stride,
                                       t2b200_fe_result* result)
{
  if (!ctx || !i_in || !q_in || !chunk || !out || !result || sample_step < 1 || stream_stride < 0 || out_stride < 0) {
    if (ctx) ctx->err = "t2b200_frontend_execute: bad argument";
    return T2B200_ERR_ARG;
  }
  FeState* f = ctx->fe;
  if (!f) { ctx->err = "front-end not configured"; return T2B200_ERR_STATE; }
  const int S = f->n_streams;
  int max_in = 0;
  for (int s = 0; s < S; ++s) {
    const t2b200_fe_chunk& c = chunk[s];
    // resample: the reference keeps it within 100 ppm of sample_rate / (2 * 64/7 MHz) (dvbt2_demodulator.cpp:55-56); the
    // resampling pass sizes its tiles for 0.25 <= resample <= 1
    if (c.len_in < 0 || c.len_in > f->max_chunk || !(c.resample >= 0.25f && c.resample <= 1.0f)) {
      ctx->err = "t2b200_frontend_execute: chunk longer than configured, or resample outside [0.25, 1]";
      return T2B200_ERR_ARG;
    }
    if (c.len_in > max_in) max_in = c.len_in;
  }
  const int nt_out = fe_max_out_tiles(reinterpret_cast<const FeChunk*>(chunk), S);
  // the worst-case output of every stream must fit its row
  for (int s = 0; s < S; ++s) {
    const long long worst = (long long)(((double)chunk[s].len_in + 1.0) / (double)chunk[s].resample + 2.0) / 2 + 1;
    if (chunk[s].len_in > 0 && worst > out_stride) { ctx->err = "t2b200_frontend_execute: out_stride too small for this chunk"; return T2B200_ERR_ARG; }
  }
  int rc;
  const void* d_i; const void* d_q;
  const size_t span = (S > 0 && max_in > 0) ? ((size_t)(S - 1) * (size_t)stream_stride + (size_t)(max_in - 1) * sample_step + 1) * sizeof(int16_t) : 0;
  if (max_in > 0) {
    if ((rc = t2_to_device(ctx, 13, i_in, span, &d_i))) return rc;
    if ((rc = t2_to_device(ctx, 14, q_in, span, &d_q))) return rc;
  } else { d_i = i_in; d_q = q_in; }
  void* d_out;
  const size_t out_bytes = (size_t)S * (size_t)out_stride * sizeof(float2);
  if ((rc = t2_out_device(ctx, 15, out, out_bytes, &d_out))) return rc;
  // the previous call's results were read before it returned, so the pinned staging buffers are free
  std::memcpy(f->h_chunk, chunk, (size_t)S * sizeof(FeChunk));
  T2_CUDA(ctx, cudaMemcpyAsync(f->d_chunk, f->h_chunk, (size_t)S * sizeof(FeChunk), cudaMemcpyHostToDevice, ctx->stream));
  FeArgs A;
  A.i_in = static_cast<const int16_t*>(d_i); A.q_in = static_cast<const int16_t*>(d_q);
  A.in_stride = stream_stride; A.step = sample_step;
  A.chunk = f->d_chunk; A.cur = f->d_state[f->cur]; A.next = f->d_state[f->cur ^ 1];
  A.plan = f->d_plan; A.dc_part = f->d_dc_part; A.theta_part = f->d_theta_part;
  A.derot = f->d_derot; A.derot_stride = f->max_chunk + 4;
  A.out = static_cast<float2*>(d_out); A.out_stride = out_stride;
  A.result = f->d_result; A.apow = f->d_apow; A.ainv = f->d_ainv;
  A.lut_cs = reinterpret_cast<const float2*>(ctx->d_lut); A.h = f->d_h;
  const int nt_in = max_in > 0 ? (max_in + FE_TILE_IN - 1) / FE_TILE_IN : 1;     // tile 0 always runs: it lays out the resampler's delay line
  if (nt_in > 0) {
    fe_dc_partial_kernel<<<dim3(nt_in, S), FE_THREADS, 0, ctx->stream>>>(A);
    ctx->launches++;
  }
  fe_plan_kernel<<<S, 64, 0, ctx->stream>>>(A);
  ctx->launches++;
  if (nt_in > 0) {
    fe_derotate_kernel<<<dim3(nt_in, S), FE_THREADS, 0, ctx->stream>>>(A);
    ctx->launches++;
  }
  fe_resample_kernel<<<dim3(nt_out + 1, S), FE_THREADS, 0, ctx->stream>>>(A);
  ctx->launches++;
  T2_CUDA(ctx, cudaGetLastError());
  f->cur ^= 1;
  T2_CUDA(ctx, cudaMemcpyAsync(f->h_result, f->d_result, (size_t)S * sizeof(FeResult), cudaMemcpyDeviceToHost, ctx->stream));
  if (out != d_out) T2_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));             // the host loop needs len_out before it can go on
  std::memcpy(result, f->h_result, (size_t)S * sizeof(FeResult));
  return T2B200_OK;
}

extern "C" int t2b200_cp_correlate(t2b200_ctx* ctx, const float* symbols, int n_symbols, long long symbol_stride, int fft_size,
                                   int guard, float* frequency_est)
{
  if (!ctx || !symbols || !frequency_est || n_symbols < 0 || fft_size < 1 || guard < 0 || symbol_stride < (long long)fft_size + guard) {
    if (ctx) ctx->err = "t2b200_cp_correlate: bad argument";
    return T2B200_ERR_ARG;
  }
  if (n_symbols == 0) return T2B200_OK;
  int rc;
  const void* d_sym; void* d_est;
  const size_t bytes = ((size_t)(n_symbols - 1) * symbol_stride + fft_size + guard) * sizeof(float2);
  if ((rc = t2_to_device(ctx, 13, symbols, bytes, &d_sym))) return rc;
  if ((rc = t2_out_device(ctx, 15, frequency_est, (size_t)n_symbols * sizeof(float), &d_est))) return rc;
  fe_cp_correlate_kernel<<<n_symbols, FE_THREADS, 0, ctx->stream>>>(static_cast<const float2*>(d_sym), symbol_stride, fft_size, guard,
                                                                    static_cast<float*>(d_est));
  ctx->launches++;
  T2_CUDA(ctx, cudaGetLastError());
  return t2_finish_out(ctx, frequency_est, d_est, (size_t)n_symbols * sizeof(float));
}

extern "C" int t2b200_p1_correlate(t2b200_ctx* ctx, const float* samples, int n, const float* history, int fq_index,
                                   float* correlation, float* out)
{
  if (!ctx || !samples || !correlation || n < 0 || n > (1 << 22)) {
    if (ctx) ctx->err = "t2b200_p1_correlate: bad argument";
    return T2B200_ERR_ARG;
  }
  if (n == 0) return T2B200_OK;
  if (!ctx->d_p1_fq) {
    std::vector<float> fq; fe_make_p1_table(fq);
    T2_CUDA(ctx, cudaMalloc(&ctx->d_p1_fq, fq.size() * sizeof(float)));
    T2_CUDA(ctx, cudaMemcpy(ctx->d_p1_fq, fq.data(), fq.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  int rc;
  const void *d_x, *d_h = nullptr; void *d_c, *d_o = nullptr, *d_p;
  if ((rc = t2_to_device(ctx, 13, samples, (size_t)n * sizeof(float2), &d_x))) return rc;
  if (history && (rc = t2_to_device(ctx, 14, history, (size_t)FE_P1_HISTORY * sizeof(float2), &d_h))) return rc;
  if ((rc = t2_out_device(ctx, 15, correlation, (size_t)n * sizeof(float), &d_c))) return rc;
  if (out && (rc = t2_out_device(ctx, 16, out, (size_t)n * sizeof(float2), &d_o))) return rc;
  if ((rc = t2_dev_scratch(ctx, 17, 2 * ((size_t)n + FE_P1_LEAD + 1) * sizeof(double2), &d_p))) return rc;
  fe_p1_correlate_kernel<<<1, FE_P1_THREADS, 0, ctx->stream>>>(static_cast<const float2*>(d_x), n, static_cast<const float2*>(d_h),
                                                               reinterpret_cast<const float2*>(ctx->d_p1_fq), fq_index & 1023,
                                                               static_cast<double2*>(d_p), static_cast<float*>(d_c), static_cast<float2*>(d_o));
  ctx->launches++;
  T2_CUDA(ctx, cudaGetLastError());
  if (out && (rc = t2_finish_out(ctx, out, d_o, (size_t)n * sizeof(float2)))) return rc;
  return t2_finish_out(ctx, correlation, d_c, (size_t)n * sizeof(float));
}
