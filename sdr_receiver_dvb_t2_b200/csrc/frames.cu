// Frame pipeline: whole T2 frames of one PLP through the hot path in ONE call, every stage chained on the device --
// FFT -> equalise / frequency de-interleave (P2, data and frame-closing symbols written straight into the frame cell
// stream) -> time / cell de-interleave -> demap -> LDPC + BCH strip / descramble.  It is the replay ("teacher-forced")
// use of the per-stage entry points (dvbt2_demodulator.cpp:332-385 feeding time_deinterleaver -> llr_demapper ->
// ldpc_decoder -> bch_decoder): the caller supplies the FFT windows of already synchronised frames.  Nothing in the call
// waits for the GPU when all buffers are device (or, for the IQ input, pinned host) memory.
#include "stages.h"
#include "fec_tables.h"
#include <algorithm>
#include <vector>

struct FramePipe {
  t2b200_frame_cfg cfg{};
  bool configured = false;
  int n_data = 0, per_frame = 0, cpf = 0, fec_bits = 0, code = 0, k_bch = 0;
  std::vector<int> blocks;                 // FEC blocks per TI block of one frame (time_deinterleaver.cpp:275-282)
  int frames_cap = 0;                      // descriptors / buffers are sized for this many frames
  TiBlockDesc* d_ti = nullptr; DemapBlockDesc* d_dm = nullptr;
  float2 *d_freq = nullptr, *d_tmp = nullptr, *d_cells = nullptr, *d_tib = nullptr, *d_iq = nullptr;
  int8_t* d_llr = nullptr; float* d_prec = nullptr; uint8_t* d_bits = nullptr; int32_t* d_trials = nullptr;
  uint8_t* d_kldpc = nullptr;              // LDPC information words of the opt-in BCH correction
  float* d_fb = nullptr;                   // sro | phase, [frames][len_frame] each
  int max_cells = 0, max_fec = 0;
  cudaEvent_t ev[7] = {};                  // stage boundaries of the last call (T2B200_OPT_STAGE_TIMING)
  bool timed = false;
};

static void pipe_free_buffers(FramePipe* p)
{
  cudaFree(p->d_ti); cudaFree(p->d_dm); cudaFree(p->d_freq); cudaFree(p->d_tmp); cudaFree(p->d_cells); cudaFree(p->d_tib);
  cudaFree(p->d_kldpc); p->d_kldpc = nullptr;
  cudaFree(p->d_iq); cudaFree(p->d_llr); cudaFree(p->d_prec); cudaFree(p->d_bits); cudaFree(p->d_trials); cudaFree(p->d_fb);
  p->d_ti = nullptr; p->d_dm = nullptr; p->d_freq = p->d_tmp = p->d_cells = p->d_tib = p->d_iq = nullptr;
  p->d_llr = nullptr; p->d_prec = nullptr; p->d_bits = nullptr; p->d_trials = nullptr; p->d_fb = nullptr;
  p->frames_cap = 0;
}

void t2_frames_free(t2b200_ctx* ctx)
{
  if (!ctx->frames) return;
  pipe_free_buffers(ctx->frames);
  for (auto& e : ctx->frames->ev) if (e) cudaEventDestroy(e);
  delete ctx->frames;
  ctx->frames = nullptr;
}

extern "C" int t2b200_frames_configure(t2b200_ctx* ctx, const t2b200_frame_cfg* c)
{
  if (!ctx || !c) return T2B200_ERR_ARG;
  if (c->fft_size < 4096 || c->len_frame < 1 || c->n_p2 < 1 || c->n_p2 > c->len_frame || c->n_blocks < 1 || c->ti_len < 1 ||
      c->ti_len > c->n_blocks || c->mod < 0 || c->mod > 3 || c->code_rate < 0 || c->code_rate > 5 || c->first_cell < 0) {
    ctx->err = "t2b200_frames_configure: bad argument"; return T2B200_ERR_ARG;
  }
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->sym[T2B200_SYM_P2] || !ctx->sym[T2B200_SYM_DATA] || (c->l_fc && !ctx->sym[T2B200_SYM_FC])) {
    ctx->err = "t2b200_frames_configure: call t2b200_eq_configure for every symbol kind first"; return T2B200_ERR_STATE;
  }
  {
    // the equaliser takes its input / output strides from this configuration but l_nulls, the carrier maps and the
    // de-interleaver addresses (up to n_out - 1) from the symbol tables: they must describe the same mode
    const int n_data = c->len_frame - c->n_p2 - (c->l_fc ? 1 : 0);
    const int want_out[3] = {c->c_p2, c->c_data, c->n_fc}, want_sym[3] = {c->n_p2, n_data, 1};
    const int want_first[3] = {0, c->n_p2, c->len_frame - 1};
    for (int kind = 0; kind < 3; ++kind) {
      if (kind == T2B200_SYM_FC && !c->l_fc) continue;
      if (kind == T2B200_SYM_DATA && n_data == 0) continue;
      int fs, no, ns, first;
      if (!t2_eq_geometry(ctx, kind, &fs, &no, &ns, &first) || fs != c->fft_size || no != want_out[kind] ||
          ns < want_sym[kind] || first != want_first[kind]) {
        ctx->err = "t2b200_frames_configure: frame geometry differs from what t2b200_eq_configure was given (fft_size, cells per "
                   "symbol, symbols per frame)";
        return T2B200_ERR_ARG;
      }
    }
  }
  int cpf = 0, nmax = 0, rc;
  if ((rc = t2_ti_geometry(ctx, c->plp, &cpf, &nmax))) return rc;
  if (!ctx->frames) ctx->frames = new FramePipe();
  FramePipe* p = ctx->frames;
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pipe_free_buffers(p);
  p->cfg = *c;
  p->n_data = c->len_frame - c->n_p2 - (c->l_fc ? 1 : 0);
  p->per_frame = c->n_p2 * c->c_p2 + p->n_data * c->c_data + (c->l_fc ? c->n_fc : 0);
  p->cpf = cpf; p->fec_bits = c->fec_type ? 64800 : 16200;
  p->code = t2b200_ldpc_code_id(c->fec_type, c->code_rate);
  p->k_bch = t2b200_ldpc_k_bch(p->code);
  p->blocks.clear();
  const int base = c->n_blocks / c->ti_len;                                  // time_deinterleaver.cpp:275-282
  for (int j = 0; j < c->ti_len; ++j) p->blocks.push_back(base + (j >= c->ti_len - c->n_blocks % c->ti_len ? 1 : 0));
  p->max_fec = *std::max_element(p->blocks.begin(), p->blocks.end());
  p->max_cells = p->max_fec * cpf;
  if (p->max_fec > nmax) { ctx->err = "t2b200_frames_configure: TI blocks larger than t2b200_ti_configure allowed"; return T2B200_ERR_ARG; }
  if ((long long)c->first_cell + (long long)c->n_blocks * cpf > p->per_frame) { ctx->err = "t2b200_frames_configure: PLP does not fit the frame"; return T2B200_ERR_ARG; }
  if (cpf != t2_cells_per_fec(c->fec_type, c->mod)) { ctx->err = "t2b200_frames_configure: PLP geometry differs from t2b200_ti_configure"; return T2B200_ERR_ARG; }
  p->configured = true;
  return T2B200_OK;
}

static int pipe_reserve(t2b200_ctx* ctx, FramePipe* p, int F, bool host_iq)
{
  if (F <= p->frames_cap && (!host_iq || p->d_iq)) return T2B200_OK;
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const bool keep_iq = F <= p->frames_cap;
  if (!keep_iq) pipe_free_buffers(p);
  const t2b200_frame_cfg& c = p->cfg;
  const size_t L = c.len_frame, N = c.fft_size, nb = c.n_blocks, nti = p->blocks.size();
  if (!keep_iq) {
    // descriptors of every TI block of F frames
    std::vector<TiBlockDesc> ti(F * nti); std::vector<DemapBlockDesc> dm(F * nti);
    long long out = 0; int fec = 0;
    for (int f = 0; f < F; ++f) {
      long long in = (long long)f * p->per_frame + c.first_cell;
      for (size_t b = 0; b < nti; ++b) {
        const long long cells = (long long)p->blocks[b] * p->cpf;
        ti[f * nti + b] = {in, out, p->blocks[b]};
        dm[f * nti + b] = {out, (int)cells, fec};
        in += cells; out += cells; fec += p->blocks[b];
      }
    }
    T2_CUDA(ctx, cudaMalloc(&p->d_ti, ti.size() * sizeof(TiBlockDesc)));
    T2_CUDA(ctx, cudaMalloc(&p->d_dm, dm.size() * sizeof(DemapBlockDesc)));
    T2_CUDA(ctx, cudaMemcpy(p->d_ti, ti.data(), ti.size() * sizeof(TiBlockDesc), cudaMemcpyHostToDevice));
    T2_CUDA(ctx, cudaMemcpy(p->d_dm, dm.data(), dm.size() * sizeof(DemapBlockDesc), cudaMemcpyHostToDevice));
    const int chunk = std::max(1, (int)((48u << 20) / (N * sizeof(float2) * 2)));
    T2_CUDA(ctx, cudaMalloc(&p->d_freq, (size_t)F * L * N * sizeof(float2)));
    T2_CUDA(ctx, cudaMalloc(&p->d_tmp, (size_t)std::min<size_t>(chunk, F * L) * N * sizeof(float2)));
    T2_CUDA(ctx, cudaMalloc(&p->d_cells, (size_t)F * p->per_frame * sizeof(float2)));
    T2_CUDA(ctx, cudaMalloc(&p->d_tib, (size_t)F * nb * p->cpf * sizeof(float2)));
    T2_CUDA(ctx, cudaMalloc(&p->d_llr, (size_t)F * nb * p->fec_bits));
    T2_CUDA(ctx, cudaMalloc(&p->d_prec, (size_t)F * nti * 2 * sizeof(float)));
    T2_CUDA(ctx, cudaMalloc(&p->d_bits, (size_t)F * nb * 54000));
    T2_CUDA(ctx, cudaMalloc(&p->d_trials, (size_t)F * nb * sizeof(int32_t)));
    T2_CUDA(ctx, cudaMalloc(&p->d_fb, (size_t)F * L * 2 * sizeof(float)));
    p->frames_cap = F;
  }
  if (host_iq && !p->d_iq) T2_CUDA(ctx, cudaMalloc(&p->d_iq, (size_t)p->frames_cap * L * N * sizeof(float2)));
  return T2B200_OK;
}

static int frames_decode(t2b200_ctx* ctx, const void* iq, bool i16, float scale, int n_frames, uint8_t* bits_out,
                         int32_t* trials_left, float* sro, float* phase, float* snr, int max_trials, unsigned ldpc_flags);

extern "C" int t2b200_frames_decode(t2b200_ctx* ctx, const float* iq, int n_frames, uint8_t* bits_out, int32_t* trials_left,
                                    float* sro, float* phase, float* snr, int max_trials, unsigned ldpc_flags)
{
  return frames_decode(ctx, iq, false, 1.0f, n_frames, bits_out, trials_left, sro, phase, snr, max_trials, ldpc_flags);
}

extern "C" int t2b200_frames_decode_i16(t2b200_ctx* ctx, const int16_t* iq, float scale, int n_frames, uint8_t* bits_out,
                                        int32_t* trials_left, float* sro, float* phase, float* snr, int max_trials,
                                        unsigned ldpc_flags)
{
  return frames_decode(ctx, iq, true, scale, n_frames, bits_out, trials_left, sro, phase, snr, max_trials, ldpc_flags);
}

static int frames_decode(t2b200_ctx* ctx, const void* iq, bool i16, float scale, int n_frames, uint8_t* bits_out,
                         int32_t* trials_left, float* sro, float* phase, float* snr, int max_trials, unsigned ldpc_flags)
{
  if (!ctx) return T2B200_ERR_ARG;
  FramePipe* p = ctx->frames;
  if (!p || !p->configured) { ctx->err = "t2b200_frames_decode: not configured"; return T2B200_ERR_STATE; }
  if (!iq || n_frames < 0 || !bits_out || (ldpc_flags & T2B200_LDPC_WANT_POST)) { ctx->err = "t2b200_frames_decode: bad argument"; return T2B200_ERR_ARG; }
  if (n_frames == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  const t2b200_frame_cfg& c = p->cfg;
  const int F = n_frames, L = c.len_frame, N = c.fft_size, nti = (int)p->blocks.size();
  const bool host_iq = !t2_is_device_ptr(iq);
  int rc;
  if ((rc = pipe_reserve(ctx, p, F, host_iq))) return rc;
  // the staging buffer is sized for float2 samples; int16 pairs use the first half of it
  const void* d_iq = iq;
  if (host_iq) {
    T2_CUDA(ctx, cudaMemcpyAsync(p->d_iq, iq, (size_t)F * L * N * (i16 ? sizeof(short2) : sizeof(float2)), cudaMemcpyHostToDevice,
                                 ctx->stream));
    d_iq = p->d_iq;
  }
  const bool timing = ctx->opt_stage_timing != 0;
  auto mark = [&](int i) -> int {
    if (!timing) return T2B200_OK;
    if (!p->ev[i]) T2_CUDA(ctx, cudaEventCreate(&p->ev[i]));
    T2_CUDA(ctx, cudaEventRecord(p->ev[i], ctx->stream));
    return T2B200_OK;
  };
  p->timed = timing;
  if ((rc = mark(0))) return rc;
  // K1 (int16 samples are converted on load: the first slice of the front-end, dvbt2_demodulator.cpp:182-186)
  if ((rc = t2_fft_device(ctx, N, i16 ? nullptr : static_cast<const float2*>(d_iq), F * L, p->d_freq, p->d_tmp,
                          i16 ? static_cast<const short2*>(d_iq) : nullptr, scale))) return rc;
  if ((rc = mark(1))) return rc;
  // K2: every symbol kind writes its cells where the frame cell stream wants them
  float* d_sro = p->d_fb; float* d_ph = p->d_fb + (size_t)F * L;
  // (one merged launch was measured 2x slower: the P2 symbols' dense pilots set the shared-memory footprint of every CTA)
  const long long frame_in = (long long)L * N;
  if ((rc = t2_equalize_device(ctx, T2B200_SYM_P2, F * c.n_p2, c.n_p2, nullptr, p->d_freq, frame_in, N, p->d_cells, p->per_frame,
                               c.c_p2, d_sro, d_ph, L))) return rc;
  if (p->n_data > 0 &&
      (rc = t2_equalize_device(ctx, T2B200_SYM_DATA, F * p->n_data, p->n_data, nullptr, p->d_freq + (size_t)c.n_p2 * N, frame_in, N,
                               p->d_cells + (size_t)c.n_p2 * c.c_p2, p->per_frame, c.c_data, d_sro + c.n_p2, d_ph + c.n_p2, L))) return rc;
  if (c.l_fc &&
      (rc = t2_equalize_device(ctx, T2B200_SYM_FC, F, 1, nullptr, p->d_freq + (size_t)(L - 1) * N, frame_in, N,
                               p->d_cells + (size_t)c.n_p2 * c.c_p2 + (size_t)p->n_data * c.c_data, p->per_frame, c.n_fc,
                               d_sro + (L - 1), d_ph + (L - 1), L))) return rc;
  if ((rc = mark(2))) return rc;
  // K3, K4
  if ((rc = t2_ti_device(ctx, c.plp, p->d_cells, p->d_tib, p->d_ti, F * nti, p->max_cells, c.mod, c.rotation))) return rc;
  if ((rc = mark(3))) return rc;
  float* d_prec = p->d_prec; float* d_snr = p->d_prec + (size_t)F * nti;
  if ((rc = t2_demap_device(ctx, p->d_tib, p->d_dm, F * nti, p->max_cells, p->max_fec, c.mod, c.rotation, true, c.fec_type,
                            c.code_rate, p->d_llr, d_prec, d_snr, nullptr))) return rc;
  if ((rc = mark(4))) return rc;
  // K5 + K6
  const int n_cw = F * c.n_blocks;
  const int k_out = (ldpc_flags & T2B200_LDPC_BCH_DESCRAMBLE) ? p->k_bch : t2b200_ldpc_k(p->code);
  const size_t out_row = (ldpc_flags & T2B200_LDPC_PACK_BITS) ? (size_t)k_out / 8 : (size_t)k_out;
  const bool host_bits = !t2_is_device_ptr(bits_out);
  uint8_t* d_bits = host_bits ? p->d_bits : bits_out;
  const bool host_tr = trials_left && !t2_is_device_ptr(trials_left);
  int32_t* d_tr = trials_left ? (host_tr ? p->d_trials : trials_left) : nullptr;
  if (ctx->opt_bch_correct && (ldpc_flags & T2B200_LDPC_BCH_DESCRAMBLE)) {
    // opt-in N3: LDPC information words (K_ldpc = N_bch bits) -> BCH correction in place -> parity strip + descramble
    if (ldpc_flags & T2B200_LDPC_PACK_BITS) { ctx->err = "t2b200_frames_decode: BCH correction needs byte-per-bit output"; return T2B200_ERR_ARG; }
    if (!p->d_kldpc) T2_CUDA(ctx, cudaMalloc(&p->d_kldpc, (size_t)p->frames_cap * c.n_blocks * 54000));
    if ((rc = t2_ldpc_device(ctx, p->code, p->d_llr, n_cw, p->d_kldpc, d_tr, nullptr, max_trials > 0 ? max_trials : 25,
                             ldpc_flags & ~(unsigned)T2B200_LDPC_BCH_DESCRAMBLE))) return rc;
    if ((rc = t2_bch_device(ctx, p->code, p->d_kldpc, t2b200_ldpc_k(p->code), n_cw, nullptr))) return rc;
    if ((rc = t2_bch_descramble_device(ctx, p->code, p->d_kldpc, n_cw, d_bits))) return rc;
  } else if ((rc = t2_ldpc_device(ctx, p->code, p->d_llr, n_cw, d_bits, d_tr, nullptr, max_trials > 0 ? max_trials : 25, ldpc_flags))) return rc;
  if ((rc = mark(5))) return rc;
  // results
  bool sync = false;
  auto give = [&](void* dst, const void* src, size_t bytes) -> int {
    if (!dst || dst == src) return T2B200_OK;
    const bool dev = t2_is_device_ptr(dst);
    T2_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    sync |= !dev;
    return T2B200_OK;
  };
  if ((rc = give(bits_out, d_bits, out_row * n_cw))) return rc;
  if ((rc = give(trials_left, d_tr, 4 * (size_t)n_cw))) return rc;
  if ((rc = give(sro, d_sro, 4 * (size_t)F * L))) return rc;
  if ((rc = give(phase, d_ph, 4 * (size_t)F * L))) return rc;
  if ((rc = give(snr, d_snr, 4 * (size_t)F * nti))) return rc;
  if ((rc = mark(6))) return rc;
  if (sync) T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // host outputs are filled when the call returns
  return T2B200_OK;
}

extern "C" int t2b200_frames_stage_ms(t2b200_ctx* ctx, float ms_out[6])
{
  if (!ctx || !ms_out) return T2B200_ERR_ARG;
  FramePipe* p = ctx->frames;
  if (!p || !p->timed || !p->ev[6]) { ctx->err = "t2b200_frames_stage_ms: no timed call (T2B200_OPT_STAGE_TIMING)"; return T2B200_ERR_STATE; }
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  T2_CUDA(ctx, cudaEventSynchronize(p->ev[6]));
  for (int i = 0; i < 5; ++i) T2_CUDA(ctx, cudaEventElapsedTime(&ms_out[i], p->ev[i], p->ev[i + 1]));
  T2_CUDA(ctx, cudaEventElapsedTime(&ms_out[5], p->ev[0], p->ev[6]));
  return T2B200_OK;
}
