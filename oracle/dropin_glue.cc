// TEST INFRASTRUCTURE ONLY -- the drop-in proof: the reference's UNMODIFIED dvbt2_demodulator.cpp,
// p1_symbol.cpp, p2_symbol.cpp, dvbt2_definition.cpp, pilot_generator.cpp, address_freq_deinterleaver.cpp and
// bb_de_header.cpp are compiled against the GPU stage classes of sdr_receiver_dvb_t2_b200/host/dropin (same file names,
// class names and signatures as the reference's fast_fourier_transform, data_symbol, fc_symbol, time_deinterleaver,
// llr_demapper, ldpc_decoder, bch_decoder) and linked with libt2b200.so: the reference's own receiver front half (P1
// detection, resampler, loops, L1 parsing) then drives the GPU hot path symbol by symbol and its own bb_de_header sends the
// TS.  This file is what moc would generate for that build (every signal a synchronous call, as in ref_chain.cc) plus a
// small C API for tests/test_dropin_gpu.py.  Built by oracle/Makefile into oracle/_ref/libdropin_chain.so.
#include <cstdint>
#include <cstring>
#include <complex>
#include <iostream>
#include <new>
#include <sstream>
#include <string>
#include <vector>
#include <immintrin.h>

#define private public
#define protected public
#include "dvbt2_demodulator.h"
#undef private
#undef protected

static std::vector<uint8_t> g_bb_bits;
static std::vector<int> g_bb_len;

// ---- dvbt2_demodulator ----
void dvbt2_demodulator::replace_null_indicator(const float, const float) {}
void dvbt2_demodulator::l1_dyn_execute(l1_postsignalling p, int n, complex* in) { deinterleaver->l1_dyn_execute(p, n, in); }
void dvbt2_demodulator::amount_plp(int) {}
void dvbt2_demodulator::data(int n, complex* in) { deinterleaver->execute(n, in); }
void dvbt2_demodulator::stop_deinterleaver() {}
void dvbt2_demodulator::finished() {}
// ---- GUI-only signals of the reference's own stages ----
void p1_symbol::replace_spectrograph(const int, complex*) {}
void p1_symbol::replace_constelation(const int, complex*) {}
void p1_symbol::replace_oscilloscope(const int, complex*) {}
void p1_symbol::bad_signal() {}
void p2_symbol::replace_spectrograph(const int, complex*) {}
void p2_symbol::replace_constelation(const int, complex*) {}
void p2_symbol::replace_oscilloscope(const int, complex*) {}
void p2_symbol::view_l1_presignalling(QString) {}
void p2_symbol::view_l1_postsignalling(QString) {}
void p2_symbol::view_l1_dynamic(QString, bool) {}
void bb_de_header::finished() {}
void bb_de_header::ts_stage(QString) {}
// ---- the stage-to-stage signals of the GPU classes ----
void time_deinterleaver::ti_block(int n, complex* cells, int plp, l1_postsignalling p) { qam->execute(n, cells, plp, p); }
void llr_demapper::soft_multiplexer_de_twist(int* idx, l1_postsignalling p, int n, int8_t* out) { decoder->execute(idx, p, n, out); }
void ldpc_decoder::bit_bch(int* idx, l1_postsignalling p, int n, uint8_t* out) { decoder->execute(idx, p, n, out); }
void bch_decoder::bit_descramble(int plp, l1_postsignalling p, int n, uint8_t* out)
{
  g_bb_bits.insert(g_bb_bits.end(), out, out + n);
  g_bb_len.push_back(n);
  deheader->execute(plp, p, n, out);
}

static dvbt2_demodulator* g_demod = nullptr;
static signal_estimate g_sig;

extern "C" {

int dropin_demod_new(float sample_rate, int need_plp)
{
  void* mem = ::operator new(sizeof(dvbt2_demodulator), std::align_val_t(64));
  std::memset(mem, 0, sizeof(dvbt2_demodulator));
  g_demod = new (mem) dvbt2_demodulator(id_sdrplay, sample_rate);
  g_bb_bits.clear(); g_bb_len.clear();
  OracleTsSink::get().bytes.clear(); OracleTsSink::get().datagram_len.clear();
  bb_de_header* bb = g_demod->deinterleaver->qam->decoder->decoder->deheader;       // main_window.cpp:319-320
  bb->set_out(bb_de_header::out_network, 7654, QString("x"), need_plp);
  return 0;
}
int dropin_demod_feed(int len, int16_t* i_in, int16_t* q_in)
{
  g_sig.frequency_changed = true; g_sig.gain_changed = true;                        // front-end side, rx_sdrplay.cpp:158-197
  g_demod->execute(len, i_in, q_in, &g_sig);
  g_sig.change_frequency = false; g_sig.change_gain = false;
  const int st = (g_demod->crc32_l1_pre ? 1 : 0) | (g_demod->demodulator_init ? 2 : 0) | (g_demod->deint_start ? 4 : 0) |
                 (g_sig.reset ? 8 : 0);
  g_sig.reset = false;
  return st;
}
// The same receiver with the FRONT-END on the GPU as well.  dvbt2_demodulator::execute keeps its per-sample
// loop, resampler and decimator inline (dvbt2_demodulator.cpp:178-221), so they cannot be swapped by a header: this is the edit
// a maintainer makes to execute() -- the chunking of :155-176 and the IQ-imbalance / level estimate of
// :223-253 around ONE t2b200_frontend_execute call per chunk -- written against the demodulator's own members, with
// symbol_acquisition (P1 detection, guard-interval correlation, loop filters, L1 parsing; unmodified) consuming what the GPU
// returns.  The loop is closed through the GPU: the NCO / resampler values of chunk n + 1 come from what the GPU stages made of
// chunk n.
int dropin_demod_feed_gpu_frontend(int len, int16_t* i_in, int16_t* q_in)
{
  dvbt2_demodulator* d = g_demod;
  t2b200_ctx* ctx = t2b200_dropin::context();
  static std::vector<std::complex<float>> out;
  if (out.empty()) {
    out.resize(1 << 17);
    if (t2b200_frontend_configure(ctx, 1, 1 << 17) != T2B200_OK) return -1;
  }
  g_sig.frequency_changed = true; g_sig.gain_changed = true;
  const float two_pi = M_PI_X_2;
  float theta1 = 0.0f, theta2 = 0.0f, theta3 = 0.0f;
  int idx_in = 0;
  while (idx_in < len) {
    if (d->est_chunk == 0) {
      if (d->next_symbol_type == SYMBOL_TYPE_P1) d->est_chunk += P1_LEN;
      d->est_chunk += d->symbol_size;
    }
    double arbitrary_resample = d->resample - d->sample_rate_est_filtered;
    if (arbitrary_resample > d->max_resample) arbitrary_resample = d->max_resample;
    int chunk = static_cast<int>(std::nearbyint(d->est_chunk * arbitrary_resample * d->upsample));
    if (chunk > len - idx_in) chunk = len - idx_in;
    d->chunk = chunk;
    d->phase_nco += d->phase_est_filtered;
    while (d->phase_nco > two_pi) d->phase_nco -= two_pi;
    while (d->phase_nco < -two_pi) d->phase_nco += two_pi;
    t2b200_fe_chunk c;
    c.len_in = chunk; c.short_to_float = d->short_to_float; c.c1 = d->c1; c.c2 = d->c2;
    c.frequency_est_filtered = d->frequency_est_filtered; c.phase_nco = d->phase_nco; c.resample = (float)arbitrary_resample;
    t2b200_fe_result r;
    const long long at = (long long)idx_in * d->convert_input;
    if (t2b200_frontend_execute(ctx, i_in + at, q_in + at, 0, d->convert_input, &c, reinterpret_cast<float*>(out.data()),
                                (long long)out.size(), &r) != T2B200_OK) {
      std::cerr << "t2b200_frontend_execute: " << t2b200_last_error(ctx) << std::endl;
      return -1;
    }
    theta1 += r.theta1; theta2 += r.theta2; theta3 += r.theta3;
    idx_in += chunk;
    d->symbol_acquisition(r.len_out, out.data(), &g_sig);
    if (g_sig.reset) {                                    // dvbt2_demodulator::reset (:111-127) cleared its loop state: so does the GPU's
      t2b200_fe_state st;
      t2b200_frontend_get_state(ctx, 0, &st);
      st.dc_re = st.dc_im = st.frequency_nco = 0.0f;
      t2b200_frontend_set_state(ctx, 0, &st);
    }
  }
  const float a1 = theta1 / len, a2 = theta2 / len, a3 = theta3 / len;          // :223-232
  d->c1 = a1 / a2;
  const float ct = a3 / a2;
  d->c2 = sqrtf(ct * ct - d->c1 * d->c1);
  d->level_detect = a2 * a3;
  if (g_sig.gain_changed) {
    if (d->level_detect < d->level_min) { g_sig.gain_offset = 1; g_sig.change_gain = true; }
    else if (d->level_detect > d->level_max) { g_sig.gain_offset = -1; g_sig.change_gain = true; }
    else { g_sig.gain_offset = 0; g_sig.change_gain = false; }
  }
  g_sig.change_frequency = false; g_sig.change_gain = false;
  const int st = (d->crc32_l1_pre ? 1 : 0) | (d->demodulator_init ? 2 : 0) | (d->deint_start ? 4 : 0) | (g_sig.reset ? 8 : 0);
  g_sig.reset = false;
  return st;
}
long long dropin_launches() { return t2b200_launch_count(t2b200_dropin::context()); }

#define TAP_GETTER(NAME, VEC, TYPE)                                                   \
  long long NAME(TYPE* dst, long long max) {                                          \
    long long n = (long long)(VEC).size();                                            \
    if (dst) std::memcpy(dst, (VEC).data(), sizeof(TYPE) * (size_t)(n < max ? n : max)); \
    return n; }
TAP_GETTER(dropin_tap_ts, OracleTsSink::get().bytes, char)
TAP_GETTER(dropin_tap_ts_datagrams, OracleTsSink::get().datagram_len, int)
TAP_GETTER(dropin_tap_bb_bits, g_bb_bits, uint8_t)
TAP_GETTER(dropin_tap_bb_len, g_bb_len, int)

}  // extern "C"
