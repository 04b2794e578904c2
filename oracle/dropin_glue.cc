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
