// Host-side mirror of the reference's per-stage C++ API (src/DSP/fast_fourier_transform.h and
// src/DVB_T2/{p2_symbol,data_symbol,fc_symbol,time_deinterleaver,llr_demapper,ldpc_decoder,bch_decoder}.h)
// on top of the t2b200 C-ABI.  Same method names, argument meaning, buffer ownership (callee owns ping-pong
// output buffers) and failure behaviour (silent drop + message on stderr) as the reference; Qt-free, so it
// compiles with or without Qt.  INTEGRATION.md shows how the reference's classes forward to these.
//
// Header-only; link with libt2b200.so.  No computation happens on the CPU here: every execute() is a
// t2b200_* call, and construction fails loudly (std::runtime_error) when no GPU context can be had.
#pragma once
#include <complex>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/t2b200.h"

namespace t2b200 {

typedef std::complex<float> complex;

// One GPU context shared by the stages of one receiver (dvbt2_demodulator owns one of each stage).
class context {
 public:
  explicit context(int device = 0) {
    if (t2b200_create(device, &ctx_) != T2B200_OK) throw std::runtime_error("t2b200: no usable CUDA device (no CPU fallback)");
  }
  ~context() { t2b200_destroy(ctx_); }
  context(const context&) = delete;
  context& operator=(const context&) = delete;
  t2b200_ctx* get() const { return ctx_; }
  void check(int rc, const char* what) const {
    if (rc != T2B200_OK) throw std::runtime_error(std::string(what) + ": " + t2b200_last_error(ctx_));
  }
 private:
  t2b200_ctx* ctx_ = nullptr;
};

// DSP/fast_fourier_transform.h:29-70
class fast_fourier_transform {
 public:
  explicit fast_fourier_transform(context& c) : c_(c) {}
  complex* init(int len_in) {                       // returns the input buffer the caller memcpy's into (:54-62)
    n_ = len_in; in_.assign(len_in, complex()); out_.assign(len_in, complex());
    return in_.data();
  }
  complex* execute() {                              // :64-70, halves already swapped
    c_.check(t2b200_fft(c_.get(), n_, reinterpret_cast<const float*>(in_.data()), 1, reinterpret_cast<float*>(out_.data())), "t2b200_fft");
    return out_.data();
  }
 private:
  context& c_; int n_ = 0; std::vector<complex> in_, out_;
};

// The fields of dvbt2_parameters (dvbt2_definition.h:215-250) the symbol stages read.
struct symbol_mode {
  int fft_size, k_total, l_nulls, n_p2, n_data, len_frame, l_fc, c_p2, c_data, n_fc;
};

// DVB_T2/data_symbol.h: init(dvbt2, pilot, address) / execute(idx_symbol, ofdm_cell, sro, phase)
class data_symbol {
 public:
  explicit data_symbol(context& c) : c_(c) {}
  // pilot->data_carrier_map / data_pilot_refer (pilot_generator.h:29,32), address->h_even_data / h_odd_data
  // (address_freq_deinterleaver.h:35-36), amplitudes as data_symbol::init derives them (data_symbol.cpp:46-84)
  void init(const symbol_mode& m, int* const* data_carrier_map, float* const* data_pilot_refer, const int* h_even_data,
            const int* h_odd_data, float amp_sp, float amp_cp) {
    m_ = m;
    const int n = m.len_frame - m.n_p2 - m.l_fc;
    std::vector<int32_t> map((size_t)n * m.k_total); std::vector<float> ref((size_t)n * m.k_total);
    for (int s = 0; s < n; ++s)
      for (int k = 0; k < m.k_total; ++k) { map[(size_t)s * m.k_total + k] = data_carrier_map[s][k]; ref[(size_t)s * m.k_total + k] = data_pilot_refer[s][k]; }
    c_.check(t2b200_eq_configure(c_.get(), T2B200_SYM_DATA, n, m.n_p2, m.fft_size, m.k_total, m.l_nulls, m.c_data, map.data(), ref.data(),
                                 h_even_data, h_odd_data, amp_sp, amp_cp), "t2b200_eq_configure(data)");
    buf_[0].assign(m.c_data, complex()); buf_[1].assign(m.c_data, complex());
  }
  complex* execute(int idx_symbol, complex* ofdm_cell, float& sample_rate_offset, float& phase_offset) {
    swap_ = !swap_;                                  // ping-pong like data_symbol.cpp:140-147
    complex* out = buf_[swap_ ? 1 : 0].data();
    int32_t idx = idx_symbol;
    c_.check(t2b200_equalize(c_.get(), T2B200_SYM_DATA, 1, &idx, reinterpret_cast<const float*>(ofdm_cell), reinterpret_cast<float*>(out),
                             &sample_rate_offset, &phase_offset), "t2b200_equalize(data)");
    return out;
  }
 private:
  context& c_; symbol_mode m_{}; std::vector<complex> buf_[2]; bool swap_ = false;
};

// DVB_T2/fc_symbol.h
class fc_symbol {
 public:
  explicit fc_symbol(context& c) : c_(c) {}
  void init(const symbol_mode& m, const int* fc_carrier_map, const float* fc_pilot_refer, const int* h_even_fc, const int* h_odd_fc,
            float amp_sp) {
    m_ = m;
    c_.check(t2b200_eq_configure(c_.get(), T2B200_SYM_FC, 1, m.len_frame - 1, m.fft_size, m.k_total, m.l_nulls, m.n_fc, fc_carrier_map,
                                 fc_pilot_refer, h_even_fc, h_odd_fc, amp_sp, 0.0f), "t2b200_eq_configure(fc)");
    buf_.assign(m.n_fc, complex());
  }
  complex* execute(complex* ofdm_cell, float& sample_rate_offset, float& phase_offset) {
    int32_t idx = m_.len_frame - 1;                  // fc_symbol.cpp:64
    c_.check(t2b200_equalize(c_.get(), T2B200_SYM_FC, 1, &idx, reinterpret_cast<const float*>(ofdm_cell), reinterpret_cast<float*>(buf_.data()),
                             &sample_rate_offset, &phase_offset), "t2b200_equalize(fc)");
    return buf_.data();
  }
 private:
  context& c_; symbol_mode m_{}; std::vector<complex> buf_;
};

// Equaliser half of DVB_T2/p2_symbol.h (p2_symbol.cpp:94-259).  The L1-pre / L1-post parse (p2_symbol.cpp:301-718)
// stays the reference's host code and reads the cells this returns.
class p2_symbol_equalizer {
 public:
  explicit p2_symbol_equalizer(context& c) : c_(c) {}
  void init(const symbol_mode& m, const int* p2_carrier_map, const float* p2_pilot_refer, const int* h_even_p2, const int* h_odd_p2,
            float amp_p2) {
    m_ = m;
    c_.check(t2b200_eq_configure(c_.get(), T2B200_SYM_P2, 1, 0, m.fft_size, m.k_total, m.l_nulls, m.c_p2, p2_carrier_map, p2_pilot_refer,
                                 h_even_p2, h_odd_p2, amp_p2, 0.0f), "t2b200_eq_configure(p2)");
    buf_.assign(m.c_p2, complex());
  }
  complex* execute(int idx_symbol, complex* ofdm_cell, float& sample_rate_offset, float& phase_offset) {
    int32_t idx = idx_symbol;
    c_.check(t2b200_equalize(c_.get(), T2B200_SYM_P2, 1, &idx, reinterpret_cast<const float*>(ofdm_cell), reinterpret_cast<float*>(buf_.data()),
                             &sample_rate_offset, &phase_offset), "t2b200_equalize(p2)");
    return buf_.data();
  }
 private:
  context& c_; symbol_mode m_{}; std::vector<complex> buf_;
};

// The fields of l1_postsignalling_plp / dynamic_plp (dvbt2_definition.h:272-312) the FEC chain reads.
struct plp_config { int id, plp_cod, plp_mod, plp_rotation, plp_fec_type, plp_num_blocks_max, time_il_length, time_il_type; };

// time_deinterleaver -> llr_demapper -> ldpc_decoder -> bch_decoder as ONE object with the reference's streaming
// interface: start(), l1_dyn_execute() at every P2, execute() per symbol; a callback plays the role of the
// bit_descramble signal (bch_decoder.h:36) and receives one BBFRAME (k_bch bytes, one bit each) at a time.
// Reference batch semantics kept: FECFRAMEs are decoded 32 at a time in lock-step, a batch that does not converge
// within 25 trials is dropped with the reference's message (ldpc_decoder.cpp:264-268), a trailing partial batch
// waits for more frames (llr_demapper.cpp:749-765).  Single PLP, type 1, P_I = 1 (the cases the reference handles,
// SURVEY 7.3-9).
class fec_chain {
 public:
  typedef std::function<void(int plp_id, int len, uint8_t* bits)> bbframe_sink;
  fec_chain(context& c, bbframe_sink sink) : c_(c), sink_(std::move(sink)) {}

  void start(const plp_config& plp, int l1_post_size) {                       // time_deinterleaver.cpp:38-145
    plp_ = plp;
    p2_start_ = 1840 + l1_post_size;
    nbits_ = plp.plp_fec_type ? 64800 : 16200;
    cpf_ = nbits_ / (2 * (plp.plp_mod + 1));
    n_ti_ = plp.time_il_type == 0 ? plp.time_il_length : 1;
    code_ = t2b200_ldpc_code_id(plp.plp_fec_type, plp.plp_cod);
    k_bch_ = t2b200_ldpc_k_bch(code_);
    c_.check(t2b200_ti_configure(c_.get(), plp.id, plp.plp_fec_type, plp.plp_mod, plp.plp_num_blocks_max, nullptr), "t2b200_ti_configure");
    llr_.clear(); started_ = true;
  }
  // time_deinterleaver.cpp:268-286: this frame's FEC-block count, then the P2 cells
  void l1_dyn_execute(int num_blocks, int len_in, complex* ofdm_cell) {
    blocks_.clear();
    const int base = num_blocks / n_ti_;
    for (int j = 0; j < n_ti_; ++j) blocks_.push_back(base + (j >= n_ti_ - num_blocks % n_ti_ ? 1 : 0));
    ti_idx_ = 0; cells_.clear();
    if (len_in > p2_start_) execute(len_in - p2_start_, ofdm_cell + p2_start_);
  }
  void execute(int len_in, complex* ofdm_cell) {                             // time_deinterleaver.cpp:288-376
    if (!started_) return;
    cells_.insert(cells_.end(), ofdm_cell, ofdm_cell + len_in);
    while (ti_idx_ < (int)blocks_.size() && (int)cells_.size() >= blocks_[ti_idx_] * cpf_) {
      const int n = blocks_[ti_idx_] * cpf_;
      ti_block(blocks_[ti_idx_], cells_.data());
      cells_.erase(cells_.begin(), cells_.begin() + n);
      ++ti_idx_;
    }
    if (ti_idx_ >= (int)blocks_.size()) cells_.clear();                      // dummy cells after the PLP
  }
  float last_snr() const { return snr_; }

 private:
  void ti_block(int n_fec, const complex* arrival) {
    std::vector<complex> ti((size_t)n_fec * cpf_);
    int32_t nf = n_fec;
    c_.check(t2b200_ti_deinterleave(c_.get(), plp_.id, reinterpret_cast<const float*>(arrival), 1, &nf, reinterpret_cast<float*>(ti.data())),
             "t2b200_ti_deinterleave");
    const size_t at = llr_.size();
    llr_.resize(at + (size_t)n_fec * nbits_);
    c_.check(t2b200_demap(c_.get(), reinterpret_cast<float*>(ti.data()), 1, &nf, plp_.plp_mod, plp_.plp_rotation, plp_.plp_fec_type,
                          plp_.plp_cod, llr_.data() + at, &snr_, nullptr, nullptr), "t2b200_demap");
    const size_t batch = (size_t)32 * nbits_;
    while (llr_.size() >= batch) {                                            // ldpc_decoder.cpp:157-301 + bch_decoder.cpp:63-164
      std::vector<uint8_t> bits((size_t)32 * k_bch_);
      int32_t trials[32];
      c_.check(t2b200_ldpc_decode(c_.get(), code_, llr_.data(), 32, bits.data(), trials, nullptr, nullptr, 25,
                                  T2B200_LDPC_GROUP32 | T2B200_LDPC_BCH_DESCRAMBLE), "t2b200_ldpc_decode");
      if (trials[0] < 0) std::fprintf(stderr, "LDPC decoder could not recover the codeword! %d\n", trials[0]);
      else for (int f = 0; f < 32; ++f) sink_(plp_.id, k_bch_, bits.data() + (size_t)f * k_bch_);
      llr_.erase(llr_.begin(), llr_.begin() + batch);
    }
  }
  context& c_; bbframe_sink sink_;
  plp_config plp_{}; bool started_ = false;
  int p2_start_ = 0, nbits_ = 0, cpf_ = 0, n_ti_ = 1, code_ = 0, k_bch_ = 0, ti_idx_ = 0;
  float snr_ = 0.f;
  std::vector<int> blocks_; std::vector<complex> cells_; std::vector<int8_t> llr_;
};

// bb_de_header.h:41-108 -- execute(plp_id, l1_post, len_in, bits): one BBFRAME in, one datagram out (the reference sends it
// with socket->writeDatagram, bb_de_header.cpp:431-441; here a sink callback receives it).  High-efficiency mode (status 0)
// and normal mode (status 3) are both built on the GPU.  need_plp as in set_out (bb_de_header.cpp:502-529).
class bb_de_header {
 public:
  typedef std::function<void(const uint8_t* datagram, int len)> ts_sink;
  bb_de_header(context& c, ts_sink sink) : c_(c), sink_(std::move(sink)) {}
  void set_out(int need_plp) { need_plp_ = need_plp; c_.check(t2b200_ts_reset(c_.get(), need_plp), "t2b200_ts_reset"); }
  void execute(int plp_id, int len_in, uint8_t* bits) {
    if (plp_id != need_plp_) return;                                          // bb_de_header.cpp:133-136
    out_.resize((size_t)len_in / 8 + 2 * 188 + 16);
    int32_t dl = 0, st = 0; long long total = 0;
    c_.check(t2b200_ts_packetize(c_.get(), plp_id, bits, 1, len_in, out_.data(), out_.size(), &dl, &st, &total), "t2b200_ts_packetize");
    if (st == 1) { std::fprintf(stderr, "Baseband header CRC8 error.\n"); return; }   // bb_de_header.cpp:109-112
    if (st == 0 || st == 3) sink_(out_.data(), dl);                                      // (a zero-length datagram is sent too, like the reference)
  }
 private:
  context& c_; ts_sink sink_; int need_plp_ = 0; std::vector<uint8_t> out_;
};

// The front-end of dvbt2_demodulator::execute (dvbt2_demodulator.cpp:178-221): the per-sample loop (int16 -> float, DC
// removal, IQ-imbalance statistics and correction, NCO), interpolator_farrow::operator() (DSP/interpolator_farrow.hh:41-68) and
// filter_decimator::execute (DSP/filter_decimator.h:72-131) of one chunk in one call; the arguments carry the member names
// execute() uses.  The carried state (DC averages, frequency_nco, resampler phase, delay lines) lives in the engine.
class frontend {
 public:
  explicit frontend(context& c, int max_chunk = 1 << 17) : c_(c) {
    c_.check(t2b200_frontend_configure(c_.get(), 1, max_chunk), "t2b200_frontend_configure");
  }
  void reset() { c_.check(t2b200_frontend_reset(c_.get(), 0), "t2b200_frontend_reset"); }      // dvbt2_demodulator::reset
  // returns len_out_decimator; theta1..3 are accumulated on, as est_1_bit_quantization does (:256-265)
  int execute(int chunk, const int16_t* i_in, const int16_t* q_in, int convert_input, float short_to_float, float c1, float c2,
              float frequency_est_filtered, float phase_nco, double arbitrary_resample, complex* out_decimator, int capacity,
              float& theta1, float& theta2, float& theta3) {
    t2b200_fe_chunk ck = {chunk, short_to_float, c1, c2, frequency_est_filtered, phase_nco, static_cast<float>(arbitrary_resample)};
    t2b200_fe_result r;
    c_.check(t2b200_frontend_execute(c_.get(), i_in, q_in, 0, convert_input, &ck, reinterpret_cast<float*>(out_decimator), capacity, &r),
             "t2b200_frontend_execute");
    theta1 += r.theta1; theta2 += r.theta2; theta3 += r.theta3;
    return r.len_out;
  }
  // guard-interval correlation of one buffered symbol (dvbt2_demodulator.cpp:321-330) -> frequency_est
  float cp_correlate(const complex* buffer_sym, int fft_size, int guard_interval_size) {
    float est = 0.f;
    c_.check(t2b200_cp_correlate(c_.get(), reinterpret_cast<const float*>(buffer_sym), 1, fft_size + guard_interval_size, fft_size,
                                 guard_interval_size, &est), "t2b200_cp_correlate");
    return est;
  }
  // p1_symbol's sliding correlator over a block (p1_symbol.cpp:141-165): correlation[n], out[n]; history = the 2046 samples
  // in front of the block or nullptr, fq_index = the block's position in the 1024-step frequency shift
  void p1_correlate(const complex* in, int n, const complex* history, int fq_index, float* correlation, complex* out) {
    c_.check(t2b200_p1_correlate(c_.get(), reinterpret_cast<const float*>(in), n, reinterpret_cast<const float*>(history), fq_index,
                                 correlation, reinterpret_cast<float*>(out)), "t2b200_p1_correlate");
  }
 private:
  context& c_;
};

}  // namespace t2b200
