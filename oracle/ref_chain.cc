// TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference stages (compiled from /root/reference
// by oracle/Makefile, Qt replaced by oracle/qtshim) one at a time so that every CUDA stage can be
// compared with the reference's own output on the same input.  Nothing here is product code and
// nothing under sdr_receiver_dvb_t2_b200/ links it.
//
// Part 1 is the moc replacement: every Qt signal of src/DVB_T2 gets a body.  Signals that feed the
// next stage call that stage synchronously (dvbt2_demodulator.cpp:86-88, time_deinterleaver.cpp:27-35,
// llr_demapper.cpp:81-89, ldpc_decoder.cpp:126-134, bch_decoder.cpp:31-39), after copying the payload
// into a tap buffer when a tap is armed; GUI signals do nothing.
// Part 2 is a C API (ref_*) used through ctypes by oracle/pyoracle.py.
#include <cstdint>
#include <cstring>
#include <complex>
#include <vector>
#include <new>
#include <dlfcn.h>
#include <execinfo.h>
#include <csignal>
#include <unistd.h>

#define private public          // the harness pokes at stage objects; the sources stay untouched
#define protected public
#include "dvbt2_demodulator.h"
#undef private
#undef protected

// ------------------------------------------------------------------------------------------
// taps
struct Taps {
  bool chain_after_ti = true, chain_after_demap = true, chain_after_ldpc = true, chain_after_bch = true;
  std::vector<std::complex<float>> ti_cells;  std::vector<int> ti_sizes, ti_plp;
  std::vector<int8_t> llr;                    std::vector<int> llr_plp;
  std::vector<uint8_t> ldpc_bits;
  std::vector<uint8_t> bb_bits;               std::vector<int> bb_len, bb_plp;
  std::vector<float> snr;
  void clear() { ti_cells.clear(); ti_sizes.clear(); ti_plp.clear(); llr.clear(); llr_plp.clear(); ldpc_bits.clear();
                 bb_bits.clear(); bb_len.clear(); bb_plp.clear(); snr.clear(); }
};
static Taps g_taps;

// ---- signals of dvbt2_demodulator ----
void dvbt2_demodulator::replace_null_indicator(const float, const float) {}
void dvbt2_demodulator::l1_dyn_execute(l1_postsignalling p, int n, complex* in) { deinterleaver->l1_dyn_execute(p, n, in); }
void dvbt2_demodulator::amount_plp(int) {}
void dvbt2_demodulator::data(int n, complex* in) { deinterleaver->execute(n, in); }
void dvbt2_demodulator::stop_deinterleaver() {}
void dvbt2_demodulator::finished() {}
// ---- GUI-only signals ----
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
void data_symbol::replace_spectrograph(const int, complex*) {}
void data_symbol::replace_constelation(const int, complex*) {}
void data_symbol::replace_oscilloscope(const int, complex*) {}
void fc_symbol::replace_spectrograph(const int, complex*) {}
void fc_symbol::replace_constelation(const int, complex*) {}
void fc_symbol::replace_oscilloscope(const int, complex*) {}
// ---- time_deinterleaver ----
void time_deinterleaver::ti_block(int n, complex* cells, int plp, l1_postsignalling p)
{
  g_taps.ti_cells.insert(g_taps.ti_cells.end(), cells, cells + n);     // before the demapper derotates in place
  g_taps.ti_sizes.push_back(n); g_taps.ti_plp.push_back(plp);
  if (g_taps.chain_after_ti) qam->execute(n, cells, plp, p);
}
void time_deinterleaver::replace_constelation(const int, complex*) {}
void time_deinterleaver::stop_qam() {}
void time_deinterleaver::finished() {}
// ---- llr_demapper ----
void llr_demapper::signal_noise_ratio(float s) { g_taps.snr.push_back(s); }
void llr_demapper::soft_multiplexer_de_twist(int* idx, l1_postsignalling p, int n, int8_t* out)
{
  g_taps.llr.insert(g_taps.llr.end(), out, out + n);
  for (int i = 0; i < 32; ++i) g_taps.llr_plp.push_back(idx[i]);
  if (g_taps.chain_after_demap) decoder->execute(idx, p, n, out);
}
void llr_demapper::stop_decoder() {}
void llr_demapper::finished() {}
// ---- ldpc_decoder ----
void ldpc_decoder::bit_bch(int* idx, l1_postsignalling p, int n, uint8_t* out)
{
  g_taps.ldpc_bits.insert(g_taps.ldpc_bits.end(), out, out + n);
  if (g_taps.chain_after_ldpc) decoder->execute(idx, p, n, out);
}
void ldpc_decoder::check(int, uint8_t*) {}
void ldpc_decoder::stop_decoder() {}
void ldpc_decoder::finished() {}
// ---- bch_decoder ----
void bch_decoder::bit_descramble(int plp, l1_postsignalling p, int n, uint8_t* out)
{
  g_taps.bb_bits.insert(g_taps.bb_bits.end(), out, out + n);
  g_taps.bb_len.push_back(n); g_taps.bb_plp.push_back(plp);
  if (g_taps.chain_after_bch) deheader->execute(plp, p, n, out);
}
void bch_decoder::check(int, uint8_t*) {}
void bch_decoder::stop_deheader() {}
void bch_decoder::finished() {}
// ---- bb_de_header ----
void bb_de_header::finished() {}
void bb_de_header::ts_stage(QString) {}

// ------------------------------------------------------------------------------------------
// Front-end taps: oracle/tap/DSP/*.h wrap the reference's resampler and decimator and report every call
// here.  At the resampler call of a chunk (dvbt2_demodulator.cpp:216-218) the demodulator's members still hold what
// the per-sample loop of that chunk used (:170-213), so one record per chunk is: the chunk length, the loop parameters,
// the loop state after the chunk, the derotated samples (the resampler's input), its output and the decimator's output.
struct FeTap {
  int first = 0, count = 0;                       // samples are kept for chunks first .. first + count - 1
  long long n_interp = 0;                         // resampler outputs so far: its parity is the decimator's phase
  std::vector<double> info;                       // FE_REC doubles per chunk, all chunks, see oracle_tap_farrow
  std::vector<std::complex<float>> derot, interp, decim;
};
enum { FE_REC = 16 };
static FeTap g_fe_tap;
static dvbt2_demodulator* g_demod = nullptr;
extern "C" void oracle_tap_farrow(int len_in, const void* in, double resample, int len_out, const void* out)
{
  if (g_fe_tap.count <= 0 || !g_demod) return;
  const dvbt2_demodulator* d = g_demod;
  const int idx = (int)(g_fe_tap.info.size() / FE_REC);
  g_fe_tap.n_interp += len_out;
  const double rec[FE_REC] = {(double)len_in, (double)len_out, resample, d->phase_nco, d->frequency_est_filtered, d->c1, d->c2,
                              d->frequency_nco, d->exp_avg_dc_real.out, d->exp_avg_dc_imag.out, 0.0, (double)d->short_to_float,
                              (double)d->interpolator.impl.x1, (double)g_fe_tap.n_interp, 0.0, 0.0};
  g_fe_tap.info.insert(g_fe_tap.info.end(), rec, rec + FE_REC);
  if (idx < g_fe_tap.first || idx >= g_fe_tap.first + g_fe_tap.count) return;
  const std::complex<float>* a = static_cast<const std::complex<float>*>(in);
  const std::complex<float>* b = static_cast<const std::complex<float>*>(out);
  g_fe_tap.derot.insert(g_fe_tap.derot.end(), a, a + len_in);
  g_fe_tap.interp.insert(g_fe_tap.interp.end(), b, b + len_out);
}
extern "C" void oracle_tap_decimator(int len_in, const void*, int len_out, const void* out)
{
  const int n = (int)(g_fe_tap.info.size() / FE_REC);
  if (n == 0 || g_fe_tap.info[FE_REC * (n - 1) + 10] != 0.0 || g_fe_tap.info[FE_REC * (n - 1) + 1] != (double)len_in) return;
  g_fe_tap.info[FE_REC * (n - 1) + 10] = (double)len_out + 0.5;   // + 0.5: marks the record complete even for len_out == 0
  if (n - 1 < g_fe_tap.first || n - 1 >= g_fe_tap.first + g_fe_tap.count) return;
  const std::complex<float>* b = static_cast<const std::complex<float>*>(out);
  g_fe_tap.decim.insert(g_fe_tap.decim.end(), b, b + len_out);
}

// ------------------------------------------------------------------------------------------
// Part 2: stage-level C API
namespace {

template <class T, class... A> T* make_zeroed(A&&... a)
{
  // some members are read before they are written (time_deinterleaver.h:95-96);
  // zeroed storage makes that deterministic.
  void* mem = ::operator new(sizeof(T), std::align_val_t(64));
  std::memset(mem, 0, sizeof(T));
  return new (mem) T(std::forward<A>(a)...);
}

struct Rx {
  dvbt2_parameters dvbt2;
  fast_fourier_transform* fft = nullptr; complex* in_fft = nullptr;
  pilot_generator* pilot = nullptr;
  address_freq_deinterleaver* fq = nullptr;
  p2_symbol* p2 = nullptr;
  data_symbol* data = nullptr;
  fc_symbol* fc = nullptr;
  // FEC side (one instance per process: the stages keep static locals)
  QMutex mtx; QWaitCondition cond;
  time_deinterleaver* ti = nullptr;
  l1_presignalling l1_pre;
  l1_postsignalling l1_post;
  std::vector<l1_postsignalling_plp> plps;
  std::vector<dynamic_plp> dyn;
};
Rx* g_rx = nullptr;

}  // namespace

extern "C" {

// Build the demodulator-side objects the way dvbt2_demodulator::init_dvbt2 (dvbt2_demodulator.cpp:129-143)
// and the post-L1-pre branch (dvbt2_demodulator.cpp:393-404) do.  SISO only, like the reference.
// out_params: fft_size,k_total,l_nulls,c_p2,c_data,n_fc,c_fc,n_data,len_frame,l_fc,n_p2,guard_interval_size,k_ext
int ref_rx_init(int fft_mode, int carrier_mode, int pilot_pattern, int guard_interval_mode, int n_data, int papr_mode,
                int* out_params)
{
  Rx* r = new Rx();
  std::memset(&r->dvbt2, 0, sizeof(r->dvbt2));
  dvbt2_parameters& d = r->dvbt2;
  d.preamble = T2_SISO; d.fft_mode = fft_mode; d.bandwidth = BANDWIDTH_8_0_MHZ; d.miso_group = MISO_TX1;
  dvbt2_p2_parameters_init(d);
  r->fft = new fast_fourier_transform; r->in_fft = r->fft->init(d.fft_size);
  r->pilot = new pilot_generator(); r->fq = new address_freq_deinterleaver();
  r->fq->init(d);
  r->p2 = make_zeroed<p2_symbol>();
  r->p2->init(d, r->pilot, r->fq);
  // what l1_pre_info writes back (p2_symbol.cpp:493-499)
  if (d.carrier_mode != carrier_mode) {
    // p2_symbol::init always falls back to extended carriers (it calls dvbt2_p2_parameters_init, dvbt2_definition.cpp:89),
    // so the reference's P2 tables exist for the extended mode only; the data / frame-closing side follows L1-pre.
    d.carrier_mode = carrier_mode;
    dvbt2_bwt_ext_parameters_init(d);
  }
  d.guard_interval_mode = guard_interval_mode; d.papr_mode = papr_mode; d.pilot_pattern = pilot_pattern; d.n_data = n_data;
  r->data = make_zeroed<data_symbol>();
  r->data->init(d, r->pilot, r->fq);            // -> dvbt2_data_parameters_init, pilot->data_generator, fq tables
  if (d.l_fc) { r->fc = make_zeroed<fc_symbol>(); r->fc->init(d, r->pilot, r->fq); }
  int v[13] = {d.fft_size, d.k_total, d.l_nulls, d.c_p2, d.c_data, d.n_fc, d.c_fc, d.n_data, d.len_frame, d.l_fc, d.n_p2,
               d.guard_interval_size, d.k_ext};
  std::memcpy(out_params, v, sizeof(v));
  g_rx = r;
  return 0;
}

// tables the drop-in facade hands to the GPU engine (pilot_generator.h:28-33, address_freq_deinterleaver.h:33-38)
void ref_rx_tables_data(int idx_data_symbol, int* carrier_map, float* pilot_refer)
{
  std::memcpy(carrier_map, g_rx->pilot->data_carrier_map[idx_data_symbol], sizeof(int) * g_rx->dvbt2.k_total);
  std::memcpy(pilot_refer, g_rx->pilot->data_pilot_refer[idx_data_symbol], sizeof(float) * g_rx->dvbt2.k_total);
}
void ref_rx_tables_p2(int* carrier_map, float* pilot_refer)
{
  std::memcpy(carrier_map, g_rx->pilot->p2_carrier_map, sizeof(int) * g_rx->dvbt2.k_total);
  std::memcpy(pilot_refer, g_rx->pilot->p2_pilot_refer[0], sizeof(float) * g_rx->dvbt2.k_total);
}
int ref_rx_tables_fc(int* carrier_map, float* pilot_refer)
{
  if (!g_rx->dvbt2.l_fc) return 0;
  std::memcpy(carrier_map, g_rx->pilot->fc_carrier_map, sizeof(int) * g_rx->dvbt2.k_total);
  std::memcpy(pilot_refer, g_rx->pilot->fc_pilot_refer, sizeof(float) * g_rx->dvbt2.k_total);
  return 1;
}
// kind: 0 p2, 1 data, 2 fc; each array has 32768 ints
void ref_rx_tables_h(int kind, int* h_even, int* h_odd)
{
  const int* e = kind == 0 ? g_rx->fq->h_even_p2 : kind == 1 ? g_rx->fq->h_even_data : g_rx->fq->h_even_fc;
  const int* o = kind == 0 ? g_rx->fq->h_odd_p2 : kind == 1 ? g_rx->fq->h_odd_data : g_rx->fq->h_odd_fc;
  std::memcpy(h_even, e, sizeof(int) * 32768); std::memcpy(h_odd, o, sizeof(int) * 32768);
}
void ref_rx_amps(float* amp_p2, float* amp_sp, float* amp_cp)
{
  *amp_p2 = g_rx->p2->amp_p2; *amp_sp = g_rx->data->amp_sp; *amp_cp = g_rx->data->amp_cp;
}

// fast_fourier_transform::execute (DSP/fast_fourier_transform.h:64-70) via the bundled FFTW
void ref_fft(const float* in, float* out)
{
  const int n = g_rx->dvbt2.fft_size;
  std::memcpy(g_rx->in_fft, in, sizeof(complex) * n);
  complex* o = g_rx->fft->execute();
  std::memcpy(out, o, sizeof(complex) * n);
}

// data_symbol::execute (data_symbol.cpp:108-335); freq holds fft_size shifted bins
void ref_data_symbol(int idx_symbol, const float* freq, float* cells, float* sro, float* phase)
{
  std::vector<complex> tmp(g_rx->dvbt2.fft_size);
  std::memcpy(tmp.data(), freq, sizeof(complex) * tmp.size());
  complex* o = g_rx->data->execute(idx_symbol, tmp.data(), *sro, *phase);
  std::memcpy(cells, o, sizeof(complex) * g_rx->dvbt2.c_data);
}
void ref_fc_symbol(const float* freq, float* cells, float* sro, float* phase)
{
  std::vector<complex> tmp(g_rx->dvbt2.fft_size);
  std::memcpy(tmp.data(), freq, sizeof(complex) * tmp.size());
  complex* o = g_rx->fc->execute(tmp.data(), *sro, *phase);
  std::memcpy(cells, o, sizeof(complex) * g_rx->dvbt2.n_fc);
}
// p2_symbol::execute (p2_symbol.cpp:89-299); returns crc flags: bit0 L1-pre ok, bit1 L1-post ok
int ref_p2_symbol(const float* freq, float* cells, float* sro, float* phase)
{
  std::vector<complex> tmp(g_rx->dvbt2.fft_size);
  std::memcpy(tmp.data(), freq, sizeof(complex) * tmp.size());
  int idx = 0; bool c1 = false, c2 = false;
  l1_presignalling pre; l1_postsignalling post;
  dvbt2_parameters d = g_rx->dvbt2;                 // l1_pre_info would overwrite mode fields from garbage L1
  complex* o = g_rx->p2->execute(d, true, idx, tmp.data(), pre, post, c1, c2, *sro, *phase);
  std::memcpy(cells, o, sizeof(complex) * g_rx->dvbt2.c_p2);
  return (c1 ? 1 : 0) | (c2 ? 2 : 0);
}

// ---- FEC side -------------------------------------------------------------------------------
// One PLP set, type-1 contiguous, described by plain ints (what l1_post_info would have parsed):
// per plp: id, plp_cod, plp_mod, plp_rotation, plp_fec_type, plp_num_blocks_max, time_il_length, time_il_type
int ref_fec_start(int num_plp, const int* plp_desc, int l1_post_size, int need_plp)
{
  Rx* r = g_rx;
  r->plps.assign(num_plp, l1_postsignalling_plp());
  r->dyn.assign(num_plp, dynamic_plp());
  for (int i = 0; i < num_plp; ++i) {
    const int* p = plp_desc + 8 * i;
    l1_postsignalling_plp& q = r->plps[i];
    q.id = p[0]; q.plp_cod = p[1]; q.plp_mod = p[2]; q.plp_rotation = p[3]; q.plp_fec_type = p[4];
    q.plp_num_blocks_max = p[5]; q.time_il_length = p[6]; q.time_il_type = p[7];
    q.frame_interval = 1; q.first_frame_idx = 0; q.plp_type = 1;
  }
  r->l1_post.num_plp = num_plp; r->l1_post.plp = r->plps.data(); r->l1_post.dyn.plp = r->dyn.data();
  r->l1_pre.l1_post_size = l1_post_size;
  r->ti = make_zeroed<time_deinterleaver>(&r->cond, &r->mtx);
  r->ti->start(r->dvbt2, r->l1_pre, r->l1_post);
  bb_de_header* bb = r->ti->qam->decoder->decoder->deheader;
  bb->set_out(bb_de_header::out_network, 7654, QString("x"), need_plp);
  g_taps.clear(); OracleTsSink::get().bytes.clear(); OracleTsSink::get().datagram_len.clear();
  return 0;
}
void ref_fec_chain(int after_ti, int after_demap, int after_ldpc, int after_bch)
{
  g_taps.chain_after_ti = after_ti; g_taps.chain_after_demap = after_demap;
  g_taps.chain_after_ldpc = after_ldpc; g_taps.chain_after_bch = after_bch;
}
// first call of a T2 frame: the c_p2 deinterleaved P2 cells + this frame's dynamic L1 (start, num_blocks per plp)
void ref_fec_feed_p2(const int* dyn_start, const int* dyn_num_blocks, int n_cells, float* cells)
{
  Rx* r = g_rx;
  for (int i = 0; i < r->l1_post.num_plp; ++i) {
    r->dyn[i].id = r->plps[i].id; r->dyn[i].start = dyn_start[i]; r->dyn[i].num_blocks = dyn_num_blocks[i];
  }
  r->ti->l1_dyn_execute(r->l1_post, n_cells, reinterpret_cast<complex*>(cells));
}
void ref_fec_feed(int n_cells, float* cells) { g_rx->ti->execute(n_cells, reinterpret_cast<complex*>(cells)); }

// ldpc_decoder::execute (ldpc_decoder.cpp:157-301) on one batch of 32 FECFRAMEs of plp, then whatever is chained behind it
// (bch_decoder::execute, bb_de_header::execute): the stock FEC back half on caller-supplied LLRs
void ref_ldpc_batch(int plp, int8_t* llr32, int fec_size)
{
  int idx[32];
  for (int i = 0; i < 32; ++i) idx[i] = plp;
  g_rx->ti->qam->decoder->execute(idx, g_rx->l1_post, 32 * fec_size, llr32);
}

// stand-alone stage entry points (stage objects of the same instance)
void ref_demap(int n_cells, float* cells, int plp) { g_rx->ti->qam->execute(n_cells, reinterpret_cast<complex*>(cells), plp, g_rx->l1_post); }

// bb_de_header::execute alone (bb_de_header.cpp:84-445) on one BBFRAME (one byte per bit): its datagram lands in the TS sink
void ref_bb_deheader(int plp, int len, uint8_t* bits)
{
  g_rx->ti->qam->decoder->decoder->deheader->execute(plp, g_rx->l1_post, len, bits);
}

// cell-deinterleaver permutation the reference built for plp (time_deinterleaver.cpp:174-266)
int ref_ti_permutation(int plp, int* out, int max)
{
  int n = g_rx->plps[plp].plp_num_blocks_max * g_rx->ti->cells_per_fec_block[plp];
  if (n > max) return -n;
  std::memcpy(out, g_rx->ti->permutations[plp], sizeof(int) * n);
  return n;
}

// the demapper's address table for (fec, mod, rate) as selected in llr_demapper.cpp:294-302,455-463,677-686
int ref_demap_address(int fec_normal, int mod, int code_rate, int* out)
{
  llr_demapper* q = g_rx->ti->qam;
  const int* a = nullptr;
  if (mod == 1) a = fec_normal ? (code_rate == C3_5 ? q->address_qam16_fecnormal_3_5 : q->address_qam16_fecnormal) : q->address_qam16_fecshort;
  if (mod == 2) a = fec_normal ? (code_rate == C3_5 ? q->address_qam64_fecnormal_3_5 : q->address_qam64_fecnormal) : q->address_qam64_fecshort;
  if (mod == 3) a = fec_normal ? (code_rate == C3_5 ? q->address_qam256_fecnormal_3_5 : code_rate == C2_3 ? q->address_qam256_fecnormal_2_3
                                                                                      : q->address_qam256_fecnormal) : q->address_qam256_fecshort;
  if (!a) return 0;
  std::memcpy(out, a, sizeof(int) * (fec_normal ? 64800 : 16200));
  return 1;
}

// p1_symbol's sliding correlator (p1_symbol.cpp:75-178) sample by sample: the thresholds are pushed out of reach so that the
// detection branch never fires, and the `correlation` member is read after every sample
int ref_p1_trace(const float* in, int n, float* correlation)
{
  static p1_symbol* p1 = nullptr;
  if (!p1) p1 = make_zeroed<p1_symbol>();
  p1->reset_buffer();
  std::vector<complex> buf(4 * P1_LEN);
  dvbt2_parameters d; std::memset(&d, 0, sizeof(d));
  for (int i = 0; i < n; ++i) {
    complex x(in[2 * i], in[2 * i + 1]);
    int consume = 0, idx_sym = 0; double cfo = 0; bool dec = false, rst = false;
    p1->execute(true, 3.0e+30f, 1, &x, consume, buf.data(), idx_sym, d, cfo, dec, rst);
    correlation[i] = p1->correlation;
  }
  return 0;
}

// tap read-out: returns element count; copies up to max elements when dst != null
#define TAP_GETTER(NAME, VEC, TYPE)                                                   \
  long long NAME(TYPE* dst, long long max) {                                          \
    long long n = (long long)(VEC).size();                                            \
    if (dst) std::memcpy(dst, (VEC).data(), sizeof(TYPE) * (size_t)(n < max ? n : max)); \
    return n; }
TAP_GETTER(ref_tap_ti_cells, g_taps.ti_cells, std::complex<float>)
TAP_GETTER(ref_tap_ti_sizes, g_taps.ti_sizes, int)
TAP_GETTER(ref_tap_llr, g_taps.llr, int8_t)
TAP_GETTER(ref_tap_ldpc_bits, g_taps.ldpc_bits, uint8_t)
TAP_GETTER(ref_tap_bb_bits, g_taps.bb_bits, uint8_t)
TAP_GETTER(ref_tap_bb_len, g_taps.bb_len, int)
TAP_GETTER(ref_tap_snr, g_taps.snr, float)
TAP_GETTER(ref_tap_ts, OracleTsSink::get().bytes, char)
TAP_GETTER(ref_tap_ts_datagrams, OracleTsSink::get().datagram_len, int)
void ref_tap_frontend_arm(int first, int count) { g_fe_tap = FeTap(); g_fe_tap.first = first; g_fe_tap.count = count; }
TAP_GETTER(ref_tap_fe_info, g_fe_tap.info, double)
TAP_GETTER(ref_tap_fe_derot, g_fe_tap.derot, std::complex<float>)
TAP_GETTER(ref_tap_fe_interp, g_fe_tap.interp, std::complex<float>)
TAP_GETTER(ref_tap_fe_decim, g_fe_tap.decim, std::complex<float>)
void ref_tap_clear() { g_taps.clear(); OracleTsSink::get().bytes.clear(); OracleTsSink::get().datagram_len.clear(); }

// full receiver: dvbt2_demodulator::execute on int16 I/Q chunks (dvbt2_demodulator.cpp:145-254)
static signal_estimate g_sig;

// Tap on the FFT input of every OFDM symbol (the `in_fft` window of dvbt2_demodulator.cpp:332): the reference calls
// fftwf_execute from an inline header function, so the call is interposed here (this library precedes libfftw3f in its
// own lookup scope) and forwarded with dlsym(RTLD_NEXT).  One record per symbol: kind (SYMBOL_TYPE_*), symbol index,
// whether the FEC chain was already running, then fft_size samples.
struct FftTap { bool armed = false; std::vector<std::complex<float>> in; std::vector<int> info; };
static FftTap g_fft_tap;
void fftwf_execute(const fftwf_plan p)
{
  static void (*real)(const fftwf_plan) = reinterpret_cast<void (*)(const fftwf_plan)>(dlsym(RTLD_NEXT, "fftwf_execute"));
  if (g_fft_tap.armed && g_demod && g_demod->in_fft && p == g_demod->fft->plan) {
    const int n = g_demod->dvbt2.fft_size;
    g_fft_tap.in.insert(g_fft_tap.in.end(), g_demod->in_fft, g_demod->in_fft + n);
    g_fft_tap.info.push_back(g_demod->next_symbol_type);
    g_fft_tap.info.push_back(g_demod->next_symbol_type == SYMBOL_TYPE_P2 ? 0 : g_demod->idx_symbol);
    g_fft_tap.info.push_back((g_demod->deint_start ? 1 : 0) | (g_demod->demodulator_init ? 2 : 0));
  }
  real(p);
}
void ref_tap_fft_arm(int on) { g_fft_tap.armed = on != 0; if (!on) { g_fft_tap.in.clear(); g_fft_tap.info.clear(); } }
TAP_GETTER(ref_tap_fft_in, g_fft_tap.in, std::complex<float>)
TAP_GETTER(ref_tap_fft_info, g_fft_tap.info, int)
void ref_tap_fft_clear() { g_fft_tap.in.clear(); g_fft_tap.info.clear(); }
// mode parameters the demodulator derived from P1 + L1-pre (dvbt2_parameters) and the L1 of the last P2
void ref_demod_params(int* out)
{
  const dvbt2_parameters& d = g_demod->dvbt2;
  int v[16] = {d.fft_size, d.k_total, d.l_nulls, d.c_p2, d.c_data, d.n_fc, d.c_fc, d.n_data, d.len_frame, d.l_fc, d.n_p2,
               d.guard_interval_size, d.k_ext, d.pilot_pattern, g_demod->l1_pre.l1_post_size, g_demod->l1_post.num_plp};
  std::memcpy(out, v, sizeof(v));
}
static void segv_backtrace(int sig)
{
  void* bt[64];
  const int n = backtrace(bt, 64);
  backtrace_symbols_fd(bt, n, 2);
  _exit(128 + sig);
}
int ref_demod_new(float sample_rate, int need_plp)
{
  if (getenv("ORACLE_BACKTRACE")) signal(SIGSEGV, segv_backtrace);
  g_demod = make_zeroed<dvbt2_demodulator>(id_sdrplay, sample_rate);
  g_taps.clear(); OracleTsSink::get().bytes.clear(); OracleTsSink::get().datagram_len.clear();
  bb_de_header* bb = g_demod->deinterleaver->qam->decoder->decoder->deheader;
  bb->set_out(bb_de_header::out_network, 7654, QString("x"), need_plp);
  return 0;
}
int ref_demod_feed(int len, int16_t* i_in, int16_t* q_in)
{
  // front-end side of the signal_estimate protocol (rx_sdrplay.cpp:158-197,232-243)
  g_sig.frequency_changed = true; g_sig.gain_changed = true;
  g_demod->execute(len, i_in, q_in, &g_sig);
  g_sig.change_frequency = false; g_sig.change_gain = false;
  int st = (g_demod->crc32_l1_pre ? 1 : 0) | (g_demod->demodulator_init ? 2 : 0) | (g_demod->deint_start ? 4 : 0) |
           (g_sig.reset ? 8 : 0);
  g_sig.reset = false;
  return st;
}

}  // extern "C"
