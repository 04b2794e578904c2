// a2: native construction of the init-time tables of one transmission mode -- the mode's cell counts
// (dvbt2_definition.cpp:20-159,161-648), the carrier-type maps and BPSK pilot references of the P2, data and
// frame-closing symbols (pilot_generator.cpp:48-132, :134-182 P2, :516-1932 continual, :1934-1960 scattered,
// :1962-2009 tone reservation, :2011-2091 frame closing, :2093-2166 modulation) -- so that the engine does not need the
// reference's pilot_generator object (or a fixture dumped from it).  SISO, 16K / 32K (one P2 symbol), PP1-PP8.
// Host code, no GPU needed.  Equality with the reference's tables, mode by mode: tests/test_pilot_tables.py.
//
// Layout differences from the reference: one sorted continual-pilot list per (FFT size, pilot pattern) instead of the
// standard's CP groups + run-time modulo; the scattered-pilot phase is computed per symbol instead of scanning the band.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/t2b200.h"
#include "fec_tables.h"

namespace {
#include "pilot_tables_data.inc"

// dvbt2_definition.h:103-113
enum { kData = 1, kP2 = 2, kP2Papr = 3, kTrPapr = 4, kScattered = 5, kContinual = 6 };
constexpr int kChips = 2624;

int fft_index(int fft_mode)
{
  switch (fft_mode) {
    case 4: case 11: return 0;      // FFTSIZE_16K, FFTSIZE_16K_T2GI
    case 5: case 7: return 1;       // FFTSIZE_32K, FFTSIZE_32K_T2GI
    default: return -1;
  }
}

// reference PRBS x^11 + x^2 + 1, all-ones start (pilot_generator.cpp:48-60)
void carrier_prbs(int len, std::vector<uint8_t>& out)
{
  out.resize(len);
  int sr = 0x7ff;
  for (int i = 0; i < len; ++i) {
    const int b = (sr ^ (sr >> 2)) & 1;
    out[i] = (uint8_t)(sr & 1);
    sr >>= 1;
    if (b) sr |= 0x400;
  }
}
inline int pn_bit(int symbol) { return (kPnSequenceBytes[symbol >> 3] >> (7 - (symbol & 7))) & 1; }

const int kDx[8] = {3, 6, 6, 12, 12, 24, 24, 6};
const int kDy[8] = {4, 2, 4, 2, 4, 2, 4, 16};

}  // namespace

extern "C" int t2b200_mode_init(int fft_mode, int carrier_mode, int pilot_pattern, int guard_interval_mode, int n_data,
                                int papr_mode, t2b200_mode* m)
{
  const int fi = fft_index(fft_mode);
  if (!m || fi < 0 || (carrier_mode != 0 && carrier_mode != 1) || pilot_pattern < 0 || pilot_pattern > 7 ||
      guard_interval_mode < 0 || guard_interval_mode > 6 || n_data < 1 || papr_mode < 0 || papr_mode > 3)
    return T2B200_ERR_ARG;
  memset(m, 0, sizeof(*m));
  m->fft_mode = fft_mode; m->carrier_mode = carrier_mode; m->pilot_pattern = pilot_pattern;
  m->guard_interval_mode = guard_interval_mode; m->papr_mode = papr_mode; m->n_data = n_data;
  // dvbt2_p2_parameters_init (SISO) + dvbt2_bwt_ext_parameters_init
  m->n_p2 = 1;
  m->c_p2 = fi ? 22432 : 8944;
  m->fft_size = fi ? 32768 : 16384;
  if (carrier_mode == 0) { m->k_total = fi ? 27265 : 13633; m->k_ext = 0; m->k_offset = fi ? 288 : 144; }
  else { m->k_total = fi ? 27841 : 13921; m->k_ext = fi ? 288 : 144; m->k_offset = 0; }
  m->l_nulls = (m->fft_size - m->k_total) / 2 + 1;
  // dvbt2_data_parameters_init
  const ModeCells& c = kModeCells[fi][carrier_mode][pilot_pattern];
  m->c_data = c.c_data; m->n_fc = c.n_fc; m->c_fc = c.c_fc;
  if (papr_mode == 2 || papr_mode == 3) {                       // PAPR_TR, PAPR_BOTH: reserved tones are not data cells
    const int tones = fi ? 288 : 144;
    if (m->c_data) m->c_data -= tones;
    if (m->n_fc) m->n_fc -= tones;
    if (m->c_fc) m->c_fc -= tones;
  }
  // combinations without a frame-closing symbol (SISO; dvbt2_definition.cpp:601-618)
  if ((guard_interval_mode == 4 && pilot_pattern == 6) || (guard_interval_mode == 0 && pilot_pattern == 3) ||
      (guard_interval_mode == 1 && pilot_pattern == 1) || (guard_interval_mode == 6 && pilot_pattern == 1)) {
    m->n_fc = 0; m->c_fc = 0;
  }
  static const int gi_num[7] = {1, 1, 1, 1, 1, 19, 19}, gi_den[7] = {32, 16, 8, 4, 128, 128, 256};
  m->guard_interval_size = m->fft_size / gi_den[guard_interval_mode] * gi_num[guard_interval_mode];
  m->l_fc = m->n_fc ? 1 : 0;
  m->len_frame = m->n_p2 + n_data;
  // amplitudes (pilot_generator.cpp:376-507)
  m->amp_p2 = fi ? sqrtf(37.0f) / 5.0f : sqrtf(31.0f) / 5.0f;
  m->amp_cp = 8.0f / 3.0f;
  static const float sp[8] = {4.0f / 3.0f, 4.0f / 3.0f, 7.0f / 4.0f, 7.0f / 4.0f, 7.0f / 3.0f, 7.0f / 3.0f, 7.0f / 3.0f, 7.0f / 3.0f};
  m->amp_sp = sp[pilot_pattern];
  m->dx = kDx[pilot_pattern]; m->dy = kDy[pilot_pattern];
  if (m->c_data == 0) return T2B200_ERR_ARG;                    // combination not defined by the standard
  return T2B200_OK;
}

extern "C" int t2b200_pilot_tables(const t2b200_mode* m, int kind, int32_t* carrier_map, float* pilot_refer)
{
  if (!m || !carrier_map || !pilot_refer) return T2B200_ERR_ARG;
  const int fi = fft_index(m->fft_mode);
  if (fi < 0) return T2B200_ERR_ARG;
  const int K = m->k_total;
  std::vector<uint8_t> prbs;
  carrier_prbs(K + m->k_offset, prbs);
  const bool tr = m->papr_mode == 2 || m->papr_mode == 3;
  const int n_tones = fi ? 288 : 144;
  const uint16_t* p2_papr = fi ? kP2Papr32k : kP2Papr16k;
  const uint16_t* tr_papr = fi ? kTrPapr32k : kTrPapr16k;
  auto modulate = [&](const int32_t* map, float* ref, int symbol, float amp_sp_or_p2) {
    const int pn = pn_bit(symbol);
    for (int n = 0; n < K; ++n) {
      const int bit = prbs[n + m->k_offset] ^ pn;
      float a = 0.0f;
      if (map[n] == kP2 || map[n] == kScattered) a = amp_sp_or_p2;
      else if (map[n] == kContinual) a = m->amp_cp;
      ref[n] = bit ? -a : a;                                    // +amp for bit 0 (pilot_generator.cpp:2102,2125)
      if (a == 0.0f) ref[n] = 0.0f;
    }
  };
  if (kind == T2B200_SYM_P2) {
    // pilots every 3rd carrier (every 6th for 32K SISO) + every carrier of the extended edges; the P2 tone-reservation
    // set is always blanked (pilot_generator.cpp:134-182,342-371)
    const int step = fi ? 6 : 3;
    for (int i = 0; i < K; ++i) carrier_map[i] = (i % step == 0) ? kP2 : kData;
    for (int i = 0; i < m->k_ext; ++i) { carrier_map[i] = kP2; carrier_map[i + K - m->k_ext] = kP2; }
    for (int i = 0; i < n_tones; ++i) carrier_map[p2_papr[i] + m->k_ext] = kP2Papr;
    modulate(carrier_map, pilot_refer, 0, m->amp_p2);
    return T2B200_OK;
  }
  if (kind == T2B200_SYM_FC) {
    if (!m->l_fc) return T2B200_ERR_STATE;
    for (int i = 0; i < K; ++i) carrier_map[i] = (i % m->dx == 0) ? kScattered : kData;
    carrier_map[0] = kScattered; carrier_map[K - 1] = kScattered;
    if (tr) for (int i = 0; i < n_tones; ++i) carrier_map[p2_papr[i] + m->k_ext] = kTrPapr;
    modulate(carrier_map, pilot_refer, m->len_frame - m->l_fc, m->amp_sp);
    return T2B200_OK;
  }
  if (kind != T2B200_SYM_DATA) return T2B200_ERR_ARG;
  const CpSet& cp = kCpSets[fi][m->pilot_pattern];
  const int n_sym = m->len_frame - m->l_fc - m->n_p2;
  const int period = m->dx * m->dy;
  for (int s = 0; s < n_sym; ++s) {
    const int symbol = m->n_p2 + s;
    int32_t* map = carrier_map + (size_t)s * K;
    for (int i = 0; i < K; ++i) map[i] = kData;
    for (int i = 0; i < cp.n_base; ++i) map[cp.base[i]] = kContinual;
    if (m->carrier_mode == 1) for (int i = 0; i < cp.n_ext; ++i) map[cp.ext[i]] = kContinual;
    // scattered pilots: (k - k_ext) mod (dx*dy) == dx * (symbol mod dy); the band edges always carry one
    int first = (m->k_ext + m->dx * (symbol % m->dy)) % period;
    for (int i = first; i < K; i += period) map[i] = kScattered;
    map[0] = kScattered; map[K - 1] = kScattered;
    if (tr) {
      const int shift = m->carrier_mode == 0 ? m->dx * (symbol % m->dy) : m->dx * ((symbol + m->k_ext / m->dx) % m->dy);
      for (int i = 0; i < n_tones; ++i) map[tr_papr[i] + shift] = kTrPapr;
    }
    modulate(map, pilot_refer + (size_t)s * K, symbol, m->amp_sp);
  }
  return T2B200_OK;
}

// Everything t2b200_eq_configure needs for the three symbol kinds of a mode, built natively.
extern "C" int t2b200_eq_configure_mode(t2b200_ctx* ctx, const t2b200_mode* m)
{
  if (!ctx || !m) return T2B200_ERR_ARG;
  const int K = m->k_total;
  const int n_sym = m->len_frame - m->l_fc - m->n_p2;
  std::vector<int32_t> map((size_t)std::max(n_sym, 1) * K), he, ho;
  std::vector<float> ref(map.size());
  int rc;
  if ((rc = t2b200_pilot_tables(m, T2B200_SYM_P2, map.data(), ref.data()))) return rc;
  if (!t2_freq_deinterleaver_tables(m->fft_size, m->c_p2, he, ho)) return T2B200_ERR_ARG;
  if ((rc = t2b200_eq_configure(ctx, T2B200_SYM_P2, 1, 0, m->fft_size, K, m->l_nulls, m->c_p2, map.data(), ref.data(),
                                he.data(), ho.data(), m->amp_p2, 0.0f))) return rc;
  if ((rc = t2b200_pilot_tables(m, T2B200_SYM_DATA, map.data(), ref.data()))) return rc;
  if (!t2_freq_deinterleaver_tables(m->fft_size, m->c_data, he, ho)) return T2B200_ERR_ARG;
  if ((rc = t2b200_eq_configure(ctx, T2B200_SYM_DATA, n_sym, m->n_p2, m->fft_size, K, m->l_nulls, m->c_data, map.data(),
                                ref.data(), he.data(), ho.data(), m->amp_sp, m->amp_cp))) return rc;
  if (m->l_fc) {
    if ((rc = t2b200_pilot_tables(m, T2B200_SYM_FC, map.data(), ref.data()))) return rc;
    if (!t2_freq_deinterleaver_tables(m->fft_size, m->n_fc, he, ho)) return T2B200_ERR_ARG;
    if ((rc = t2b200_eq_configure(ctx, T2B200_SYM_FC, 1, m->len_frame - 1, m->fft_size, K, m->l_nulls, m->n_fc, map.data(),
                                  ref.data(), he.data(), ho.data(), m->amp_sp, 0.0f))) return rc;
  }
  return T2B200_OK;
}
