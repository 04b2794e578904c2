#include "fec_tables.h"
#include <cmath>

int t2_bits_per_cell(int mod) { return 2 * (mod + 1); }
int t2_cells_per_fec(int fec_type, int mod)
{
  const int n = fec_type ? 64800 : 16200;
  return n / t2_bits_per_cell(mod);
}

void t2_cell_deinterleaver_permutation(int n_fec_blocks, int cells, std::vector<int32_t>& perm)
{
  int deg = 0;
  while ((1 << deg) < cells) ++deg;                       // ceil(log2(cells))
  const int states = 1 << deg;
  // taps of the degree-(deg-1) register per EN 302 755 table 22 (as the receiver uses them)
  static const int taps11[] = {0, 3}, taps12[] = {0, 2}, taps13[] = {0, 1, 4, 6}, taps14[] = {0, 1, 4, 5, 9, 11},
                   taps15[] = {0, 1, 2, 12};
  const int* taps; int ntaps;
  switch (deg) {
    case 11: taps = taps11; ntaps = 2; break;
    case 12: taps = taps12; ntaps = 2; break;
    case 13: taps = taps13; ntaps = 4; break;
    case 15: taps = taps15; ntaps = 4; break;
    default: taps = taps14; ntaps = 6; break;             // 14, and the reference's fall-back
  }
  const int low_mask = (1 << (deg - 1)) - 1;
  std::vector<int> base; base.reserve(cells);             // L(w), w = 0..cells-1
  int reg = 0;
  for (int i = 0; i < states; ++i) {
    if (i < 2) reg = 0;
    else if (i == 2) reg = 1;
    else {
      int fb = 0;
      for (int k = 0; k < ntaps; ++k) fb ^= (reg >> taps[k]) & 1;
      reg = ((reg & low_mask) >> 1) | (fb << (deg - 2));
    }
    reg |= (i & 1) << (deg - 1);                          // toggling MSB
    if (reg < cells) base.push_back(reg);
  }
  perm.assign((size_t)n_fec_blocks * cells, 0);
  int counter = 0;
  for (int r = 0; r < n_fec_blocks; ++r) {
    int shift;
    do {                                                  // bit-reversed counter, one extra left shift,
      int t = counter++, rev = 0;                         // values >= cells skipped
      for (int p = 0; p < deg; ++p) { rev |= t & 1; rev <<= 1; t >>= 1; }
      shift = rev;
    } while (shift >= cells);
    for (int w = 0; w < cells; ++w) perm[(size_t)r * cells + (base[w] + shift) % cells] = r * cells + w;
  }
}

bool t2_demap_address_table(int fec_type, int mod, int code_rate, std::vector<int32_t>& address)
{
  // column-twist parameters tc and demux tables, EN 302 755 tables 6.1.3-x / 6.2.1-x in the receive
  // orientation used by the reference (llr_demapper.h:64-78)
  static const int tc16s[8] = {0, 0, 0, 1, 7, 20, 20, 21}, tc16n[8] = {0, 0, 2, 4, 4, 5, 7, 7};
  static const int tc64s[12] = {0, 0, 0, 2, 2, 2, 3, 3, 3, 6, 7, 7}, tc64n[12] = {0, 0, 2, 2, 3, 4, 4, 5, 5, 7, 8, 9};
  static const int tc256s[8] = {0, 0, 0, 1, 7, 20, 20, 21};
  static const int tc256n[16] = {0, 2, 2, 2, 2, 3, 7, 15, 16, 20, 22, 22, 27, 27, 28, 32};
  static const int dm16[8] = {7, 1, 3, 5, 2, 4, 6, 0}, dm16_35[8] = {0, 2, 3, 6, 4, 1, 7, 5};
  static const int dm64[12] = {11, 8, 5, 2, 10, 7, 4, 1, 9, 6, 3, 0}, dm64_35[12] = {4, 6, 0, 5, 8, 10, 2, 1, 7, 3, 11, 9};
  static const int dm256s[8] = {7, 2, 4, 1, 6, 3, 5, 0};
  static const int dm256n[16] = {15, 1, 13, 3, 10, 7, 9, 11, 4, 6, 8, 5, 12, 2, 14, 0};
  static const int dm256n_35[16] = {4, 6, 0, 2, 3, 14, 12, 10, 7, 5, 8, 1, 15, 9, 11, 13};
  static const int dm256n_23[16] = {3, 15, 1, 7, 4, 11, 5, 0, 12, 2, 9, 14, 13, 6, 8, 10};
  address.clear();
  if (mod == 0) return true;                                // QPSK: no bit interleaver / demux on this path
  const int n = fec_type ? 64800 : 16200;
  int ncols; const int *tc, *dm;
  switch (mod) {
    case 1: ncols = 8;  tc = fec_type ? tc16n : tc16s; dm = (fec_type && code_rate == 1) ? dm16_35 : dm16; break;
    case 2: ncols = 12; tc = fec_type ? tc64n : tc64s; dm = (fec_type && code_rate == 1) ? dm64_35 : dm64; break;
    case 3:
      if (fec_type) { ncols = 16; tc = tc256n; dm = code_rate == 1 ? dm256n_35 : code_rate == 2 ? dm256n_23 : dm256n; }
      else { ncols = 8; tc = tc256s; dm = dm256s; }
      break;
    default: return false;
  }
  const int nrows = n / ncols;
  std::vector<int32_t> twist((size_t)n);
  for (int r = 0; r < nrows; ++r)
    for (int c = 0; c < ncols; ++c) twist[(size_t)r * ncols + c] = nrows * c + (r + nrows - tc[c]) % nrows;
  address.resize(n);
  for (int i = 0; i < n; ++i) address[i] = twist[(i / ncols) * ncols + dm[i % ncols]];
  return true;
}

// ---- frequency de-interleaver ----------------------------------------------------------------------------------
// The interleaver permutes the cells of a symbol by H(q): a shift register R' of Nr - 1 bits (Nr = log2 FFT size) steps
// through a maximal-length sequence, a fixed wire permutation (different for even and odd symbols below 32K) turns R'_i
// into R_i, the toggling top bit (i mod 2) is put in front, and values outside the symbol are skipped.  The receiver
// needs the inverse mapping; in 32K the even-symbol permutation IS the inverse of the odd one
// (address_freq_deinterleaver.cpp:149-155,185-196), so there h_even is the odd forward table itself.
bool t2_freq_deinterleaver_tables(int fft_size, int n_cells, std::vector<int32_t>& h_even, std::vector<int32_t>& h_odd)
{
  static const int perm16_even[13] = {9, 7, 6, 10, 12, 5, 1, 11, 0, 2, 3, 4, 8};       // EN 302 755 table 74 (16K)
  static const int perm16_odd[13] = {6, 8, 10, 12, 2, 0, 4, 1, 11, 3, 5, 9, 7};
  static const int perm32[14] = {7, 13, 3, 4, 9, 2, 12, 11, 1, 8, 10, 0, 5, 6};        // (32K)
  static const int taps16[6] = {0, 1, 4, 5, 9, 11}, taps32[4] = {0, 1, 2, 12};
  const int *pe, *po, *taps; int ntaps, nbits;
  if (fft_size == 16384) { pe = perm16_even; po = perm16_odd; taps = taps16; ntaps = 6; nbits = 13; }
  else if (fft_size == 32768) { pe = perm32; po = perm32; taps = taps32; ntaps = 4; nbits = 14; }
  else return false;
  if (n_cells <= 0 || n_cells > fft_size) return false;
  std::vector<int32_t> fwd_even, fwd_odd;                       // H(q) of the transmitter, even / odd symbols
  fwd_even.reserve(n_cells); fwd_odd.reserve(n_cells);
  unsigned reg = 0;
  for (int i = 0; i < fft_size; ++i) {
    if (i < 2) reg = 0;
    else if (i == 2) reg = 1;
    else {
      unsigned fb = 0;
      for (int k = 0; k < ntaps; ++k) fb ^= (reg >> taps[k]) & 1u;
      reg = ((reg & ((1u << nbits) - 1u)) >> 1) | (fb << (nbits - 1));
    }
    unsigned e = 0, o = 0;
    for (int n = 0; n < nbits; ++n) { e |= ((reg >> n) & 1u) << pe[n]; o |= ((reg >> n) & 1u) << po[n]; }
    const unsigned top = (unsigned)(i & 1) * (unsigned)(fft_size / 2);
    if ((int)(e + top) < n_cells) fwd_even.push_back((int32_t)(e + top));
    if ((int)(o + top) < n_cells) fwd_odd.push_back((int32_t)(o + top));
  }
  if ((int)fwd_even.size() != n_cells || (int)fwd_odd.size() != n_cells) return false;
  h_even.assign(n_cells, 0); h_odd.assign(n_cells, 0);
  for (int i = 0; i < n_cells; ++i) h_odd[fwd_odd[i]] = i;
  if (fft_size == 32768) h_even = fwd_odd;                      // inverse of the inverse
  else for (int i = 0; i < n_cells; ++i) h_even[fwd_even[i]] = i;
  return true;
}
