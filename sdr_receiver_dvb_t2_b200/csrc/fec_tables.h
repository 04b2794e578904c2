// Host-side construction of the permutation tables of the FEC front half (K3, K4):
//   * cell de-interleaver permutation per FEC block   (reference time_deinterleaver.cpp:174-266)
//   * bit de-interleaver (column twist) + demux address table (reference llr_demapper.cpp:110-130 with
//     the constants of llr_demapper.h:64-78)
// Re-derived from the description in EN 302 755 6.1.3 / 6.2 / 6.4 the way the reference receiver
// applies them (receive direction), validated table-for-table against the reference in tests.
#pragma once
#include <cstdint>
#include <vector>

// perm[r*cells + ((L(w) + shift_r) mod cells)] = r*cells + w  for FEC block r of a TI block
void t2_cell_deinterleaver_permutation(int n_fec_blocks, int cells_per_fec, std::vector<int32_t>& perm);

// cells per FEC block (time_deinterleaver.cpp:61-116): fec_type 0 short / 1 normal, mod 0..3
int t2_cells_per_fec(int fec_type, int mod);
int t2_bits_per_cell(int mod);

// address[n] for the n-th soft bit the demapper produces inside a FEC frame (cell-major, per cell the
// order L0(I),L0(Q),L1(I),L1(Q),...): index into the de-interleaved FECFRAME.  Empty for QPSK (identity).
bool t2_demap_address_table(int fec_type, int mod, int code_rate, std::vector<int32_t>& address);

// Frequency de-interleaver address tables of the receiver (EN 302 755 clause 8.5; address_freq_deinterleaver.cpp:28-209):
// h_even / h_odd [n_cells] such that the d-th data cell of a symbol (ascending carrier order) is cell h[d] of the
// interleaving frame.  16K and 32K (the FFT sizes with one P2 symbol, the ones the reference receives).
bool t2_freq_deinterleaver_tables(int fft_size, int n_cells, std::vector<int32_t>& h_even, std::vector<int32_t>& h_odd);
