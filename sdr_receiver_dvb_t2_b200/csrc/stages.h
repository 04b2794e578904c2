// Device-level entry points of the stages (all pointers device memory, everything asynchronous on ctx->stream):
// what the C-ABI functions call after staging host buffers, and what the frame pipeline (frames.cu) chains directly.
#pragma once
#include "ctx.h"

struct TiBlockDesc { long long in_off, out_off; int n_fec; };           // cell offsets of one TI block
struct DemapBlockDesc { long long cell_off; int n_cells; int first_fec; };

int t2_fft_device(t2b200_ctx* ctx, int n, const float2* d_in, int batch, float2* d_out, float2* d_tmp,
                  const short2* d_in16 = nullptr, float scale = 1.0f);
int t2_equalize_device(t2b200_ctx* ctx, int kind, int n_symbols, int per_frame, const int* d_idx, const float2* d_freq,
                       long long in_frame, long long in_sym, float2* d_out, long long out_frame, long long out_sym,
                       float* d_sro, float* d_phase, long long fb_frame);
int t2_ti_device(t2b200_ctx* ctx, int plp, const float2* d_in, float2* d_out, const TiBlockDesc* d_desc, int n_ti_blocks,
                 int max_cells, int fuse_mod, int rotate);
int t2_ti_geometry(t2b200_ctx* ctx, int plp, int* cells_per_fec, int* n_fec_max);
int t2_demap_device(t2b200_ctx* ctx, float2* d_cells, const DemapBlockDesc* d_desc, int n_ti_blocks, int max_cells,
                    int max_fec, int mod, int rotation, bool prepared, int fec_type, int code_rate, int8_t* d_llr,
                    float* d_prec, float* d_snr, const float* d_prec_in);
int t2_ldpc_device(t2b200_ctx* ctx, int code, const int8_t* d_llr, int n_cw, uint8_t* d_bits, int32_t* d_trials,
                   int32_t* d_iters, int max_trials, unsigned flags);
// geometry the symbol tables of `kind` were configured with (t2b200_eq_configure); false when they are missing
bool t2_eq_geometry(const t2b200_ctx* ctx, int kind, int* fft_size, int* n_out, int* n_symbols, int* first_symbol);
// opt-in BCH decoding in place: bits uint8[n_words][row_stride], one byte per bit, the first K_ldpc of a row are the BCH word
int t2_bch_device(t2b200_ctx* ctx, int code, uint8_t* d_bits, int row_stride, int n_words, int32_t* d_corrected);
int t2_bch_descramble_device(t2b200_ctx* ctx, int code, const uint8_t* d_in, int n_words, uint8_t* d_out);
