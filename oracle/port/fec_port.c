/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, strict IEEE: built with -fno-fast-math
 * -ffp-contract=off) of the reference's time de-interleaver and soft demapper.  Pinned against the
 * compiled reference (oracle/_ref/libref_chain.so) by tests/test_oracle_fec.py and against the golden
 * vectors that reference produced (tests/golden/fec_ref.npz).  Never linked by the product.
 *
 * Follows (paths relative to /root/reference/src/DVB_T2):
 *   time_deinterleaver.cpp:174-266   address_cell_deinterleaving
 *   time_deinterleaver.cpp:316-374   the per-cell loop of execute() incl. the deferred Q write
 *   llr_demapper.cpp:110-130         address_generator
 *   llr_demapper.cpp:160-228         qpsk      :230-364 qam16      :366-535 qam64      :537-768 qam256
 *   llr_demapper.cpp:770-776         quantize (QPSK only)
 * Floating point: the reference is built -Ofast, so its own summation order for sum_s / sum_e is the
 * compiler's choice; this port adds in program order in float.  Everything else is order-free.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* time_deinterleaver.cpp:174-266 */
void port_cell_permutation(int block_max, int cells_size, int* permutations)
{
  int pn_degree = (int)ceil(log2((double)cells_size));
  int max_states = 1 << pn_degree;
  static const int l11[] = {0, 3}, l12[] = {0, 2}, l13[] = {0, 1, 4, 6}, l14[] = {0, 1, 4, 5, 9, 11}, l15[] = {0, 1, 2, 12};
  const int* logic; int xor_size, pn_mask;
  switch (pn_degree) {
    case 11: logic = l11; xor_size = 2; pn_mask = 0x3ff; break;
    case 12: logic = l12; xor_size = 2; pn_mask = 0x7ff; break;
    case 13: logic = l13; xor_size = 4; pn_mask = 0xfff; break;
    case 15: logic = l15; xor_size = 4; pn_mask = 0x3fff; break;
    default: logic = l14; xor_size = 6; pn_mask = 0x1fff; break;
  }
  int* first = (int*)malloc(sizeof(int) * max_states);
  int q = 0, lfsr = 0;
  for (int i = 0; i < max_states; ++i) {
    if (i == 0 || i == 1) lfsr = 0;
    else if (i == 2) lfsr = 1;
    else {
      int result = 0;
      for (int k = 0; k < xor_size; ++k) result ^= (lfsr >> logic[k]) & 1;
      lfsr &= pn_mask; lfsr >>= 1; lfsr |= result << (pn_degree - 2);
    }
    lfsr |= (i % 2) << (pn_degree - 1);
    if (lfsr < cells_size) first[q++] = lfsr;
  }
  int n = 0, index = 0, address = 0;
  for (int r = 0; r < block_max; r++) {
    int shift = cells_size;
    while (shift >= cells_size) {
      int temp = n; shift = 0;
      for (int p = 0; p < pn_degree; ++p) { shift |= temp & 1; shift <<= 1; temp >>= 1; }
      n++;
    }
    for (int w = 0; w < cells_size; ++w) permutations[((first[w] + shift) % cells_size) + index] = address++;
    index += cells_size;
  }
  free(first);
}

/* One TI block through the cell loop of time_deinterleaver::execute (time_deinterleaver.cpp:316-340).
 * in/out: interleaved re,im floats.  The pending first-cell Q write is carried in *end_cell / *q_first
 * across calls exactly like the member variables (pass zero-initialised storage for a fresh receiver). */
void port_ti_block(const float* in, int n_fec, int cells_per_fec, const int* perm, float* out,
                   int* end_cell, float* q_first)
{
  const int num_rows = cells_per_fec / 5, ti_block_size = 5 * n_fec * num_rows;
  int idx_step = 0, idx_row = 0;
  for (int i = 0; i < ti_block_size; ++i) {
    int d = idx_step + idx_row;
    int i_address = perm[d];
    int q_address = i_address - 1;
    if (i_address % cells_per_fec == 0) {
      if (i_address != 0) out[2 * (*end_cell) + 1] = *q_first;
      *q_first = in[2 * i + 1];
      *end_cell = q_address + cells_per_fec;
    } else {
      out[2 * q_address + 1] = in[2 * i + 1];
    }
    out[2 * i_address] = in[2 * i];
    idx_step += num_rows;
    if (idx_step == ti_block_size) {
      out[2 * (*end_cell) + 1] = *q_first;
      idx_step = 0;
      ++idx_row;
    }
  }
}

/* llr_demapper.cpp:110-130 (argument names as there: _column is the longer dimension) */
void port_demap_address(int column, int row, const int* tc, const int* demux, int* out)
{
  int* address = (int*)malloc(sizeof(int) * column * row);
  for (int c = 0; c < column; ++c)
    for (int r = 0; r < row; ++r) address[c * row + r] = column * r + (c + column - tc[r]) % column;
  int k = 0, n = 0;
  for (int i = 0; i < column * row; ++i) {
    out[i] = address[demux[n] + k];
    if (++n == row) { n = 0; k += row; }
  }
  free(address);
}

static const float kRot[4] = {0.506145483f, 0.293215314f, 0.150098316f, 0.062418810f};
static const float kNorm[4] = {0.707106781f, 0.316227766f, 0.15430335f, 0.076696499f};

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

/* the table selection of qam16/qam64/qam256 (llr_demapper.cpp:294-302, 455-463, 677-686) */
void port_demap_address_for(int fec_normal, int mod, int code_rate, int* out)
{
  if (mod == 1) {
    if (fec_normal) port_demap_address(8100, 8, tc16n, code_rate == 1 ? dm16_35 : dm16, out);
    else port_demap_address(2025, 8, tc16s, dm16, out);
  } else if (mod == 2) {
    if (fec_normal) port_demap_address(5400, 12, tc64n, code_rate == 1 ? dm64_35 : dm64, out);
    else port_demap_address(1350, 12, tc64s, dm64, out);
  } else if (mod == 3) {
    if (fec_normal) port_demap_address(4050, 16, tc256n, code_rate == 1 ? dm256n_35 : code_rate == 2 ? dm256n_23 : dm256n, out);
    else port_demap_address(2025, 8, tc256s, dm256s, out);
  } else {
    int n = fec_normal ? 64800 : 16200;
    for (int i = 0; i < n; ++i) out[i] = i;
  }
}

/* hard slicing ladders, one axis */
static float slice(int mod, float x, float a)
{
  if (mod == 0) return x > 0 ? a : -a;
  if (mod == 1) {
    if (x > 0) return x > a * 2.0f ? a * 3.0f : a;
    return x < -(a * 2.0f) ? -(a * 3.0f) : -a;
  }
  if (mod == 2) {
    if (x > 0) {
      if (x > a * 4.0f) return x > a * 6.0f ? a * 7.0f : a * 5.0f;
      return x > a * 2.0f ? a * 3.0f : a;
    }
    if (x < -(a * 4.0f)) return x > a * 6.0f ? -(a * 7.0f) : -(a * 5.0f);     /* :407,:427 */
    return x < -(a * 2.0f) ? -(a * 3.0f) : -a;
  }
  if (x > 0) {
    if (x > a * 8.0f) { if (x > a * 12.0f) return x > a * 14.0f ? a * 15.0f : a * 13.0f; return x > a * 10.0f ? a * 11.0f : a * 9.0f; }
    if (x > a * 4.0f) return x > a * 6.0f ? a * 7.0f : a * 5.0f;
    return x > a * 2.0f ? a * 3.0f : a;
  }
  if (x < -(a * 8.0f)) { if (x < -(a * 12.0f)) return x < -(a * 14.0f) ? -(a * 15.0f) : -(a * 13.0f); return x < -(a * 10.0f) ? -(a * 11.0f) : -(a * 9.0f); }
  if (x < -(a * 4.0f)) return x < -(a * 6.0f) ? -(a * 7.0f) : -(a * 5.0f);
  return x < -(a * 2.0f) ? -(a * 3.0f) : -a;
}

/* Restates the PRODUCT option T2B200_OPT_DEMAP_SATURATE (include/t2b200.h), not a reference behaviour: clamp the LLR to
 * [-128, 127] instead of the reference's wrapping cast.  Off by default; bench.py turns it on to give the reference's LDPC
 * decoder the same (decodable) LLRs the GPU arm decodes, tests use it to check the option bit for bit. */
static int g_demap_saturate = 0;
void port_set_demap_saturate(int on) { g_demap_saturate = on != 0; }

/* (int8_t)(float) the way x86-64 gcc does it: cvttss2si r32 (0x80000000 if out of range), low byte */
static int8_t cast_i8(float r)
{
  if (g_demap_saturate) return (int8_t)fminf(fmaxf(r, -128.0f), 127.0f);
  if (!(fabsf(r) < 2147483648.0f)) return 0;
  return (int8_t)((int32_t)r & 0xff);
}

/*
 * One TI block through llr_demapper::execute.  cells (re,im floats) are derotated IN PLACE when
 * rotation != 0.  llr: int8[n_fec][fec_size].  precision_in > 0 overrides the computed precision.
 * Returns the precision used; *snr gets the emitted SNR value.
 */
float port_demap(float* cells, int n_cells, int mod, int rotation, int fec_normal, int code_rate,
                 int8_t* llr, float* snr, float precision_in)
{
  const float a = kNorm[mod];
  const int fec_size = fec_normal ? 64800 : 16200;
  const int bpc = 2 * (mod + 1);
  if (rotation) {
    const float rc = (float)cos(-(double)kRot[mod]), rs = (float)sin(-(double)kRot[mod]);
    for (int i = 0; i < n_cells; ++i) {
      float re = cells[2 * i] * rc - cells[2 * i + 1] * rs;
      float im = cells[2 * i] * rs + cells[2 * i + 1] * rc;
      cells[2 * i] = re; cells[2 * i + 1] = im;
    }
  }
  float sum_s = 0, sum_e = 0;
  const int n_stat = mod == 0 ? (n_cells < 2048 ? n_cells : 2048) : n_cells;
  for (int i = 0; i < n_stat; ++i) {
    float sx = slice(mod, cells[2 * i], a), sy = slice(mod, cells[2 * i + 1], a);
    float ex = cells[2 * i] - sx, ey = cells[2 * i + 1] - sy;
    sum_s += sx * sx + sy * sy;
    sum_e += ex * ex + ey * ey;
  }
  *snr = (mod == 0 ? 10.0f : 20.0f) * log10f(sum_s / sum_e);
  float precision = 8.0f * a * sum_s / sum_e;
  if (precision_in > 0) precision = precision_in;
  int* address = (int*)malloc(sizeof(int) * fec_size);
  port_demap_address_for(fec_normal, mod, code_rate, address);
  const int cpf = fec_size / bpc;
  for (int i = 0; i < n_cells; ++i) {
    int8_t* out = llr + (size_t)(i / cpf) * fec_size;
    const int* ad = address + bpc * (i % cpf);
    float xi = cells[2 * i], xq = cells[2 * i + 1];
    for (int l = 0; l <= mod; ++l) {
      float ri = nearbyintf(xi * precision), rq = nearbyintf(xq * precision);
      if (mod == 0) {                                         /* quantize(): saturating */
        ri = fminf(fmaxf(ri, -128.0f), 127.0f); rq = fminf(fmaxf(rq, -128.0f), 127.0f);
      }
      out[ad[2 * l]] = cast_i8(ri);
      out[ad[2 * l + 1]] = cast_i8(rq);
      if (l < mod) {
        float t = a * (float)(1 << (mod - l));
        xi = fabsf(xi) - t; xq = fabsf(xq) - t;
      }
    }
  }
  free(address);
  return precision;
}
