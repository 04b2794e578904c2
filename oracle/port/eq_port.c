/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, strict IEEE) of the reference's per-symbol
 * channel estimation / equalisation / frequency de-interleaving and of the FFT convention.
 * Pinned against the compiled reference (oracle/_ref/libref_chain.so) in tests/test_oracle_eq.py:
 * the reference is built -Ofast (reciprocal maths, approximate sqrt), so that comparison carries the
 * tolerance SURVEY 8d states (cells <= 1e-3 relative); the CUDA path is compared with THIS port.
 *
 * Follows (paths relative to /root/reference/src):
 *   DSP/fast_math.h:25-42        65 536-entry sin/cos table, index = int(x * k + 32767) & 65535
 *   DSP/fast_math.h:61-81        atan2_approx
 *   DVB_T2/data_symbol.cpp:108-335   data_symbol::execute
 *   DVB_T2/fc_symbol.cpp:82-271      fc_symbol::execute (only scattered pilots estimate)
 *   DVB_T2/p2_symbol.cpp:89-259      equaliser half of p2_symbol::execute (centre carrier skipped)
 *   DSP/fast_fourier_transform.h:64-70  forward DFT, unnormalised, halves swapped
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float lut_sin[65536], lut_cos[65536];
static int lut_ready = 0;
#define PI_F 3.14159265358979323846f
#define PI_2_F 1.57079632679489661923f

static void lut_init(void)
{
  if (lut_ready) return;
  const float k_table = 32767.0f / (2.0f * PI_F);
  memset(lut_sin, 0, sizeof(lut_sin)); memset(lut_cos, 0, sizeof(lut_cos));
  for (int i = -32767; i < 32768; i++) {
    lut_sin[i + 32767] = sinf(i / k_table);
    lut_cos[i + 32767] = cosf(i / k_table);
  }
  lut_ready = 1;
}
void port_sincos_lut(float* s, float* c) { lut_init(); memcpy(s, lut_sin, sizeof(lut_sin)); memcpy(c, lut_cos, sizeof(lut_cos)); }

static float sin_l(float x) { const float k = 32767.0f / (2.0f * PI_F); return lut_sin[(int)(x * k + 32767) & 65535]; }
static float cos_l(float x) { const float k = 32767.0f / (2.0f * PI_F); return lut_cos[(int)(x * k + 32767) & 65535]; }

float port_atan2_approx(float y, float x)
{
  if (x == 0.0f) return y > 0.0f ? PI_2_F : -PI_2_F;
  if (y == 0.0f) return x > 0.0f ? 0.0f : -PI_F;
  float ax = fabsf(x), ay = fabsf(y);
  int min_x = ax < ay;
  float a = min_x ? ax / ay : ay / ax;
  float s = a * a;
  float r = ((-4.6496475e-2f * s + 1.5931422e-1f) * s - 3.2762276e-1f) * s * a + a;
  if (min_x) r = PI_2_F - r;
  if (x < 0.0f) r = PI_F - r;
  if (y < 0.0f) r = -r;
  return r;
}

enum { DATA_CARRIER = 1, P2CARRIER, P2PAPR_CARRIER, TRPAPR_CARRIER, SCATTERED_CARRIER, CONTINUAL_CARRIER,
       P2CARRIER_INVERTED, SCATTERED_CARRIER_INVERTED, CONTINUAL_CARRIER_INVERTED };

/*
 * kind 0: P2 (p2_symbol.cpp), 1: data symbol (data_symbol.cpp), 2: frame closing (fc_symbol.cpp).
 * freq: fft_size complex bins (re,im), already shifted; map / refer: k_total entries for THIS symbol;
 * h: the de-interleaver table the reference would pick for the symbol's parity; out: n_out complex.
 * amp_main = amp_p2 (kind 0) or amp_sp; amp_cp used by kind 1 only.
 */
void port_equalize(int kind, const float* freq, int l_nulls, int k_total, const int* map, const float* refer,
                   const int* h, float amp_main, float amp_cp, float* out, float* sro, float* phase)
{
  lut_init();
  const float* cellp = freq + 2 * l_nulls;
  const int half_total = k_total / 2;
  float angle = 0, delta_angle = 0, angle_est = 0, amp = 0, delta_amp = 0, amp_est = 0, dif_angle = 0;
  float sum_angle_1 = 0, sum_angle_2 = 0;
  float sp1r = 0, sp1i = 0, sp2r = 0, sp2i = 0;
  float amp_pilot = amp_main;
  float* buffer = (float*)malloc(sizeof(float) * 2 * k_total);
  int idx_data = 0, d = 0;
  /* first pilot (carrier 0 is always an edge pilot) */
  {
    float cr = cellp[0], ci = cellp[1], pr = refer[0];
    float er = cr * pr, ei = ci * pr;
    sp1r += er; sp1i += ei;
    angle_est = port_atan2_approx(ei, er);
    amp_est = sqrtf(cr * cr + ci * ci) / amp_pilot;
  }
  for (int i = 1; i < k_total; ++i) {
    const int second = i > half_total;
    float cr = cellp[2 * i], ci = cellp[2 * i + 1], pr = refer[i];
    int t = map[i];
    int is_pilot = 0;
    if (kind == 0) is_pilot = (t == P2CARRIER || t == P2CARRIER_INVERTED);
    else if (kind == 1) is_pilot = (t == SCATTERED_CARRIER || t == SCATTERED_CARRIER_INVERTED || t == CONTINUAL_CARRIER ||
                                    t == CONTINUAL_CARRIER_INVERTED);
    else is_pilot = (t == SCATTERED_CARRIER || t == SCATTERED_CARRIER_INVERTED);
    if (i == half_total) {
      /* centre carrier: P2 skips it altogether; data / FC keep a data cell but never estimate on a pilot */
      if (kind != 0 && t == DATA_CARRIER) { buffer[2 * idx_data] = cr; buffer[2 * idx_data + 1] = ci; ++idx_data; }
      continue;
    }
    if (t == DATA_CARRIER) { buffer[2 * idx_data] = cr; buffer[2 * idx_data + 1] = ci; ++idx_data; continue; }
    if (!is_pilot) continue;
    if (kind == 1 && (t == CONTINUAL_CARRIER || t == CONTINUAL_CARRIER_INVERTED)) amp_pilot = amp_cp;
    float er = cr * pr, ei = ci * pr;
    if (second) { sp2r += er; sp2i += ei; } else { sp1r += er; sp1i += ei; }
    angle = port_atan2_approx(ei, er);
    dif_angle = angle - angle_est;
    if (dif_angle > PI_F) dif_angle = PI_F * 2.0f - dif_angle;
    else if (dif_angle < -PI_F) dif_angle = PI_F * 2.0f + dif_angle;
    if (second) sum_angle_2 += angle; else sum_angle_1 += angle;
    delta_angle = dif_angle / (idx_data + 1);
    amp = sqrtf(cr * cr + ci * ci) / amp_pilot;
    amp_pilot = amp_main;
    delta_amp = (amp - amp_est) / (idx_data + 1);
    for (int j = 0; j < idx_data; ++j) {
      angle_est += delta_angle;
      amp_est += delta_amp;
      float dr = cos_l(angle_est) / amp_est, di = sin_l(angle_est) / amp_est;
      float br = buffer[2 * j], bi = buffer[2 * j + 1];
      /* buffer_cell[j] * conj(derotate) */
      out[2 * h[d]] = br * dr + bi * di;
      out[2 * h[d] + 1] = bi * dr - br * di;
      ++d;
    }
    idx_data = 0;
    angle_est = angle;
    amp_est = amp;
  }
  *phase = port_atan2_approx(sp2i, sp2r) + port_atan2_approx(sp1i, sp1r);
  *sro = sum_angle_2 - sum_angle_1;
  free(buffer);
}

/* fast_fourier_transform::execute: unnormalised forward DFT (sign -1) in double precision, halves swapped.
 * O(n log n) radix-2; the reference uses FFTW single precision, so comparisons carry ~1e-6 relative. */
void port_fft_shift(const float* in, int n, float* out)
{
  double* re = (double*)malloc(sizeof(double) * n); double* im = (double*)malloc(sizeof(double) * n);
  int bits = 0; while ((1 << bits) < n) ++bits;
  for (int i = 0; i < n; ++i) {
    int r = 0; for (int b = 0; b < bits; ++b) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    re[r] = in[2 * i]; im[r] = in[2 * i + 1];
  }
  for (int len = 2; len <= n; len <<= 1) {
    double ang = -2.0 * 3.14159265358979323846 / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        double wr = cos(ang * k), wi = sin(ang * k);
        double ur = re[i + k], ui = im[i + k];
        double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
        double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
        re[i + k] = ur + vr; im[i + k] = ui + vi;
        re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
      }
  }
  for (int i = 0; i < n; ++i) { int o = (i + n / 2) % n; out[2 * o] = (float)re[i]; out[2 * o + 1] = (float)im[i]; }
  free(re); free(im);
}
