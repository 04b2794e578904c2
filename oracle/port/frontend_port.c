/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, strict IEEE, the reference's own serial order) of the receiver
 * front-end that precedes the FFT.  Pinned against the compiled reference by tests/test_oracle_frontend.py:
 * the unmodified dvbt2_demodulator::execute is run on synthetic int16 I/Q with taps around its resampler and decimator
 * (oracle/tap/DSP, oracle/ref_chain.cc) and every chunk's derotated samples, resampler output and decimator output are
 * compared with this file's (the reference is built -Ofast, so agreement is to float rounding, not bit for bit).
 * Never linked by the product.
 *
 * Follows (paths relative to /root/reference/src):
 *   DVB_T2/dvbt2_demodulator.cpp:178-213   per-sample loop: int16 -> float, DC removal (DSP/loop_filters.hh:58-73, ratio
 *                                          1e-6), 1-bit IQ-imbalance statistics (:256-265) and correction, NCO
 *                                          (frequency_nco decremented per sample, wrapped to +-2pi, minus phase_nco) through
 *                                          the 65536-entry sin / cos tables of DSP/fast_math.h:27-43
 *   DSP/interpolator_farrow.hh:41-68       cubic Farrow resampler, mu in [-0.5, 0.5)
 *   DSP/filter_decimator.h:72-131          64-tap half-band FIR, decimation by 2, the AVX summation order
 *   DVB_T2/dvbt2_demodulator.cpp:321-330   guard-interval correlation -> fine frequency estimate (DSP/fast_math.h:62-80)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define FE_TAPS 64
#define TWO_PI_F (3.14159265358979323846f * 2.0f)

typedef struct {
  float dc_re, dc_im;           /* exponential_averager::out */
  float frequency_nco;
  float x1;                     /* interpolator_farrow::x1 */
  float delay[3][2];            /* delay_data_1, _2, _3 */
  float hist[FE_TAPS - 1][2];   /* the 63 resampler outputs before the next one, oldest first */
  int parity;                   /* filter_decimator::execute's static d */
} port_fe_state;

static float g_sin[65536], g_cos[65536];
static int g_lut_ready;
static const float k_table = 32767.0f / (2.0f * 3.14159265358979323846f);

static void lut_init(void)
{
  for (int i = -32767; i < 32768; i++) {                /* fast_math.h:33-38 */
    g_sin[i + 32767] = sinf(i / k_table);
    g_cos[i + 32767] = cosf(i / k_table);
  }
  g_lut_ready = 1;
}

/* The coefficients as floats, as filter_decimator's constructor makes them (filter_decimator.h:22-35,52-55). */
static const double h_fir[FE_TAPS] = {
  9.1776e-04, -8.7999e-05, -1.5371e-03, -3.5994e-04, 2.2031e-03, 1.2190e-03, -2.7671e-03, -2.5573e-03, 3.0238e-03, 4.3827e-03,
  -2.7246e-03, -6.6208e-03, 1.5959e-03, 9.0978e-03, 6.3727e-04, -1.1531e-02, -4.2324e-03, 1.3522e-02, 9.4232e-03, -1.4551e-02,
  -1.6447e-02, 1.3930e-02, 2.5643e-02, -1.0675e-02, -3.7747e-02, 3.0430e-03, 5.4821e-02, 1.3260e-02, -8.4349e-02, -5.5651e-02,
  1.7580e-01, 4.1952e-01, 4.1952e-01, 1.7580e-01, -5.5651e-02, -8.4349e-02, 1.3260e-02, 5.4821e-02, 3.0430e-03, -3.7747e-02,
  -1.0675e-02, 2.5643e-02, 1.3930e-02, -1.6447e-02, -1.4551e-02, 9.4232e-03, 1.3522e-02, -4.2324e-03, -1.1531e-02, 6.3727e-04,
  9.0978e-03, 1.5959e-03, -6.6208e-03, -2.7246e-03, 4.3827e-03, 3.0238e-03, -2.5573e-03, -2.7671e-03, 1.2190e-03, 2.2031e-03,
  -3.5994e-04, -1.5371e-03, -8.7999e-05, 9.1776e-04};

int port_fe_state_size(void) { return (int)sizeof(port_fe_state); }

void port_fe_reset(port_fe_state* s)
{
  memset(s, 0, sizeof(*s));
  s->x1 = -0.5f;
}

static float wrap2pi(float x)
{
  while (x > TWO_PI_F) x -= TWO_PI_F;
  while (x < -TWO_PI_F) x += TWO_PI_F;
  return x;
}

/* One chunk of dvbt2_demodulator::execute (:178-221).  theta[3] are accumulated on (the reference adds every sample of
 * one execute() call into them, :256-265).  derot / interp may be NULL.  Returns the number of decimator outputs. */
int port_fe_chunk(port_fe_state* s, const int16_t* i_in, const int16_t* q_in, int stride, int len_in, float short_to_float,
                  float c1, float c2, float frequency_est_filtered, float phase_nco, double arbitrary_resample,
                  float* derot, float* interp, int* len_interp, float* out, float* theta)
{
  if (!g_lut_ready) lut_init();
  const float dc_ratio = 1.0e-6f;
  const float delay_x = (float)arbitrary_resample;
  int n_interp = 0, n_out = 0;
  for (int i = 0; i < len_in; ++i) {
    float real = i_in[(long)i * stride] * short_to_float;
    float imag = q_in[(long)i * stride] * short_to_float;
    s->dc_re = s->dc_re + dc_ratio * (real - s->dc_re);
    real -= s->dc_re;
    s->dc_im = s->dc_im + dc_ratio * (imag - s->dc_im);
    imag -= s->dc_im;
    float sgn = real < 0 ? -1.0f : 1.0f;
    theta[0] -= imag * sgn;
    theta[1] += real * sgn;
    sgn = imag < 0 ? -1.0f : 1.0f;
    theta[2] += imag * sgn;
    real *= c2;
    imag += c1 * real;
    s->frequency_nco -= frequency_est_filtered;
    s->frequency_nco = wrap2pi(s->frequency_nco);
    const float offset_nco = wrap2pi(s->frequency_nco - phase_nco);
    const int idx = (int)(offset_nco * k_table + 32767) & 65535;
    const float nco_real = g_cos[idx], nco_imag = g_sin[idx];
    const float dre = real * nco_real - imag * nco_imag;
    const float dim = imag * nco_real + real * nco_imag;
    if (derot) { derot[2 * i] = dre; derot[2 * i + 1] = dim; }
    /* interpolator_farrow.hh:46-66, on both components */
    float a0[2], a1[2], a2[2], a3[2];
    const float in[2] = {dre, dim};
    for (int c = 0; c < 2; ++c) {
      const float even1 = s->delay[2][c] + in[c];
      const float even2 = s->delay[1][c] + s->delay[0][c];
      const float odd1 = s->delay[2][c] - in[c];
      const float odd2 = s->delay[1][c] - s->delay[0][c];
      a0[c] = 0.5625f * even2 - 0.0625f * even1;
      a1[c] = 0.125f * odd1 - 1.375f * odd2;
      a2[c] = 0.25f * (even1 - even2);
      a3[c] = 1.5f * odd2 - 0.5f * odd1;
    }
    while (s->x1 < 0.5f) {
      const float x2 = s->x1 * s->x1, x3 = x2 * s->x1;
      float v[2];
      for (int c = 0; c < 2; ++c) v[c] = a3[c] * x3 + a2[c] * x2 + a1[c] * s->x1 + a0[c];
      if (interp) { interp[2 * n_interp] = v[0]; interp[2 * n_interp + 1] = v[1]; }
      ++n_interp;
      s->x1 += delay_x;
      /* filter_decimator.h:83-128: the window is the 64 samples that end with this one */
      if (++s->parity == 2) {
        s->parity = 0;
        float lane[4][2];
        for (int l = 0; l < 4; ++l) lane[l][0] = lane[l][1] = 0.0f;
        for (int b = 0; b < 4; ++b)
          for (int l = 0; l < 4; ++l)
            for (int c = 0; c < 2; ++c) {
              float p[4];
              for (int q = 0; q < 4; ++q) {
                const int t = 16 * b + 4 * q + l;
                const float x = t < FE_TAPS - 1 ? s->hist[t][c] : v[c];
                p[q] = x * (float)h_fir[t];
              }
              lane[l][c] = lane[l][c] + ((p[0] + p[1]) + (p[2] + p[3]));
            }
        out[2 * n_out] = lane[0][0] + lane[1][0] + lane[2][0] + lane[3][0];
        out[2 * n_out + 1] = lane[0][1] + lane[1][1] + lane[2][1] + lane[3][1];
        ++n_out;
      }
      memmove(&s->hist[0][0], &s->hist[1][0], sizeof(float) * 2 * (FE_TAPS - 2));
      s->hist[FE_TAPS - 2][0] = v[0];
      s->hist[FE_TAPS - 2][1] = v[1];
    }
    s->x1 -= 1.0f;
    for (int c = 0; c < 2; ++c) { s->delay[2][c] = s->delay[1][c]; s->delay[1][c] = s->delay[0][c]; s->delay[0][c] = in[c]; }
  }
  if (len_interp) *len_interp = n_interp;
  return n_out;
}

static float atan2_approx(float y, float x)            /* fast_math.h:62-80 */
{
  const float pi = 3.14159265358979323846f, pi_2 = 1.57079632679489661923f;
  if (x == 0.0f) return y > 0.0f ? pi_2 : -pi_2;
  if (y == 0.0f) return x > 0.0f ? 0.0f : -pi;
  const float ax = fabsf(x), ay = fabsf(y);
  const int min_x = ax < ay;
  const float a = min_x ? ax / ay : ay / ax;
  const float s = a * a;
  float r = ((-4.6496475e-2f * s + 1.5931422e-1f) * s - 3.2762276e-1f) * s * a + a;
  if (min_x) r = pi_2 - r;
  if (x < 0.0f) r = pi - r;
  if (y < 0.0f) r = -r;
  return r;
}

/* dvbt2_demodulator.cpp:321-330: sym holds guard interval + fft_size samples; returns frequency_est */
float port_fe_cp_correlate(const float* sym, int fft_size, int guard)
{
  const float* cp = sym + 2 * (long)fft_size;
  float sr = 0.0f, si = 0.0f;
  for (int i = 4; i < guard - 4; ++i) {
    const float ar = cp[2 * i], ai = cp[2 * i + 1], br = sym[2 * i], bi = -sym[2 * i + 1];
    sr += ar * br - ai * bi;
    si += ar * bi + ai * br;
  }
  return atan2_approx(si, sr) / (float)(fft_size << 1);
}

/* ---- P1 correlator (DVB_T2/p1_symbol.cpp:75-178, the chain block diagram at :56-74; DSP/buffers.hh) ----------------------
 * data -> x exp(-j 2 pi f_sh t) [fq_shift, :30-35] -> delay Tc -> conj-multiply with data -> running sum over Tc -> delay 2 Tb -+
 * data -> delay Tb -> conj-multiply with the shifted data -> running sum over Tb -> delay 2 ---------------------------------x-> out
 * correlation = |out|^2.  delay_buffer<T, D> returns the sample written D steps earlier; sum_of_buffer<T, LEN> holds the sum of
 * the last LEN - 1 inputs (it subtracts the slot it is about to overwrite NEXT, buffers.hh:33-39). */
#define P1_C 542
#define P1_B 482
typedef struct {
  float fq[1024][2];
  int idx_fq, ready;
  float dc[P1_C + 1][2]; int ic;          /* delay_c */
  float db[P1_B + 1][2]; int ib;          /* delay_b */
  float dx[2 * P1_B + 1][2]; int ix;      /* delay_b_x2 */
  float d2[3][2]; int i2;                 /* delay_2 */
  float sc[P1_C][2], sum_c[2]; int isc;   /* average_c */
  float sb[P1_B][2], sum_b[2]; int isb;   /* average_b */
} port_p1_state;

int port_p1_state_size(void) { return (int)sizeof(port_p1_state); }
void port_p1_reset(port_p1_state* s)
{
  memset(s, 0, sizeof(*s));
  const float angle_shift = TWO_PI_F / 1024.0f;
  float angle = 0.0f;
  for (int i = 0; i < 1024; ++i) { s->fq[i][0] = sinf(angle); s->fq[i][1] = cosf(angle); angle += angle_shift; }   /* :30-35 */
  s->ready = 1;
}
static void p1_delay(float (*buf)[2], int len, int* idx, const float in[2], float out[2])
{
  buf[*idx][0] = in[0]; buf[*idx][1] = in[1];
  *idx = (*idx + 1) % len;
  out[0] = buf[*idx][0]; out[1] = buf[*idx][1];
}
static void p1_sum(float (*buf)[2], int len, int* idx, float sum[2], const float in[2])
{
  buf[*idx][0] = in[0]; buf[*idx][1] = in[1];
  *idx = (*idx + 1) % len;
  sum[0] = sum[0] - buf[*idx][0] + in[0];
  sum[1] = sum[1] - buf[*idx][1] + in[1];
}
/* n samples -> correlation[n] and out[n] (complex) */
void port_p1_correlate(port_p1_state* s, const float* in, int n, float* correlation, float* out)
{
  for (int i = 0; i < n; ++i) {
    const float d[2] = {in[2 * i], in[2 * i + 1]};
    const float* f = s->fq[s->idx_fq];
    s->idx_fq = (s->idx_fq + 1) & 0x3FF;
    const float sh[2] = {d[0] * f[0] - d[1] * f[1], d[0] * f[1] + d[1] * f[0]};
    float c[2], b[2], a[2], dd[2];
    p1_delay(s->dc, P1_C + 1, &s->ic, sh, c);
    const float avc[2] = {d[0] * c[0] + d[1] * c[1], d[1] * c[0] - d[0] * c[1]};          /* data * conj(c) */
    p1_delay(s->db, P1_B + 1, &s->ib, d, b);
    const float avb[2] = {sh[0] * b[0] + sh[1] * b[1], sh[1] * b[0] - sh[0] * b[1]};      /* shifted * conj(b) */
    p1_sum(s->sc, P1_C, &s->isc, s->sum_c, avc);
    p1_sum(s->sb, P1_B, &s->isb, s->sum_b, avb);
    p1_delay(s->dx, 2 * P1_B + 1, &s->ix, s->sum_c, a);
    p1_delay(s->d2, 3, &s->i2, s->sum_b, dd);
    const float o[2] = {a[0] * dd[0] - a[1] * dd[1], a[0] * dd[1] + a[1] * dd[0]};
    correlation[i] = o[0] * o[0] + o[1] * o[1];
    if (out) { out[2 * i] = o[0]; out[2 * i + 1] = o[1]; }
  }
}
