// Host-side tables and launch geometry of the front-end (frontend.cu; also used by the host emulation of its kernels).
#pragma once
#include <vector>
#include <cmath>
#include "frontend_kernels.h"

// The 64 taps of the reference's half-band filter for 8 MHz channels (DSP/filter_decimator.h:22-35), kept as doubles and
// converted to float the way its constructor does (:52-55).
static const double fe_h_fir[FE_TAPS] = {
  9.1776e-04, -8.7999e-05, -1.5371e-03, -3.5994e-04, 2.2031e-03, 1.2190e-03, -2.7671e-03, -2.5573e-03, 3.0238e-03, 4.3827e-03,
  -2.7246e-03, -6.6208e-03, 1.5959e-03, 9.0978e-03, 6.3727e-04, -1.1531e-02, -4.2324e-03, 1.3522e-02, 9.4232e-03, -1.4551e-02,
  -1.6447e-02, 1.3930e-02, 2.5643e-02, -1.0675e-02, -3.7747e-02, 3.0430e-03, 5.4821e-02, 1.3260e-02, -8.4349e-02, -5.5651e-02,
  1.7580e-01, 4.1952e-01, 4.1952e-01, 1.7580e-01, -5.5651e-02, -8.4349e-02, 1.3260e-02, 5.4821e-02, 3.0430e-03, -3.7747e-02,
  -1.0675e-02, 2.5643e-02, 1.3930e-02, -1.6447e-02, -1.4551e-02, 9.4232e-03, 1.3522e-02, -4.2324e-03, -1.1531e-02, 6.3727e-04,
  9.0978e-03, 1.5959e-03, -6.6208e-03, -2.7246e-03, 4.3827e-03, 3.0238e-03, -2.5573e-03, -2.7671e-03, 1.2190e-03, 2.2031e-03,
  -3.5994e-04, -1.5371e-03, -8.7999e-05, 9.1776e-04};

// apow[k] = (1 - r)^k, ainv[k] = (1 - r)^-k for k = 0 .. FE_TILE_IN; lut = {cos, sin} pairs of DSP/fast_math.h:27-43 (entry
// 65535 stays zero); h = the taps as floats
static inline void fe_make_tables(std::vector<double>& apow, std::vector<double>& ainv, std::vector<float>& lut, std::vector<float>& h)
{
  apow.resize(FE_TILE_IN + 1); ainv.resize(FE_TILE_IN + 1);
  const double a = 1.0 - (double)FE_DC_RATIO;
  for (int k = 0; k <= FE_TILE_IN; ++k) { apow[k] = std::pow(a, k); ainv[k] = std::pow(a, -k); }
  lut.assign(2 * 65536, 0.0f);
  const float k_table = FE_K_TABLE;
  for (int i = -32767; i < 32768; i++) { lut[2 * (i + 32767)] = cosf(i / k_table); lut[2 * (i + 32767) + 1] = sinf(i / k_table); }
  h.resize(FE_TAPS);
  for (int i = 0; i < FE_TAPS; ++i) h[i] = (float)fe_h_fir[i];
}

// upper bound of the FE_TILE_OUT-output tiles any stream of the launch needs (the exact counts are only known on the device: they
// depend on the resampler phase carried in the stream state, which lies in [-0.5, 0.5 + d))
static inline int fe_max_out_tiles(const FeChunk* chunk, int n_streams)
{
  int worst = 0;
  for (int s = 0; s < n_streams; ++s) {
    const double m = ((double)chunk[s].len_in + 1.0) / (double)chunk[s].resample + 2.0;
    const int k = (int)(m / 2.0) + 2;
    if (k > worst) worst = k;
  }
  return (worst + FE_TILE_OUT - 1) / FE_TILE_OUT;
}

// p1_symbol's frequency-shift table (p1_symbol.cpp:30-35): {sin, cos} of an angle accumulated in float
static inline void fe_make_p1_table(std::vector<float>& fq)
{
  fq.resize(2 * 1024);
  const float angle_shift = FE_TWO_PI_F / 1024.0f;
  float angle = 0.0f;
  for (int i = 0; i < 1024; ++i) { fq[2 * i] = sinf(angle); fq[2 * i + 1] = cosf(angle); angle += angle_shift; }
}
