// Host execution of the front-end kernel bodies (sdr_receiver_dvb_t2_b200/csrc/frontend_kernels.h): the same source the
// GPU runs, every phase of every CTA as a plain loop, CTAs one after the other in launch order.  Built by
// tests/test_frontend_emu.py with -ffp-contract=off; compared there with the oracle (oracle/port/frontend_port.c).
#include <cstring>
#include <vector>
#include <cmath>
#include "../../sdr_receiver_dvb_t2_b200/csrc/frontend_kernels.h"
#include "../../sdr_receiver_dvb_t2_b200/csrc/frontend_tables.h"

extern "C" {

int emu_fe_state_size() { return (int)sizeof(FeStream); }

// one launch sequence of t2b200_frontend_execute: all streams, one chunk each
int emu_fe_chunk(int n_streams, const FeStream* cur, FeStream* next, const FeChunk* chunk, const int16_t* i_in, const int16_t* q_in,
                 long long in_stride, int step, float* out, long long out_stride, float* derot_out, FeResult* result)
{
  int max_in = 0;
  for (int s = 0; s < n_streams; ++s) if (chunk[s].len_in > max_in) max_in = chunk[s].len_in;
  const int nt_in = max_in > 0 ? (max_in + FE_TILE_IN - 1) / FE_TILE_IN : 1;      // tile 0 always runs: it lays out the delay line
  if (nt_in > FE_MAX_TILES) return 1;
  std::vector<FePlan> plan(n_streams);
  std::vector<double2> dc_part((size_t)n_streams * FE_MAX_TILES);
  std::vector<double> theta_part((size_t)n_streams * FE_MAX_TILES * 3);
  std::vector<float2> derot((size_t)n_streams * (max_in + 4));
  std::vector<double> apow, ainv; std::vector<float> lut, h;
  fe_make_tables(apow, ainv, lut, h);
  FeArgs A;
  A.i_in = i_in; A.q_in = q_in; A.in_stride = in_stride; A.step = step; A.chunk = chunk; A.cur = cur; A.next = next;
  A.plan = plan.data(); A.dc_part = dc_part.data(); A.theta_part = theta_part.data();
  A.derot = derot.data(); A.derot_stride = max_in + 4; A.out = reinterpret_cast<float2*>(out); A.out_stride = out_stride;
  A.result = result; A.apow = apow.data(); A.ainv = ainv.data(); A.lut_cs = reinterpret_cast<const float2*>(lut.data()); A.h = h.data();
  for (int s = 0; s < n_streams; ++s) for (int t = 0; t < nt_in; ++t) fe_dc_partial_body(A, s, t);
  for (int s = 0; s < n_streams; ++s) fe_plan_body(A, s);
  for (int s = 0; s < n_streams; ++s) for (int t = 0; t < nt_in; ++t) fe_derotate_body(A, s, t);
  const int nt_out = fe_max_out_tiles(chunk, n_streams) + 1;
  for (int s = 0; s < n_streams; ++s) for (int t = 0; t < nt_out; ++t) fe_resample_body(A, s, t, nt_out);
  if (derot_out)
    for (int s = 0; s < n_streams; ++s) std::memcpy(derot_out + 2 * (size_t)s * max_in, &derot[(size_t)s * (max_in + 4) + 3], sizeof(float2) * chunk[s].len_in);
  return 0;
}

// the NCO recurrence in jumps (fe_nco_run, with segments as the derotation pass uses them) -> all n values
void emu_nco_run(float v, float c, int n, float* values, int* n_segments)
{
  int k0[FE_MAX_SEG]; float v0[FE_MAX_SEG]; double inc[FE_MAX_SEG];
  int ns = 0, nd = 0;
  float w = fe_nco_run(v, c, n, FE_MAX_SEG, k0, v0, inc, &ns, &nd);
  for (int k = nd; k < n; ++k) { w = fe_wrap(fe_add(w, c)); values[k] = w; }
  for (int i = 0; i < nd; ++i) {
    int g = 0;
    while (g + 1 < ns && k0[g + 1] <= i) ++g;
    values[i] = (float)((double)v0[g] + (double)(i + 1 - k0[g]) * inc[g]);
  }
  *n_segments = ns;
}

// stress: `cases` random (v, c, n) triples, the jumps against n real float additions; returns the number of mismatching cases
int emu_nco_stress(unsigned seed, int cases, float* bad_v, float* bad_c, int* bad_n)
{
  unsigned long long st = seed * 2654435761ull + 88172645463325252ull;
  auto rnd = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  int bad = 0;
  std::vector<float> got(FE_TILE_IN);
  for (int k = 0; k < cases; ++k) {
    float v = rnd() < 0.7 ? (float)((rnd() * 2 - 1) * 6.2831) : (float)((rnd() * 2 - 1) * std::pow(10.0, -8.0 * rnd()));
    if (rnd() < 0.05) v = std::ldexp(1.0f, (int)(rnd() * 6) - 3) * (rnd() < 0.5 ? 1.f : -1.f);          // exact powers of two
    float c = (float)((rnd() < 0.5 ? 1 : -1) * std::pow(10.0, -9.0 + 7.5 * rnd()));
    if (rnd() < 0.1) c = std::ldexp(1.0f, -(int)(rnd() * 30) - 3) * (rnd() < 0.5 ? 1.f : -1.f) * (rnd() < 0.5 ? 1.0f : 1.5f);   // ties
    const int n = 1 + (int)(rnd() * (FE_TILE_IN - 1));
    int ns = 0;
    emu_nco_run(v, c, n, got.data(), &ns);
    const float end = fe_nco_run(v, c, n, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    float w = v; bool ok = true;
    for (int i = 0; i < n; ++i) {
      w = fe_wrap(fe_add(w, c));
      unsigned a, b; std::memcpy(&a, &w, 4); std::memcpy(&b, &got[i], 4);
      if (a != b) { ok = false; break; }
    }
    unsigned a, b; std::memcpy(&a, &w, 4); std::memcpy(&b, &end, 4);
    if (ok && a != b) ok = false;
    if (!ok) { if (bad == 0) { *bad_v = v; *bad_c = c; *bad_n = n; } ++bad; }
  }
  return bad;
}

// the same recurrence without segments: only the end value (as fe_plan_body walks tile boundaries)
float emu_nco_end(float v, float c, int n) { return fe_nco_run(v, c, n, 0, nullptr, nullptr, nullptr, nullptr, nullptr); }

void emu_p1_correlate(const float* x, int n, const float* hist, int i0, float* correlation, float* out)
{
  std::vector<float> fq; fe_make_p1_table(fq);
  std::vector<double2> prefix(2 * (size_t)(n + FE_P1_LEAD + 1));
  fe_p1_correlate_body(reinterpret_cast<const float2*>(x), n, reinterpret_cast<const float2*>(hist),
                       reinterpret_cast<const float2*>(fq.data()), i0, prefix.data(), correlation, reinterpret_cast<float2*>(out));
}

void emu_cp_correlate(const float* sym, int fft_size, int guard, float* est)
{
  fe_cp_correlate_body(reinterpret_cast<const float2*>(sym), fft_size, guard, est);
}

}
