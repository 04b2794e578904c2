// Receiver front-end: the bodies of the CUDA kernels of frontend.cu.
//
// What they replace (paths relative to the reference's src/): the per-sample loop of dvbt2_demodulator::execute
// (DVB_T2/dvbt2_demodulator.cpp:178-213: int16 -> float, DC removal, IQ-imbalance statistics and correction, NCO), the
// Farrow resampler (DSP/interpolator_farrow.hh:41-68) and the half-band decimator (DSP/filter_decimator.h:72-131).  The
// reference walks one stream sample by sample with five recurrences (DC average, NCO phase, resampler phase, delay lines,
// decimator phase); here every recurrence is put in closed form per chunk so that all samples of all streams of a launch are
// independent:
//   * DC average   y_i = (1 - r) y_{i-1} + r x_i      -> weighted prefix sums (double), one partial per 1 024-sample tile
//   * NCO phase    s_{i+1} = wrap(fl(s_i - f))        -> evaluated EXACTLY: inside one binade of s a float addition of a
//                                                        constant moves s by a constant number of ulps, so the chunk falls
//                                                        apart into a few linear segments (fe_nco_run); the steps that leave
//                                                        a binade, hit an exact tie or wrap at 2 pi are done in real float
//                                                        arithmetic.  The phase indexes a 65 536-entry table (fast_math.h), so
//                                                        an approximate phase would pick other entries than the reference.
//   * resampler    output m sits at X_m = x1 + m d    -> input index floor(X_m + 1/2), mu = X_m - index (double)
//   * decimator    output k = resampler output m0 + 2k, 64 taps, summed in the order of the reference's AVX loop
//
// The file compiles for the device (frontend.cu) and for the host (tests/cpp/frontend_emu.cpp): a kernel body is a sequence
// of PHASES, each a loop over "threads" followed by a barrier; on the host a phase is a plain loop, so the same source is
// executed on the CPU against the oracle (tests/test_frontend_emu.py) before it ever sees a GPU.  Data that lives across a
// barrier is in (what is on the device) shared memory; within a phase no thread reads what another thread writes.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define FE_FN __device__ __forceinline__
#define FE_SHARED __shared__
#define FE_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define FE_SYNC() __syncthreads()
#define FE_ONE() if (threadIdx.x == 0)
#define FE_LDG(p) __ldg(p)
#define FE_UNROLL4 _Pragma("unroll 4")
FE_FN float fe_mul(float a, float b) { return __fmul_rn(a, b); }      // never contracted into an FMA: the CPU does not either
FE_FN float fe_add(float a, float b) { return __fadd_rn(a, b); }
FE_FN float fe_sub(float a, float b) { return __fsub_rn(a, b); }
#else
#define FE_FN static inline
#define FE_SHARED
#define FE_FOR(i, n) for (int i = 0; i < (n); ++i)
#define FE_SYNC()
#define FE_ONE()
#define FE_LDG(p) (*(p))
#define FE_UNROLL4
FE_FN float fe_mul(float a, float b) { return a * b; }                // the emulation is built with -ffp-contract=off
FE_FN float fe_add(float a, float b) { return a + b; }
FE_FN float fe_sub(float a, float b) { return a - b; }
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
#endif

enum {
  FE_TILE_IN = 1024,         // input samples per CTA of the derotation pass
  FE_TILE_OUT = 1024,        // decimator outputs per CTA of the resampling pass (four per thread)
  FE_THREADS = 256,
  FE_TAPS = 64,
  FE_MAX_SEG = 48,           // linear NCO segments kept per tile before the planner falls back to single steps
  FE_MAX_CHUNK_SEG = 96,     // linear NCO segments of a whole chunk (more: every tile plans for itself)
  FE_MAX_TILES = 256         // input tiles per chunk (262 144 samples)
};
#define FE_DC_RATIO 1.0e-6f  /* dvbt2_demodulator.h:88 */
#define FE_TWO_PI_F (3.14159265358979323846f * 2.0f)
#define FE_K_TABLE (32767.0f / (2.0f * 3.14159265358979323846f))

struct FeStream {            // state of one stream between chunks (two copies: a launch reads one and writes the other)
  float dc_re, dc_im;        // exponential_averager::out
  float frequency_nco;
  float x1;                  // interpolator_farrow::x1
  float2 delay[3];           // delay_data_1, _2, _3: the last three derotated samples, newest first
  float2 hist[FE_TAPS - 1];  // the 63 resampler outputs before the next one, oldest first
  int parity;                // filter_decimator::execute's static d
  int pad;
};

struct FeChunk {             // what the host loop knows about one chunk of one stream (t2b200_fe_chunk)
  int len_in;
  float short_to_float, c1, c2, frequency_est_filtered, phase_nco, resample;
};

struct FePlan {              // written by fe_plan_body, read by the two passes
  int n_interp, n_out, m0, n_tiles_in;
  float x1_next; int parity_next;
  float nco_next; float dc_re_next, dc_im_next;
  float nco_start[FE_MAX_TILES];         // frequency_nco before the first sample of the tile
  int n_seg;                             // > 0: the NCO phase of the whole chunk as linear segments (fe_nco_run) ...
  int seg_k0[FE_MAX_CHUNK_SEG]; float seg_v0[FE_MAX_CHUNK_SEG]; double seg_inc[FE_MAX_CHUNK_SEG];
                                         // ... 0: the chunk needs more segments than that, every tile plans from nco_start
  double2 dc_start[FE_MAX_TILES];        // DC average before the first sample of the tile
};

struct FeResult { int len_out, len_interp; float theta1, theta2, theta3; };   // t2b200_fe_result

struct FeArgs {
  const int16_t* i_in; const int16_t* q_in;     // sample n of stream s at [s * in_stride + n * step]
  long long in_stride; int step;
  const FeChunk* chunk;                          // [n_streams]
  const FeStream* cur; FeStream* next;           // [n_streams]
  FePlan* plan;                                  // [n_streams]
  double2* dc_part;                              // [n_streams][FE_MAX_TILES]
  double* theta_part;                            // [n_streams][FE_MAX_TILES][3]
  float2* derot; long long derot_stride;         // [n_streams][derot_stride]: delay_data_3, _2, _1 of the carried state, then the chunk
  float2* out; long long out_stride;             // [n_streams][out_stride]
  FeResult* result;                              // [n_streams]
  const double* apow;                            // (1 - r)^k, k = 0 .. FE_TILE_IN
  const double* ainv;                            // (1 - r)^-k
  const float2* lut_cs;                          // {cos, sin} tables of DSP/fast_math.h
  const float* h;                                // the 64 decimator taps as floats
};

#ifdef __CUDACC__
__constant__ float fe_c_h[FE_TAPS];                // the taps again: with compile-time indices they are operands, not loads
#define FE_H(A, t) fe_c_h[t]
#else
#define FE_H(A, t) (A).h[t]
#endif

// ---- NCO ---------------------------------------------------------------------------------------------------------------
FE_FN float fe_wrap(float x)
{
  while (x > FE_TWO_PI_F) x = fe_sub(x, FE_TWO_PI_F);
  while (x < -FE_TWO_PI_F) x = fe_add(x, FE_TWO_PI_F);
  return x;
}

// Advance the recurrence v <- wrap(fl(v + c)) by n steps, bit for bit as n float additions would, in a few jumps.  When seg_*
// are given, the run is recorded as linear segments: v_{k0 + j} = v0 + j * inc for 0 <= j <= len (k0 counted from the
// start of the run, v_0 = the start value); at most max_seg segments, *n_done steps were covered (the caller finishes the rest
// one step at a time).  Returns the value after the steps it covered.
FE_FN float fe_nco_run(float v, float c, int n, int max_seg, int* seg_k0, float* seg_v0, double* seg_inc, int* n_seg, int* n_done)
{
  int k = 0, ns = 0;
  while (k < n && (max_seg == 0 || ns < max_seg)) {
    int len = 0;
    double inc = 0.0;
    const float av = fabsf(v);
    if (av >= 1.17549435e-38f && av <= FE_TWO_PI_F) {
      int e;
      frexpf(av, &e);                                       // av = m * 2^e, m in [0.5, 1): the binade is [2^(e-1), 2^e)
      const double u = ldexp(1.0, e - 24);                  // its ulp
      const double lo = ldexp(1.0, e - 1), hi = ldexp(1.0, e);
      const double sgn = v < 0 ? -1.0 : 1.0;
      const double cm = sgn * (double)c;                    // > 0: the magnitude grows
      const double d = cm / u;
      const double q = nearbyint(d);
      if (fabs(d - q) != 0.5 && fabs(q) < 16777216.0) {     // an exact tie depends on the parity of the sum: single step
        const double qi = q * u;                            // |v| moves by exactly qi per step while |v| + cm stays in the binade
        // step j+1 (from |v| + j qi) is regular iff lo <= |v| + j qi + cm < hi and the result stays <= 2 pi
        double jmax;
        const double a = (double)av;
        if (a + cm < lo || a + cm >= hi || a + qi > (double)FE_TWO_PI_F) jmax = 0;
        else if (qi == 0.0) jmax = n - k;
        else {
          double est = qi > 0 ? floor((hi - a - cm) / qi) + 1 : floor((a + cm - lo) / (-qi)) + 1;
          if (qi > 0 && est > floor(((double)FE_TWO_PI_F - a) / qi)) est = floor(((double)FE_TWO_PI_F - a) / qi);
          if (est > n - k) est = n - k;
          if (est < 1) est = 1;
          // cond(j): the j-th step is regular
          #define FE_COND(j) ((a + ((j) - 1) * qi + cm >= lo) && (a + ((j) - 1) * qi + cm < hi) && (a + (j) * qi <= (double)FE_TWO_PI_F))
          while (est > 1 && !FE_COND(est)) est -= 1;
          while (est < n - k && FE_COND(est + 1)) est += 1;
          #undef FE_COND
          jmax = est;
        }
        if (jmax >= 1) { len = (int)jmax; inc = sgn * qi; }
      }
    }
    float v_next;
    if (len == 0) {                                          // one step in real float arithmetic
      v_next = fe_wrap(fe_add(v, c));
      len = v_next == v ? n - k : 1;                         // a step that does not move v never will (c = 0, or |c| below half an ulp)
      inc = (double)v_next - (double)v;
    } else {
      v_next = (float)((double)v + len * inc);               // exact: a multiple of the ulp inside the binade
    }
    if (seg_k0) { seg_k0[ns] = k; seg_v0[ns] = v; seg_inc[ns] = inc; }
    ++ns;
    k += len;
    v = v_next;
  }
  if (n_seg) *n_seg = ns;
  if (n_done) *n_done = k;
  return v;
}

// ---- pass 0: weighted sum of every input tile for the DC average ------------------------------------------------------
FE_FN void fe_dc_partial_body(const FeArgs& A, int s, int t)
{
  FE_SHARED double sh_re[FE_THREADS], sh_im[FE_THREADS];
  const FeChunk ck = A.chunk[s];
  const int i0 = t * FE_TILE_IN;
  if (i0 >= ck.len_in) return;
  const int L = ck.len_in - i0 < FE_TILE_IN ? ck.len_in - i0 : FE_TILE_IN;
  const int16_t* pi = A.i_in + s * A.in_stride;
  const int16_t* pq = A.q_in + s * A.in_stride;
  FE_FOR(k, FE_THREADS) {
    double re = 0.0, im = 0.0;
    for (int j = 4 * k; j < 4 * k + 4 && j < L; ++j) {
      const double w = FE_LDG(A.apow + (L - 1 - j));
      re += w * (double)fe_mul((float)pi[(long long)(i0 + j) * A.step], ck.short_to_float);
      im += w * (double)fe_mul((float)pq[(long long)(i0 + j) * A.step], ck.short_to_float);
    }
    sh_re[k] = re; sh_im[k] = im;
  }
  FE_SYNC();
  for (int step = FE_THREADS / 2; step > 0; step >>= 1) {
    FE_FOR(k, step) { sh_re[k] += sh_re[k + step]; sh_im[k] += sh_im[k + step]; }
    FE_SYNC();
  }
  FE_ONE() {
    double2 w; w.x = (double)FE_DC_RATIO * sh_re[0]; w.y = (double)FE_DC_RATIO * sh_im[0];
    A.dc_part[s * FE_MAX_TILES + t] = w;
  }
}

// ---- pass 1: per stream, the state at every tile boundary and the output counts (one thread) ---------------------------
FE_FN void fe_plan_body(const FeArgs& A, int s)
{
  FE_SHARED double2 sh_w[FE_MAX_TILES];                      // the tile sums, fetched by all threads: the walk below is serial
  FE_FOR(t, (A.chunk[s].len_in + FE_TILE_IN - 1) / FE_TILE_IN) sh_w[t] = A.dc_part[s * FE_MAX_TILES + t];
  FE_SYNC();
  FE_ONE() {
    const FeChunk ck = A.chunk[s];
    const FeStream st = A.cur[s];
    FePlan& P = A.plan[s];
    const int n = ck.len_in;
    const int nt = (n + FE_TILE_IN - 1) / FE_TILE_IN;
    P.n_tiles_in = nt;
    // DC average at the tile boundaries
    double yr = st.dc_re, yi = st.dc_im;
    for (int t = 0; t < nt; ++t) {
      double2 y0; y0.x = yr; y0.y = yi;
      P.dc_start[t] = y0;
      const int L = n - t * FE_TILE_IN < FE_TILE_IN ? n - t * FE_TILE_IN : FE_TILE_IN;
      const double a = FE_LDG(A.apow + L);
      const double2 w = sh_w[t];
      yr = a * yr + w.x; yi = a * yi + w.y;
    }
    P.dc_re_next = (float)yr; P.dc_im_next = (float)yi;
    // NCO: the whole chunk in linear segments; the tile boundaries read off them
    float v = st.frequency_nco;
    const float c = -ck.frequency_est_filtered;
    int ns = 0, nd = 0;
    const float v_end = fe_nco_run(v, c, n, FE_MAX_CHUNK_SEG, P.seg_k0, P.seg_v0, P.seg_inc, &ns, &nd);
    if (nd == n) {
      P.n_seg = ns;
      int g = 0;
      for (int t = 0; t < nt; ++t) {
        const int k = t * FE_TILE_IN;
        while (g + 1 < ns && P.seg_k0[g + 1] <= k) ++g;
        P.nco_start[t] = ns ? (float)((double)P.seg_v0[g] + (double)(k - P.seg_k0[g]) * P.seg_inc[g]) : v;
      }
      v = v_end;
    } else {                                                  // too many segments: tile by tile
      P.n_seg = 0;
      for (int t = 0; t < nt; ++t) {
        P.nco_start[t] = v;
        const int L = n - t * FE_TILE_IN < FE_TILE_IN ? n - t * FE_TILE_IN : FE_TILE_IN;
        v = fe_nco_run(v, c, L, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
      }
    }
    P.nco_next = v;
    // resampler: output m sits at X_m = x1 + m d and is made while input floor(X_m + 1/2) is the newest sample
    const double d = (double)ck.resample, x1 = (double)st.x1;
    double M = ceil(((double)n - 0.5 - x1) / d);
    if (M < 0) M = 0;
    while (M > 0 && !(x1 + (M - 1) * d < (double)n - 0.5)) M -= 1;      // output M-1 must belong to this chunk ...
    while (x1 + M * d < (double)n - 0.5) M += 1;                        // ... and output M to the next one
    P.n_interp = (int)M;
    P.x1_next = (float)(x1 + M * d - (double)n);
    // decimator: it emits when its counter reaches 2 (filter_decimator.h:91-93)
    const int m0 = st.parity == 1 ? 0 : 1;
    P.m0 = m0;
    P.n_out = (int)M > m0 ? ((int)M - m0 + 1) / 2 : 0;
    P.parity_next = (st.parity + (int)M) & 1;
  }
}

// ---- pass 2: int16 -> float, DC removal, IQ statistics and correction, NCO derotation; one CTA per input tile --------------
FE_FN void fe_derotate_body(const FeArgs& A, int s, int t)
{
  FE_SHARED float sh_phase[FE_TILE_IN];
  FE_SHARED double2 sh_z[FE_TILE_IN];                    // prefix of x_j (1-r)^-j inside the thread's group of four
  FE_SHARED double2 sh_scan[2][FE_THREADS];
  FE_SHARED double sh_th[3][FE_THREADS];
  FE_SHARED int sh_seg_k0[FE_MAX_CHUNK_SEG];
  FE_SHARED float sh_seg_v0[FE_MAX_CHUNK_SEG];
  FE_SHARED double sh_seg_inc[FE_MAX_CHUNK_SEG];
  FE_SHARED int sh_nseg, sh_ndone, sh_base, sh_g0;
  const FeChunk ck = A.chunk[s];
  const int i0 = t * FE_TILE_IN;
  if (t == 0) { FE_FOR(j, 3) A.derot[s * A.derot_stride + j] = A.cur[s].delay[2 - j]; }      // the resampler's delay line
  if (i0 >= ck.len_in) return;
  const int L = ck.len_in - i0 < FE_TILE_IN ? ck.len_in - i0 : FE_TILE_IN;
  const FePlan& P = A.plan[s];
  const int16_t* pi = A.i_in + s * A.in_stride;
  const int16_t* pq = A.q_in + s * A.in_stride;
  const float c = -ck.frequency_est_filtered;
  // phase A: the NCO segments that cover this tile -- the chunk's (fe_plan_stream) or, if the chunk has too many, the tile's own
  // (one thread); the local prefix sums of the DC average (all threads)
  if (P.n_seg > 0) {
    FE_FOR(g, P.n_seg) { sh_seg_k0[g] = P.seg_k0[g]; sh_seg_v0[g] = P.seg_v0[g]; sh_seg_inc[g] = P.seg_inc[g]; }
    FE_ONE() {
      sh_nseg = P.n_seg; sh_ndone = L; sh_base = i0;
      int lo = 0, hi = P.n_seg - 1;                              // the segment of the tile's first sample: k0 < i0 + 1
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (P.seg_k0[mid] < i0 + 1) lo = mid; else hi = mid - 1; }
      sh_g0 = lo;
    }
  } else {
    FE_ONE() {
      int ns = 0, nd = 0;
      float v = fe_nco_run(P.nco_start[t], c, L, FE_MAX_SEG, sh_seg_k0, sh_seg_v0, sh_seg_inc, &ns, &nd);
      for (int k = nd; k < L; ++k) { v = fe_wrap(fe_add(v, c)); sh_phase[k] = v; }      // beyond FE_MAX_SEG segments: step by step
      sh_nseg = ns; sh_ndone = nd; sh_base = 0; sh_g0 = 0;
    }
  }
  FE_FOR(k, FE_THREADS) {
    double2 acc; acc.x = 0.0; acc.y = 0.0;
    for (int j = 4 * k; j < 4 * k + 4 && j < L; ++j) {
      const double w = FE_LDG(A.ainv + j);
      acc.x += w * (double)fe_mul((float)pi[(long long)(i0 + j) * A.step], ck.short_to_float);
      acc.y += w * (double)fe_mul((float)pq[(long long)(i0 + j) * A.step], ck.short_to_float);
      sh_z[j] = acc;
    }
    sh_scan[0][k] = acc;
  }
  FE_SYNC();
  // phase B: inclusive scan of the 256 group sums (ping-pong), the phase of every sample from its segment
  int src = 0;
  for (int off = 1; off < FE_THREADS; off <<= 1) {
    FE_FOR(k, FE_THREADS) {
      double2 v = sh_scan[src][k];
      if (k >= off) { v.x += sh_scan[src][k - off].x; v.y += sh_scan[src][k - off].y; }
      sh_scan[src ^ 1][k] = v;
    }
    FE_SYNC();
    src ^= 1;
  }
  FE_FOR(i, L) {
    if (i < sh_ndone) {
      const int step_idx = sh_base + i + 1;                     // sample i uses v_{i+1}: the segment with k0 < i + 1 <= k0 + len
      int lo = sh_g0;                                           // the tile's first segment; a tile rarely spans more than two
      while (lo + 1 < sh_nseg && sh_seg_k0[lo + 1] < step_idx) ++lo;
      sh_phase[i] = (float)((double)sh_seg_v0[lo] + (double)(step_idx - sh_seg_k0[lo]) * sh_seg_inc[lo]);
    }
  }
  FE_SYNC();
  // phase C: the samples; thread k takes samples k, k + 256, ... (coalesced) and keeps its share of the statistics in registers
  const double2 y0 = P.dc_start[t];
  FE_FOR(k, FE_THREADS) {
    float th0 = 0.0f, th1 = 0.0f, th2 = 0.0f;                    // four samples per thread in float, the tree below in double
    for (int i = k; i < L; i += FE_THREADS) {
      const int g = i >> 2;
      double2 pre = sh_z[i];
      if (g > 0) { pre.x += sh_scan[src][g - 1].x; pre.y += sh_scan[src][g - 1].y; }
      const double a1 = FE_LDG(A.apow + i + 1), a0 = FE_LDG(A.apow + i);
      const float dc_re = (float)(a1 * y0.x + (double)FE_DC_RATIO * a0 * pre.x);
      const float dc_im = (float)(a1 * y0.y + (double)FE_DC_RATIO * a0 * pre.y);
      float real = fe_sub(fe_mul((float)pi[(long long)(i0 + i) * A.step], ck.short_to_float), dc_re);
      float imag = fe_sub(fe_mul((float)pq[(long long)(i0 + i) * A.step], ck.short_to_float), dc_im);
      const float sr = real < 0 ? -1.0f : 1.0f, si = imag < 0 ? -1.0f : 1.0f;
      th0 -= fe_mul(imag, sr);
      th1 += fe_mul(real, sr);
      th2 += fe_mul(imag, si);
      real = fe_mul(real, ck.c2);
      imag = fe_add(imag, fe_mul(ck.c1, real));
      const float off_nco = fe_wrap(fe_sub(sh_phase[i], ck.phase_nco));
      const int idx = (int)fe_add(fe_mul(off_nco, FE_K_TABLE), 32767.0f) & 65535;
      const float2 cs = FE_LDG(A.lut_cs + idx);
      float2 o;
      o.x = fe_sub(fe_mul(real, cs.x), fe_mul(imag, cs.y));
      o.y = fe_add(fe_mul(imag, cs.x), fe_mul(real, cs.y));
      A.derot[s * A.derot_stride + 3 + i0 + i] = o;
    }
    sh_th[0][k] = th0; sh_th[1][k] = th1; sh_th[2][k] = th2;
  }
  FE_SYNC();
  for (int step = FE_THREADS / 2; step > 0; step >>= 1) {
    FE_FOR(k, step) { sh_th[0][k] += sh_th[0][k + step]; sh_th[1][k] += sh_th[1][k + step]; sh_th[2][k] += sh_th[2][k + step]; }
    FE_SYNC();
  }
  FE_ONE() {
    double* tp = A.theta_part + ((long long)s * FE_MAX_TILES + t) * 3;
    tp[0] = sh_th[0][0]; tp[1] = sh_th[1][0]; tp[2] = sh_th[2][0];
  }
}

// ---- pass 3: Farrow resampler + half-band decimator; one CTA per 1 024 outputs, one more per stream commits the state ----
FE_FN float2 fe_derot_at(const FeArgs& A, int s, long long i)
{
  return A.derot[s * A.derot_stride + 3 + i];               // i >= -3: the row starts with the three samples before the chunk
}

FE_FN float2 fe_interp_at(const FeArgs& A, int s, int m, double x1, double d)
{
  if (m < 0) return A.cur[s].hist[FE_TAPS - 1 + m];
  const double X = x1 + m * d;
  const double fi = floor(X + 0.5);
  const long long i = (long long)fi;
  const float mu = (float)(X - fi);
  const float2 in = fe_derot_at(A, s, i), d1 = fe_derot_at(A, s, i - 1), d2 = fe_derot_at(A, s, i - 2), d3 = fe_derot_at(A, s, i - 3);
  const float x2 = fe_mul(mu, mu), x3 = fe_mul(x2, mu);
  float2 v;
  {
    const float even1 = fe_add(d3.x, in.x), even2 = fe_add(d2.x, d1.x), odd1 = fe_sub(d3.x, in.x), odd2 = fe_sub(d2.x, d1.x);
    const float a0 = fe_sub(fe_mul(0.5625f, even2), fe_mul(0.0625f, even1));
    const float a1 = fe_sub(fe_mul(0.125f, odd1), fe_mul(1.375f, odd2));
    const float a2 = fe_mul(0.25f, fe_sub(even1, even2));
    const float a3 = fe_sub(fe_mul(1.5f, odd2), fe_mul(0.5f, odd1));
    v.x = fe_add(fe_add(fe_add(fe_mul(a3, x3), fe_mul(a2, x2)), fe_mul(a1, mu)), a0);
  }
  {
    const float even1 = fe_add(d3.y, in.y), even2 = fe_add(d2.y, d1.y), odd1 = fe_sub(d3.y, in.y), odd2 = fe_sub(d2.y, d1.y);
    const float a0 = fe_sub(fe_mul(0.5625f, even2), fe_mul(0.0625f, even1));
    const float a1 = fe_sub(fe_mul(0.125f, odd1), fe_mul(1.375f, odd2));
    const float a2 = fe_mul(0.25f, fe_sub(even1, even2));
    const float a3 = fe_sub(fe_mul(1.5f, odd2), fe_mul(0.5f, odd1));
    v.y = fe_add(fe_add(fe_add(fe_mul(a3, x3), fe_mul(a2, x2)), fe_mul(a1, mu)), a0);
  }
  return v;
}

FE_FN void fe_resample_body(const FeArgs& A, int s, int tile, int n_tiles)
{
  FE_SHARED float4 sh_v4[(2 * FE_TILE_OUT + FE_TAPS) / 2 * 5 / 4 + 4];
  const FePlan& P = A.plan[s];
  const FeChunk ck = A.chunk[s];
  const double x1 = (double)A.cur[s].x1, d = (double)ck.resample;
  if (tile == n_tiles - 1) {                               // the committing CTA: state of the next chunk, results
    FE_FOR(h, FE_TAPS - 1) A.next[s].hist[h] = fe_interp_at(A, s, P.n_interp - (FE_TAPS - 1) + h, x1, d);
    FE_FOR(j, 3) A.next[s].delay[j] = fe_derot_at(A, s, (long long)ck.len_in - 1 - j);
    FE_ONE() {
      FeStream& N = A.next[s];
      N.dc_re = P.dc_re_next; N.dc_im = P.dc_im_next; N.frequency_nco = P.nco_next; N.x1 = P.x1_next; N.parity = P.parity_next;
      N.pad = 0;
      double th[3] = {0.0, 0.0, 0.0};
      for (int t = 0; t < P.n_tiles_in; ++t)
        for (int q = 0; q < 3; ++q) th[q] += A.theta_part[((long long)s * FE_MAX_TILES + t) * 3 + q];
      FeResult r;
      r.len_out = P.n_out; r.len_interp = P.n_interp; r.theta1 = (float)th[0]; r.theta2 = (float)th[1]; r.theta3 = (float)th[2];
      A.result[s] = r;
    }
    return;
  }
  const int k_lo = tile * FE_TILE_OUT;
  if (k_lo >= P.n_out) return;
  const int nk = P.n_out - k_lo < FE_TILE_OUT ? P.n_out - k_lo : FE_TILE_OUT;
  const int m_lo = P.m0 + 2 * k_lo - (FE_TAPS - 1);
  const int count = 2 * (nk - 1) + FE_TAPS;
  // resampler outputs of the tile into shared memory; 16 bytes of padding after every 64 (the decimator below reads 16-byte
  // words at a stride of 64 bytes per thread)
  float2* sh_v = reinterpret_cast<float2*>(sh_v4);
  FE_UNROLL4
  FE_FOR(j, count) sh_v[j + ((j >> 3) << 1)] = fe_interp_at(A, s, m_lo + j, x1, d);
  FE_SYNC();
  // filter_decimator.h:95-123: four 8-float lanes (one complex sample each), blocks of 16 samples, (m0 + m1) + (m2 + m3) per lane
  // and block, lanes added left to right at the end.  A thread makes FOUR consecutive outputs: their windows are 2 samples apart,
  // so one block of the four windows is 22 samples held in registers (eleven 16-byte loads instead of 4 x 16 8-byte ones).
  FE_FOR(g, (nk + 3) >> 2) {
    float lr[4][4], li[4][4];
    for (int o = 0; o < 4; ++o) for (int l = 0; l < 4; ++l) { lr[o][l] = 0.0f; li[o][l] = 0.0f; }
    #pragma unroll
    for (int b = 0; b < 4; ++b) {
      float2 r[22];
      #pragma unroll
      for (int q = 0; q < 11; ++q) {
        const int w4 = 4 * g + 8 * b + q;                       // 16-byte word of the unpadded window
        const float4 x = sh_v4[w4 + (w4 >> 2)];
        r[2 * q].x = x.x; r[2 * q].y = x.y; r[2 * q + 1].x = x.z; r[2 * q + 1].y = x.w;
      }
      #pragma unroll
      for (int o = 0; o < 4; ++o)
        #pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int t0 = 16 * b + l, j0 = 2 * o + l;
          lr[o][l] = fe_add(lr[o][l], fe_add(fe_add(fe_mul(r[j0].x, FE_H(A, t0)), fe_mul(r[j0 + 4].x, FE_H(A, t0 + 4))),
                                             fe_add(fe_mul(r[j0 + 8].x, FE_H(A, t0 + 8)), fe_mul(r[j0 + 12].x, FE_H(A, t0 + 12)))));
          li[o][l] = fe_add(li[o][l], fe_add(fe_add(fe_mul(r[j0].y, FE_H(A, t0)), fe_mul(r[j0 + 4].y, FE_H(A, t0 + 4))),
                                             fe_add(fe_mul(r[j0 + 8].y, FE_H(A, t0 + 8)), fe_mul(r[j0 + 12].y, FE_H(A, t0 + 12)))));
        }
    }
    #pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (4 * g + o < nk) {
        float2 v;
        v.x = fe_add(fe_add(fe_add(lr[o][0], lr[o][1]), lr[o][2]), lr[o][3]);
        v.y = fe_add(fe_add(fe_add(li[o][0], li[o][1]), li[o][2]), li[o][3]);
        A.out[s * A.out_stride + k_lo + 4 * g + o] = v;
      }
    }
  }
}

// ---- guard-interval correlation (dvbt2_demodulator.cpp:321-330): one CTA per symbol ------------------------------------
FE_FN float fe_atan2_approx(float y, float x)            // DSP/fast_math.h:62-80
{
  const float pi = 3.14159265358979323846f, pi_2 = 1.57079632679489661923f;
  if (x == 0.0f) return y > 0.0f ? pi_2 : -pi_2;
  if (y == 0.0f) return x > 0.0f ? 0.0f : -pi;
  const float ax = fabsf(x), ay = fabsf(y);
  const bool min_x = ax < ay;
  const float a = min_x ? ax / ay : ay / ax;
  const float q = fe_mul(a, a);
  float r = fe_add(fe_mul(fe_mul(fe_sub(fe_mul(fe_add(fe_mul(-4.6496475e-2f, q), 1.5931422e-1f), q), 3.2762276e-1f), q), a), a);
  if (min_x) r = fe_sub(pi_2, r);
  if (x < 0.0f) r = fe_sub(pi, r);
  if (y < 0.0f) r = -r;
  return r;
}

FE_FN void fe_cp_correlate_body(const float2* sym, int fft_size, int guard, float* frequency_est)
{
  FE_SHARED double sh_r[FE_THREADS], sh_i[FE_THREADS];
  FE_FOR(k, FE_THREADS) { sh_r[k] = 0.0; sh_i[k] = 0.0; }
  FE_SYNC();
  const int n = guard - 8;
  FE_FOR(j, n > 0 ? n : 0) {
    const float2 a = sym[fft_size + 4 + j], b = sym[4 + j];
    sh_r[j & (FE_THREADS - 1)] += (double)fe_add(fe_mul(a.x, b.x), fe_mul(a.y, b.y));      // a * conj(b)
    sh_i[j & (FE_THREADS - 1)] += (double)fe_sub(fe_mul(a.y, b.x), fe_mul(a.x, b.y));
  }
  FE_SYNC();
  for (int step = FE_THREADS / 2; step > 0; step >>= 1) {
    FE_FOR(k, step) { sh_r[k] += sh_r[k + step]; sh_i[k] += sh_i[k + step]; }
    FE_SYNC();
  }
  FE_ONE() { *frequency_est = fe_atan2_approx((float)sh_i[0], (float)sh_r[0]) / (float)(fft_size << 1); }
}

// ---- P1 correlator (p1_symbol.cpp:75-178; the chain block diagram at :56-74; DSP/buffers.hh) -----------------------------
// The reference pushes every sample through two delay lines, two running sums and two more delays.  In closed form, with
// s[i] = x[i] fq[(i0 + i) mod 1024] (the frequency shift, :30-35), pc[i] = x[i] conj(s[i - 542]), pb[i] = s[i] conj(x[i - 482]):
//   out[i] = (sum of pc over the 541 samples up to i - 964) * (sum of pb over the 481 samples up to i - 2),  correlation = |out|^2
// (sum_of_buffer<LEN> holds LEN - 1 terms: it subtracts the slot it will overwrite next, buffers.hh:33-39).  The window sums are
// differences of prefix sums kept in double; one CTA scans the block (P1 search runs on one symbol's worth of samples per frame).
enum { FE_P1_THREADS = 512, FE_P1_HISTORY = 2046, FE_P1_LEAD = 1505 };   // samples before the block that still matter; prefix origin

FE_FN float2 fe_p1_x(const float2* x, const float2* hist, int j)
{
  if (j >= 0) return x[j];
  if (hist && j >= -FE_P1_HISTORY) return hist[FE_P1_HISTORY + j];
  float2 z; z.x = 0.f; z.y = 0.f;
  return z;
}
FE_FN float2 fe_p1_shift(const float2* x, const float2* hist, const float2* fq, int i0, int j)
{
  const float2 d = fe_p1_x(x, hist, j), f = fq[(i0 + j) & 1023];
  float2 r;
  r.x = fe_sub(fe_mul(d.x, f.x), fe_mul(d.y, f.y));
  r.y = fe_add(fe_mul(d.x, f.y), fe_mul(d.y, f.x));
  return r;
}
// element j of the two product sequences (j >= -FE_P1_LEAD + 1)
FE_FN void fe_p1_terms(const float2* x, const float2* hist, const float2* fq, int i0, int j, float2& pc, float2& pb)
{
  const float2 d = fe_p1_x(x, hist, j), sh = fe_p1_shift(x, hist, fq, i0, j);
  const float2 c = fe_p1_shift(x, hist, fq, i0, j - 542), b = fe_p1_x(x, hist, j - 482);
  pc.x = fe_add(fe_mul(d.x, c.x), fe_mul(d.y, c.y)); pc.y = fe_sub(fe_mul(d.y, c.x), fe_mul(d.x, c.y));
  pb.x = fe_add(fe_mul(sh.x, b.x), fe_mul(sh.y, b.y)); pb.y = fe_sub(fe_mul(sh.y, b.x), fe_mul(sh.x, b.y));
}

// prefix: double2[2][n + FE_P1_LEAD + 1] scratch (inclusive prefix sums of pc / pb from j = -FE_P1_LEAD + 1; entry 0 is zero)
FE_FN void fe_p1_correlate_body(const float2* x, int n, const float2* hist, const float2* fq, int i0, double2* prefix,
                                float* correlation, float2* out)
{
  FE_SHARED double2 sh_c[2][FE_P1_THREADS], sh_b[2][FE_P1_THREADS];
  const int total = n + FE_P1_LEAD;                          // elements j = -FE_P1_LEAD + 1 .. n - 1, stored at j + FE_P1_LEAD
  const int per = (total + FE_P1_THREADS - 1) / FE_P1_THREADS;
  double2* Pc = prefix; double2* Pb = prefix + (total + 1);
  FE_FOR(t, FE_P1_THREADS) {
    double2 ac, ab; ac.x = ac.y = ab.x = ab.y = 0.0;
    for (int e = t * per + 1; e <= (t + 1) * per && e <= total; ++e) {
      float2 pc, pb;
      fe_p1_terms(x, hist, fq, i0, e - FE_P1_LEAD, pc, pb);
      ac.x += pc.x; ac.y += pc.y; ab.x += pb.x; ab.y += pb.y;
    }
    sh_c[0][t] = ac; sh_b[0][t] = ab;
  }
  FE_SYNC();
  int src = 0;
  for (int off = 1; off < FE_P1_THREADS; off <<= 1) {
    FE_FOR(t, FE_P1_THREADS) {
      double2 vc = sh_c[src][t], vb = sh_b[src][t];
      if (t >= off) { vc.x += sh_c[src][t - off].x; vc.y += sh_c[src][t - off].y; vb.x += sh_b[src][t - off].x; vb.y += sh_b[src][t - off].y; }
      sh_c[src ^ 1][t] = vc; sh_b[src ^ 1][t] = vb;
    }
    FE_SYNC();
    src ^= 1;
  }
  FE_FOR(t, FE_P1_THREADS) {
    double2 ac, ab; ac.x = ac.y = ab.x = ab.y = 0.0;
    if (t > 0) { ac = sh_c[src][t - 1]; ab = sh_b[src][t - 1]; }
    if (t == 0) { Pc[0] = ac; Pb[0] = ab; }
    for (int e = t * per + 1; e <= (t + 1) * per && e <= total; ++e) {
      float2 pc, pb;
      fe_p1_terms(x, hist, fq, i0, e - FE_P1_LEAD, pc, pb);
      ac.x += pc.x; ac.y += pc.y; ab.x += pb.x; ab.y += pb.y;
      Pc[e] = ac; Pb[e] = ab;
    }
  }
  FE_SYNC();
  FE_FOR(i, n) {
    const int ec = i - 964 + FE_P1_LEAD, eb = i - 2 + FE_P1_LEAD;      // prefix entries of the window ends
    float2 a, d;
    a.x = (float)(Pc[ec].x - Pc[ec - 541].x); a.y = (float)(Pc[ec].y - Pc[ec - 541].y);
    d.x = (float)(Pb[eb].x - Pb[eb - 481].x); d.y = (float)(Pb[eb].y - Pb[eb - 481].y);
    float2 o;
    o.x = fe_sub(fe_mul(a.x, d.x), fe_mul(a.y, d.y));
    o.y = fe_add(fe_mul(a.x, d.y), fe_mul(a.y, d.x));
    correlation[i] = fe_add(fe_mul(o.x, o.x), fe_mul(o.y, o.y));
    if (out) out[i] = o;
  }
}
