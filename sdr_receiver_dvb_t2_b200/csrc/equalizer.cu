// K2: pilot-based channel estimation, per-carrier equalisation and frequency de-interleaving for P2,
// data and frame-closing symbols.
//
// Reference semantics reproduced (paths relative to the reference's src/):
//   DVB_T2/data_symbol.cpp:108-335, fc_symbol.cpp:82-271, p2_symbol.cpp:94-259
//     every estimating pilot gives angle = atan2_approx(cell * ref) and amp = |cell| / amp_pilot; the data
//     cells between two consecutive pilots get angle / amp interpolated linearly BY REPEATED ADDITION of the
//     step (so the n-th cell carries n rounded float adds), with the asymmetric +-pi unwrap of
//     data_symbol.cpp:190-191; derotation through the 65 536-entry sin/cos table (DSP/fast_math.h:25-42);
//     out[h[d]] = cell * conj(e^{j angle} / amp); the centre carrier never estimates; continual pilots use
//     amp_cp, scattered / edge ones amp_sp; phase = atan2(sum pilots left) + atan2(sum pilots right),
//     sro = sum angle right - sum angle left, both accumulated in carrier order.
// B200 design: the serial scan of the reference is turned inside out.  At table-upload time the host
// compiles each symbol's carrier map into a PLAN: the list of estimating pilots and, per data cell, its
// carrier, its left pilot, its position inside the interval and its de-interleaved address (8 bytes per
// cell, identical plans shared between symbols), plus the first-cell index of every pilot interval.  Eight
// CTAs share a symbol (1 024-cell chunks dealt round-robin; see equalize_kernel): pilots are estimated in
// parallel into shared memory, the interpolation chains run once per interval, the cells of a chunk are
// equalised independently and leave through the de-interleaver as 8-byte scattered stores that merge in L2
// because few symbols are in flight.  Arithmetic is written with explicit round-to-nearest intrinsics so no
// FMA contraction changes a bit relative to the CPU oracle.
#include "stages.h"
#include <cuda_pipeline.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace {

enum { DATA_CARRIER = 1, P2CARRIER, P2PAPR_CARRIER, TRPAPR_CARRIER, SCATTERED_CARRIER, CONTINUAL_CARRIER,
       P2CARRIER_INVERTED, SCATTERED_CARRIER_INVERTED, CONTINUAL_CARRIER_INVERTED };     // dvbt2_definition.h:103-113

struct PlanPilot { uint16_t k; uint8_t second_half; uint8_t pad; float ref; float amp; };
struct PlanHost {
  std::vector<PlanPilot> pilots;
  std::vector<int> first;            // per pilot: index of the first data cell to its right
  std::vector<uint2> cells;          // x = k | left_pilot << 16 ; y = j | n << 8 | h << 16
};
struct PlanDev { PlanPilot* pilots = nullptr; uint2* cells = nullptr; int* first = nullptr; int* chunk_lo = nullptr; int* chunk_k = nullptr; int n_pilots = 0, n_cells = 0, n_first = 0, pad = 0; };
// chunk_lo[c] = interval (left pilot) holding data cell c * kEqChunk; chunk_k[2c], [2c+1] = first / last carrier of chunk c
// first[ip] = index in cells[] of the first data cell right of pilot ip (first[n_pilots - 1] = n_cells)

}  // namespace

struct SymbolTables {
  int kind = 0, k_total = 0, l_nulls = 0, fft_size = 0, n_out = 0, first_symbol = 0, n_symbols = 0;
  std::vector<PlanDev> plans;            // unique plans
  std::vector<int> plan_even, plan_odd;  // per symbol of the kind: plan when idx_symbol is even / odd
  int* d_plan_even = nullptr; int* d_plan_odd = nullptr;
  PlanDev* d_plans = nullptr;
  int max_pilots = 0;
};

namespace {

__device__ __forceinline__ float atan2_approx_dev(float y, float x)      // DSP/fast_math.h:61-81
{
  const float PI = 3.14159265358979323846f, PI_2 = 1.57079632679489661923f;
  if (x == 0.0f) return y > 0.0f ? PI_2 : -PI_2;
  if (y == 0.0f) return x > 0.0f ? 0.0f : -PI;
  const float ax = fabsf(x), ay = fabsf(y);
  const bool min_x = ax < ay;
  const float a = min_x ? __fdiv_rn(ax, ay) : __fdiv_rn(ay, ax);
  const float s = __fmul_rn(a, a);
  float r = __fadd_rn(__fmul_rn(-4.6496475e-2f, s), 1.5931422e-1f);
  r = __fsub_rn(__fmul_rn(r, s), 3.2762276e-1f);
  r = __fadd_rn(__fmul_rn(__fmul_rn(r, s), a), a);
  if (min_x) r = __fsub_rn(PI_2, r);
  if (x < 0.0f) r = __fsub_rn(PI, r);
  if (y < 0.0f) r = -r;
  return r;
}

// Symbol tables of one kind as a launch sees them: symbols r0 .. of a frame are of this kind, their cells go to
// out + f * out_frame + out_off + (r - r0) * out_sym
struct EqKind {
  const PlanDev* plans; const int* plan_even; const int* plan_odd;
  int first_symbol, n_symbols_kind, r0, pad;
  long long out_off, out_sym;
};
struct EqParams {
  const float2* freq; float2* out; float* sro; float* phase; const int* idx_symbol;
  const float2* lut_cs;           // {cos, sin} of DSP/fast_math.h's tables, interleaved
  int l_nulls;
  int n_symbols;                 // symbols in this launch
  int split;                     // CTAs per symbol
  // symbol s of the launch = symbol r = s % per_frame of frame f = s / per_frame (a plain batch is one "frame"):
  //   spectrum at freq + f * in_frame + r * in_sym, feedback floats at [f * fb_frame + r]; idx_symbol == nullptr means
  //   idx = first_symbol of its kind + (r - r0)
  int per_frame;
  long long in_frame, in_sym, out_frame, fb_frame;
  int n_kinds;                   // 1: a batch of one kind; 3 (2 without a frame-closing symbol): whole frames, P2 | data | FC
  EqKind kind[3];
};

constexpr int kEqThreads = 256;
constexpr int kEqChunk = 1024;          // data cells whose interpolated (angle, amplitude) are staged in shared memory
constexpr int kEqSpan = 2 * kEqChunk;   // carriers of the spectrum staged next to them (pilots and reserved tones between the cells)
constexpr int kEqU = kEqChunk / kEqThreads;
constexpr int kEqSplitDefault = 8;             // CTAs sharing one symbol (chunks dealt round-robin): keeps few symbols in flight so that the
                                        // scattered 8-byte stores of a symbol meet in L2 before their sectors are evicted

// One CTA per symbol.  Pilots are estimated in parallel into shared memory.  The data cells are then handled in
// chunks of 4096 (carrier order): phase A -- one THREAD per pilot interval runs the reference's repeated float
// additions of the interpolation step ONCE per interval (the n-th cell carries n rounded adds; a chain is
// serial, the ~45 chains of a chunk are not) and leaves (angle, amplitude) per cell in shared memory; phase B --
// all threads equalise the chunk's cells independently, fully coalesced, and send them through the frequency
// de-interleaver as 8-byte scattered stores that merge in L2.  The last thread forms the ordered pilot sums.
__global__ void __launch_bounds__(kEqThreads) equalize_kernel(const EqParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const float PI = 3.14159265358979323846f;
  const float k_table = 32767.0f / (2.0f * PI);
  float2* chain = reinterpret_cast<float2*>(smem_raw);
  const int part = blockIdx.x % p.split;
  for (int s = blockIdx.x / p.split; s < p.n_symbols; s += gridDim.x / p.split) {
    const int fr = s / p.per_frame, rs = s - fr * p.per_frame;
    const int kd = (p.n_kinds > 1 && rs >= p.kind[1].r0) ? ((p.n_kinds > 2 && rs >= p.kind[2].r0) ? 2 : 1) : 0;
    const EqKind& K = p.kind[kd];
    const int rk = rs - K.r0;
    const int idx = p.idx_symbol ? p.idx_symbol[s] : K.first_symbol + rk;
    int rel = idx - K.first_symbol;
    rel = min(max(rel, 0), K.n_symbols_kind - 1);
    const PlanDev pl = K.plans[(idx & 1) ? K.plan_odd[rel] : K.plan_even[rel]];
    float2* spec = chain + kEqChunk;
    float2* est = spec + kEqSpan;
    float* ang = reinterpret_cast<float*>(est + pl.n_pilots);
    float* amp = ang + pl.n_pilots;
    int* first = reinterpret_cast<int*>(amp + pl.n_pilots);
    const float2* cell = p.freq + (size_t)fr * p.in_frame + (size_t)rs * p.in_sym + p.l_nulls;
    float2* out = p.out + (size_t)fr * p.out_frame + (size_t)K.out_off + (size_t)rk * K.out_sym;
    const size_t fb = (size_t)fr * p.fb_frame + rs;

    for (int i = threadIdx.x; i < pl.n_pilots; i += blockDim.x) {
      const PlanPilot pp = pl.pilots[i];
      const float2 c = __ldg(cell + pp.k);
      const float er = __fmul_rn(c.x, pp.ref), ei = __fmul_rn(c.y, pp.ref);        // cell * pilot_refer
      est[i] = make_float2(er, ei);
      ang[i] = atan2_approx_dev(ei, er);
      amp[i] = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y))), pp.amp);
      first[i] = __ldg(pl.first + i);
    }
    __syncthreads();

    for (int c0 = part * kEqChunk; c0 < pl.n_cells; c0 += p.split * kEqChunk) {
      const int c1 = min(c0 + kEqChunk, pl.n_cells);
      // Loads that depend on nothing go first and stay in flight under phase A: the carriers spanned by the chunk's
      // cells (cp.async straight into shared memory) and the plan words of this thread's cells.
      const int ci = c0 / kEqChunk;
      const int k_lo = __ldg(pl.chunk_k + 2 * ci), k_hi = __ldg(pl.chunk_k + 2 * ci + 1);
      const bool staged = k_hi - k_lo < kEqSpan;
      if (staged)
        for (int k = k_lo + threadIdx.x; k <= k_hi; k += blockDim.x)
          __pipeline_memcpy_async(spec + (k - k_lo), cell + k, sizeof(float2));
      __pipeline_commit();
      uint2 w[kEqU];
#pragma unroll
      for (int u = 0; u < kEqU; ++u) {
        const int d = c0 + threadIdx.x + u * kEqThreads;
        w[u] = d < c1 ? __ldg(pl.cells + d) : make_uint2(0u, 0u);
      }
      // ---- phase A: one thread per interval that overlaps the chunk (shared memory only) ----
      for (int ip = __ldg(pl.chunk_lo + ci) + threadIdx.x; ip + 1 < pl.n_pilots; ip += blockDim.x) {
        const int d0 = first[ip], d1 = first[ip + 1];
        if (d0 >= c1) break;
        if (d0 >= d1) continue;
        const float ang_l = ang[ip], ang_r = ang[ip + 1], amp_l = amp[ip], amp_r = amp[ip + 1];
        float dif = __fsub_rn(ang_r, ang_l);
        if (dif > PI) dif = __fsub_rn(__fmul_rn(PI, 2.0f), dif);                      // data_symbol.cpp:190-191
        else if (dif < -PI) dif = __fadd_rn(__fmul_rn(PI, 2.0f), dif);
        const float fn = (float)(d1 - d0 + 1);
        const float da = __fdiv_rn(dif, fn), dm = __fdiv_rn(__fsub_rn(amp_r, amp_l), fn);
        float a = ang_l, m = amp_l;
        const int dend = min(d1, c1);
        for (int d = d0; d < dend; ++d) {
          a = __fadd_rn(a, da); m = __fadd_rn(m, dm);
          if (d >= c0) chain[d - c0] = make_float2(a, m);
        }
      }
      __pipeline_wait_prior(0);
      __syncthreads();
      // ---- phase B: all cells of the chunk, independently; the table look-ups of a thread's cells are in flight together ----
      float2 am[kEqU], cs[kEqU];
#pragma unroll
      for (int u = 0; u < kEqU; ++u) {
        const int d = c0 + threadIdx.x + u * kEqThreads;
        if (d < c1) {
          am[u] = chain[d - c0];
          const int li = __float2int_rz(__fadd_rn(__fmul_rn(am[u].x, k_table), 32767.0f)) & 65535;
          cs[u] = __ldg(p.lut_cs + li);
        }
      }
#pragma unroll
      for (int u = 0; u < kEqU; ++u) {
        const int d = c0 + threadIdx.x + u * kEqThreads;
        if (d < c1) {
          const int k = (int)(w[u].x & 0xffff);
          const float2 c = staged ? spec[k - k_lo] : __ldg(cell + k);
          const float dr = __fdiv_rn(cs[u].x, am[u].y), di = __fdiv_rn(cs[u].y, am[u].y);
          out[w[u].y >> 16] = make_float2(__fadd_rn(__fmul_rn(c.x, dr), __fmul_rn(c.y, di)),
                                          __fsub_rn(__fmul_rn(c.y, dr), __fmul_rn(c.x, di)));
        }
      }
      __syncthreads();
    }

    if (part == 0 && threadIdx.x == blockDim.x - 1 && (p.phase || p.sro)) {
      // ordered sums (data_symbol.cpp:162-163,183-193,319-324): the first pilot feeds sum_pilot_1 only.
      // Operands are fetched eight at a time so that only the three FADD chains are serial.
      float s1r = est[0].x, s1i = est[0].y, s2r = 0.f, s2i = 0.f, a1 = 0.f, a2 = 0.f;
      int i = 1;
      for (; i + 8 <= pl.n_first; i += 8) {
        float2 e[8]; float g[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { e[u] = est[i + u]; g[u] = ang[i + u]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s1r = __fadd_rn(s1r, e[u].x); s1i = __fadd_rn(s1i, e[u].y); a1 = __fadd_rn(a1, g[u]); }
      }
      for (; i < pl.n_first; ++i) {                                  // pilots left of the centre carrier
        const float2 e = est[i];
        s1r = __fadd_rn(s1r, e.x); s1i = __fadd_rn(s1i, e.y); a1 = __fadd_rn(a1, ang[i]);
      }
      for (; i + 8 <= pl.n_pilots; i += 8) {
        float2 e[8]; float g[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { e[u] = est[i + u]; g[u] = ang[i + u]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s2r = __fadd_rn(s2r, e[u].x); s2i = __fadd_rn(s2i, e[u].y); a2 = __fadd_rn(a2, g[u]); }
      }
      for (; i < pl.n_pilots; ++i) {                                 // right of it
        const float2 e = est[i];
        s2r = __fadd_rn(s2r, e.x); s2i = __fadd_rn(s2i, e.y); a2 = __fadd_rn(a2, ang[i]);
      }
      if (p.phase) p.phase[fb] = __fadd_rn(atan2_approx_dev(s2i, s2r), atan2_approx_dev(s1i, s1r));
      if (p.sro) p.sro[fb] = __fsub_rn(a2, a1);
    }
    __syncthreads();                        // the pilot arrays are reused by the next symbol
  }
}

// carrier map + pilot references of ONE symbol -> plan (the reference's scan, run once on the host)
bool build_plan(int kind, int k_total, int n_out, const int32_t* map, const float* refer, const int32_t* h, float amp_main,
                float amp_cp, PlanHost& out, std::string& err)
{
  const int half_total = k_total / 2;
  {
    // the de-interleaver tables hold n_out addresses: count the data cells before indexing them
    int cells = 0;
    for (int i = 1; i < k_total; ++i) cells += map[i] == DATA_CARRIER && !(i == half_total && kind == 0);
    if (cells > n_out) { err = "carrier map holds more data cells than n_out"; return false; }
  }
  out.pilots.clear(); out.cells.clear(); out.first.clear();
  out.pilots.push_back({0, 0, 0, refer[0], amp_main});            // carrier 0: always the first (edge) pilot
  out.first.push_back(0);
  std::vector<int> pending;                                        // carriers of buffered data cells
  int d = 0;
  for (int i = 1; i < k_total; ++i) {
    const int t = map[i];
    bool pilot;
    if (kind == 0) pilot = (t == P2CARRIER || t == P2CARRIER_INVERTED);
    else if (kind == 1) pilot = (t == SCATTERED_CARRIER || t == SCATTERED_CARRIER_INVERTED || t == CONTINUAL_CARRIER ||
                                 t == CONTINUAL_CARRIER_INVERTED);
    else pilot = (t == SCATTERED_CARRIER || t == SCATTERED_CARRIER_INVERTED);
    if (i == half_total) {                                         // centre carrier never estimates
      if (kind != 0 && t == DATA_CARRIER) pending.push_back(i);
      continue;
    }
    if (t == DATA_CARRIER) { pending.push_back(i); continue; }
    if (!pilot) continue;
    const bool cp = kind == 1 && (t == CONTINUAL_CARRIER || t == CONTINUAL_CARRIER_INVERTED);
    const int left = (int)out.pilots.size() - 1;
    out.pilots.push_back({(uint16_t)i, (uint8_t)(i > half_total), 0, refer[i], cp ? amp_cp : amp_main});
    const int n = (int)pending.size();
    if (n > 255) { err = "more than 255 data cells between two pilots"; return false; }
    for (int j = 0; j < n; ++j, ++d) {
      if (h[d] < 0 || h[d] >= n_out) { err = "de-interleaver address out of range"; return false; }
      out.cells.push_back(make_uint2((uint32_t)pending[j] | ((uint32_t)left << 16),
                                     (uint32_t)(j + 1) | ((uint32_t)n << 8) | ((uint32_t)h[d] << 16)));
    }
    pending.clear();
    out.first.push_back((int)out.cells.size());        // cells right of the pilot just added start here
  }
  if (out.pilots.size() > 65535) { err = "too many pilots"; return false; }
  return true;
}

void free_tables(SymbolTables* t)
{
  if (!t) return;
  for (auto& p : t->plans) { cudaFree(p.pilots); cudaFree(p.cells); cudaFree(p.first); cudaFree(p.chunk_lo); cudaFree(p.chunk_k); }
  cudaFree(t->d_plan_even); cudaFree(t->d_plan_odd); cudaFree(t->d_plans);
  delete t;
}

}  // namespace

bool t2_eq_geometry(const t2b200_ctx* ctx, int kind, int* fft_size, int* n_out, int* n_symbols, int* first_symbol)
{
  const SymbolTables* t = (kind >= 0 && kind < 3) ? ctx->sym[kind] : nullptr;
  if (!t) return false;
  *fft_size = t->fft_size; *n_out = t->n_out; *n_symbols = t->n_symbols; *first_symbol = t->first_symbol;
  return true;
}

void t2_eq_free(t2b200_ctx* ctx)
{
  for (auto& s : ctx->sym) { free_tables(s); s = nullptr; }
  if (ctx->d_lut) { cudaFree(ctx->d_lut); ctx->d_lut = nullptr; }
}

int t2_ensure_lut(t2b200_ctx* ctx)
{
  if (ctx->d_lut) return T2B200_OK;
  std::vector<float> h(2 * 65536, 0.0f);                          // DSP/fast_math.h:25-40: entry 65535 stays 0; {cos, sin} pairs
  const float k_table = 32767.0f / (2.0f * 3.14159265358979323846f);
  for (int i = -32767; i < 32768; i++) { h[2 * (i + 32767) + 1] = sinf(i / k_table); h[2 * (i + 32767)] = cosf(i / k_table); }
  T2_CUDA(ctx, cudaMalloc(&ctx->d_lut, h.size() * sizeof(float)));
  T2_CUDA(ctx, cudaMemcpy(ctx->d_lut, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return T2B200_OK;
}

extern "C" int t2b200_eq_configure(t2b200_ctx* ctx, int kind, int n_symbols, int first_symbol, int fft_size, int k_total,
                                   int l_nulls, int n_out, const int32_t* carrier_map, const float* pilot_refer,
                                   const int32_t* h_even, const int32_t* h_odd, float amp_main, float amp_cp)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (kind < 0 || kind > 2 || n_symbols <= 0 || !carrier_map || !pilot_refer || !h_even || !h_odd || k_total <= 0 ||
      k_total > 32768 || l_nulls < 0 || l_nulls + k_total > fft_size || n_out <= 0) {
    ctx->err = "t2b200_eq_configure: bad argument"; return T2B200_ERR_ARG;
  }
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = t2_ensure_lut(ctx))) return rc;
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  free_tables(ctx->sym[kind]); ctx->sym[kind] = nullptr;
  SymbolTables* t = new SymbolTables();
  t->kind = kind; t->k_total = k_total; t->l_nulls = l_nulls; t->fft_size = fft_size; t->n_out = n_out;
  t->first_symbol = first_symbol; t->n_symbols = n_symbols;
  std::vector<PlanHost> uniq;
  auto intern = [&](PlanHost& ph) -> int {
    for (size_t u = 0; u < uniq.size(); ++u)
      if (uniq[u].pilots.size() == ph.pilots.size() && uniq[u].cells.size() == ph.cells.size() &&
          !memcmp(uniq[u].pilots.data(), ph.pilots.data(), ph.pilots.size() * sizeof(PlanPilot)) &&
          !memcmp(uniq[u].cells.data(), ph.cells.data(), ph.cells.size() * sizeof(uint2))) return (int)u;
    uniq.push_back(ph);
    return (int)uniq.size() - 1;
  };
  for (int s = 0; s < n_symbols; ++s) {
    for (int parity = 0; parity < 2; ++parity) {
      PlanHost ph;
      // idx_symbol even -> h_odd, odd -> h_even (data_symbol.cpp:148-149)
      if (!build_plan(kind, k_total, n_out, carrier_map + (size_t)s * k_total, pilot_refer + (size_t)s * k_total,
                      parity ? h_even : h_odd, amp_main, amp_cp, ph, ctx->err)) { delete t; return T2B200_ERR_ARG; }
      if ((int)ph.cells.size() > n_out) { ctx->err = "carrier map holds more data cells than n_out"; delete t; return T2B200_ERR_ARG; }
      (parity ? t->plan_odd : t->plan_even).push_back(intern(ph));
    }
  }
  for (auto& ph : uniq) {
    PlanDev d;
    d.n_pilots = (int)ph.pilots.size(); d.n_cells = (int)ph.cells.size();
    d.n_first = 0;
    for (auto& pp : ph.pilots) d.n_first += pp.second_half ? 0 : 1;
    T2_CUDA(ctx, cudaMalloc(&d.pilots, ph.pilots.size() * sizeof(PlanPilot)));
    T2_CUDA(ctx, cudaMalloc(&d.cells, std::max<size_t>(1, ph.cells.size()) * sizeof(uint2)));
    T2_CUDA(ctx, cudaMemcpy(d.pilots, ph.pilots.data(), ph.pilots.size() * sizeof(PlanPilot), cudaMemcpyHostToDevice));
    T2_CUDA(ctx, cudaMemcpy(d.cells, ph.cells.data(), ph.cells.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    T2_CUDA(ctx, cudaMalloc(&d.first, ph.first.size() * sizeof(int)));
    T2_CUDA(ctx, cudaMemcpy(d.first, ph.first.data(), ph.first.size() * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<int> clo, ck;
    for (int c0 = 0, ip = 0; c0 < (int)ph.cells.size() || clo.empty(); c0 += kEqChunk) {
      while (ip + 1 < (int)ph.first.size() && ph.first[ip + 1] <= c0) ++ip;       // largest ip with first[ip] <= c0
      clo.push_back(ip);
      const int c1 = std::min<int>(c0 + kEqChunk, (int)ph.cells.size());
      ck.push_back(ph.cells.empty() ? 0 : (int)(ph.cells[c0].x & 0xffff));
      ck.push_back(ph.cells.empty() ? 0 : (int)(ph.cells[c1 - 1].x & 0xffff));
    }
    T2_CUDA(ctx, cudaMalloc(&d.chunk_k, ck.size() * sizeof(int)));
    T2_CUDA(ctx, cudaMemcpy(d.chunk_k, ck.data(), ck.size() * sizeof(int), cudaMemcpyHostToDevice));
    T2_CUDA(ctx, cudaMalloc(&d.chunk_lo, clo.size() * sizeof(int)));
    T2_CUDA(ctx, cudaMemcpy(d.chunk_lo, clo.data(), clo.size() * sizeof(int), cudaMemcpyHostToDevice));
    t->plans.push_back(d);
    t->max_pilots = std::max(t->max_pilots, d.n_pilots);
  }
  T2_CUDA(ctx, cudaMalloc(&t->d_plans, t->plans.size() * sizeof(PlanDev)));
  T2_CUDA(ctx, cudaMemcpy(t->d_plans, t->plans.data(), t->plans.size() * sizeof(PlanDev), cudaMemcpyHostToDevice));
  T2_CUDA(ctx, cudaMalloc(&t->d_plan_even, n_symbols * sizeof(int)));
  T2_CUDA(ctx, cudaMalloc(&t->d_plan_odd, n_symbols * sizeof(int)));
  T2_CUDA(ctx, cudaMemcpy(t->d_plan_even, t->plan_even.data(), n_symbols * sizeof(int), cudaMemcpyHostToDevice));
  T2_CUDA(ctx, cudaMemcpy(t->d_plan_odd, t->plan_odd.data(), n_symbols * sizeof(int), cudaMemcpyHostToDevice));
  ctx->sym[kind] = t;
  return T2B200_OK;
}

static void eq_fill_kind(EqKind& k, const SymbolTables* t, int r0, long long out_off, long long out_sym)
{
  k.plans = t->d_plans; k.plan_even = t->d_plan_even; k.plan_odd = t->d_plan_odd;
  k.first_symbol = t->first_symbol; k.n_symbols_kind = t->n_symbols; k.r0 = r0; k.pad = 0;
  k.out_off = out_off; k.out_sym = out_sym;
}

static int eq_launch(t2b200_ctx* ctx, EqParams& p, int max_pilots)
{
  p.lut_cs = reinterpret_cast<const float2*>(ctx->d_lut);
  const size_t smem = (size_t)(kEqChunk + kEqSpan) * sizeof(float2) + (size_t)(max_pilots + 2) * 20;
  T2_CUDA(ctx, cudaFuncSetAttribute(equalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.split = kEqSplitDefault;
  if (const char* e = getenv("T2B200_EQ_SPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 64) p.split = v; }   // development aid
  equalize_kernel<<<p.n_symbols * p.split, kEqThreads, smem, ctx->stream>>>(p);
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return T2B200_OK;
}

// Device-level launch (all pointers device memory): n_symbols symbols of one kind, per_frame of them per frame (see EqParams).
int t2_equalize_device(t2b200_ctx* ctx, int kind, int n_symbols, int per_frame, const int* d_idx, const float2* d_freq,
                       long long in_frame, long long in_sym, float2* d_out, long long out_frame, long long out_sym,
                       float* d_sro, float* d_phase, long long fb_frame)
{
  SymbolTables* t = ctx->sym[kind];
  if (!t) { ctx->err = "equalise: symbol tables not configured"; return T2B200_ERR_STATE; }
  if (n_symbols == 0) return T2B200_OK;
  EqParams p;
  memset(&p, 0, sizeof(p));
  p.freq = d_freq; p.out = d_out; p.sro = d_sro; p.phase = d_phase; p.idx_symbol = d_idx;
  p.l_nulls = t->l_nulls;
  p.n_symbols = n_symbols; p.per_frame = per_frame;
  p.in_frame = in_frame; p.in_sym = in_sym; p.out_frame = out_frame; p.fb_frame = fb_frame;
  p.n_kinds = 1;
  eq_fill_kind(p.kind[0], t, 0, 0, out_sym);
  return eq_launch(ctx, p, t->max_pilots);
}

extern "C" int t2b200_equalize(t2b200_ctx* ctx, int kind, int n_symbols, const int32_t* idx_symbol, const float* freq,
                               float* cells_out, float* sro, float* phase)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (kind < 0 || kind > 2 || n_symbols < 0 || !idx_symbol || !freq || !cells_out) { ctx->err = "t2b200_equalize: bad argument"; return T2B200_ERR_ARG; }
  SymbolTables* t = ctx->sym[kind];
  if (!t) { ctx->err = "t2b200_equalize: symbol tables not configured"; return T2B200_ERR_STATE; }
  if (n_symbols == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc; const void *dfreq, *didx; void *dout, *dsro = nullptr, *dph = nullptr;
  if ((rc = t2_to_device(ctx, 0, freq, (size_t)n_symbols * t->fft_size * 8, &dfreq))) return rc;
  if ((rc = t2_to_device(ctx, 5, idx_symbol, (size_t)n_symbols * 4, &didx))) return rc;
  if ((rc = t2_out_device(ctx, 1, cells_out, (size_t)n_symbols * t->n_out * 8, &dout))) return rc;
  if (sro && (rc = t2_out_device(ctx, 2, sro, (size_t)n_symbols * 4, &dsro))) return rc;
  if (phase && (rc = t2_out_device(ctx, 3, phase, (size_t)n_symbols * 4, &dph))) return rc;
  if ((rc = t2_equalize_device(ctx, kind, n_symbols, n_symbols, (const int*)didx, (const float2*)dfreq, 0, t->fft_size,
                               (float2*)dout, 0, t->n_out, (float*)dsro, (float*)dph, 0))) return rc;
  if ((rc = t2_finish_out(ctx, cells_out, dout, (size_t)n_symbols * t->n_out * 8))) return rc;
  if (sro && (rc = t2_finish_out(ctx, sro, dsro, (size_t)n_symbols * 4))) return rc;
  if (phase && (rc = t2_finish_out(ctx, phase, dph, (size_t)n_symbols * 4))) return rc;
  return T2B200_OK;
}
