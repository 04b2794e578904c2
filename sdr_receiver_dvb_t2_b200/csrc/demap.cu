// K3: time / cell de-interleaver + cyclic-Q-delay removal;  K4: rotated-QAM soft demapper with the bit
// de-interleaver (column twist) and demux folded into its store addresses.
//
// Reference semantics reproduced (paths relative to the reference's src/DVB_T2):
//   time_deinterleaver.cpp:316-374   cell k of a TI block (arrival order) -> d = (k mod cols)*rows + k div cols,
//                                    a = perm[d]; out[a].re = in.re; out[a-1 cyclic inside the FEC block].im = in.im
//                                    (the Q component is ALWAYS moved back one cell, rotated constellation or not)
//   llr_demapper.cpp:160-776         optional derotation by e^{-j theta}; sum_s / sum_e over the whole TI block from
//                                    hard slicing (incl. the 64-QAM outer-level quirk, :407,:427, and QPSK's first-2048
//                                    -cells rule, :185); precision = 8*a*sum_s/sum_e; per axis
//                                    L0 = x, L1 = |x| - 8a, L2 = |L1| - 4a, L3 = |L2| - 2a (256-QAM; fewer for 64/16);
//                                    round-to-nearest-even(L * precision); (int8) cast that WRAPS (QPSK saturates);
//                                    store through address[] (:110-130)
// B200 design: both stages are HBM-streaming permutations.  K3 is a gather pass (two 4-byte reads per output cell from the
// L2-resident TI block, <= 4.4 MB, one coalesced 8-byte store) that also applies the demapper's derotation in the frame
// pipeline, so the TI block is written once, and leaves tree-order chunk sums of the statistics terms.  K4 is two passes over
// it: the order-dependent sums (every 2048-cell chunk in parallel as one integer increment for the predicted binade, then one
// warp per TI block and sum stitching the chunks -- see "Pass 1b") and the LLR pass, one CTA per FECFRAME: LLRs are
// produced into a 64.8 KB shared-memory image of the de-interleaved frame and leave the SM as coalesced 16-byte stores, so
// the column-twist scatter never reaches HBM as byte writes.  40 bytes of traffic per cell in all (was 72).
#include "stages.h"
#include "fec_tables.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <climits>

struct TiPlp {
  int cells_per_fec = 0, n_fec_max = 0, rows = 0;
  int32_t* d_perm = nullptr;
  uint32_t* d_src = nullptr;      // per OUTPUT cell a: (row << 16 | column) of the memory cell d = perm^-1[a] that lands there
};
struct DemapTable { int32_t* d_addr = nullptr; };
struct TiDemapState {
  std::map<int, TiPlp> plp;
  std::map<int, DemapTable> addr;        // key fec_type*100 + mod*10 + code_rate class
};

namespace {

constexpr float kRot[4] = {0.506145483f, 0.293215314f, 0.150098316f, 0.062418810f};      // dvbt2_definition.h:45-48
constexpr float kNorm[4] = {0.707106781f, 0.316227766f, 0.15430335f, 0.076696499f};      // dvbt2_definition.h:49-52

// ---- K3 ------------------------------------------------------------------------------------

// Gather form: one thread per OUTPUT cell.  The table gives, for output cell a, the (row, column) of the interleaver
// memory cell that lands there, i.e. arrival index k = row * cols + column; the Q component comes from the cell that lands
// on a + 1 (cyclic inside the FEC block).  Two coalesced table reads, two 4-byte gathers from the L2-resident TI block,
// one coalesced 8-byte store -- no scattered stores.
constexpr int kSumChunk = 2048;          // cells per chunk of the ordered sums (and per CTA step of the fused TI pass)
constexpr int kChunkThreads = 256;

// (sum over the CTA of v.x, of v.y) -> thread 0
__device__ __forceinline__ float2 block_sum2(float2 v, float2* red)
{
#pragma unroll
  for (int off = 16; off; off >>= 1) { v.x += __shfl_down_sync(0xffffffffu, v.x, off); v.y += __shfl_down_sync(0xffffffffu, v.y, off); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kChunkThreads / 32 ? red[lane] : make_float2(0.f, 0.f);
#pragma unroll
    for (int off = 4; off; off >>= 1) { v.x += __shfl_down_sync(0xffffffffu, v.x, off); v.y += __shfl_down_sync(0xffffffffu, v.y, off); }
  }
  __syncthreads();
  return v;
}

template <int MOD> __device__ __forceinline__ float2 demap_term(float2 v, float a);

// MOD >= 0 (frame pipeline): the demapper's derotation (llr_demapper.cpp:555-557, _in[i] *= derotate) is applied on the way
// out, so that the TI block is written once, already derotated, and the pass also leaves per chunk of kSumChunk cells the
// (tree-order) sums of the two statistics terms: the estimate the ordered-sum kernels predict the running sum's binade
// from.  MOD < 0 (stand-alone stage): the cells leave as the reference's time_deinterleaver writes them.
template <int MOD>
__global__ void __launch_bounds__(kChunkThreads) ti_deinterleave_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                         const uint32_t* __restrict__ srcmap,
                                                                         const TiBlockDesc* __restrict__ blocks, int rows, int cpf,
                                                                         int rotate, float rc, float rs, float2* __restrict__ partial,
                                                                         int chunks_max)
{
  __shared__ float2 red[kChunkThreads / 32];
  const TiBlockDesc b = blocks[blockIdx.y];
  const int cols = 5 * b.n_fec;
  const int n = cols * rows;
  const float* src = reinterpret_cast<const float*>(in + b.in_off);
  float2* dst = out + b.out_off;
  for (int c0 = blockIdx.x * kSumChunk; c0 < n; c0 += gridDim.x * kSumChunk) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kSumChunk / kChunkThreads; ++u) {
      const int a = c0 + threadIdx.x + u * kChunkThreads;
      if (a < n) {
        const int r = a % cpf;
        const int an = r == cpf - 1 ? a - (cpf - 1) : a + 1;
        const uint32_t s1 = __ldg(srcmap + a), s2 = __ldg(srcmap + an);
        const int k1 = (int)(s1 >> 16) * cols + (int)(s1 & 0xffffu), k2 = (int)(s2 >> 16) * cols + (int)(s2 & 0xffffu);
        float2 v = make_float2(__ldg(src + 2 * k1), __ldg(src + 2 * k2 + 1));
        if (MOD >= 0) {
          if (rotate) v = make_float2(__fsub_rn(__fmul_rn(v.x, rc), __fmul_rn(v.y, rs)), __fadd_rn(__fmul_rn(v.x, rs), __fmul_rn(v.y, rc)));
          const float2 t = demap_term<(MOD < 0 ? 0 : MOD)>(v, kNorm[MOD < 0 ? 0 : MOD]);
          acc.x += t.x; acc.y += t.y;
        }
        dst[a] = v;
      }
    }
    if (MOD >= 0) {
      acc = block_sum2(acc, red);
      if (threadIdx.x == 0) partial[(size_t)blockIdx.y * chunks_max + c0 / kSumChunk] = acc;
    }
  }
}

// ---- K4 pass 1: derotate (in place, like the reference) + hard-decision statistics ------------
template <int MOD>
__device__ __forceinline__ float slice_axis(float x, float a)
{
  // nearest constellation level the way the reference's if-ladders pick it (strict > / < tests)
  if (MOD == 0) return x > 0 ? a : -a;
  // Branch-free: the reference's ladders (strict > on the positive side, strict < on the negative one, x == 0 on the negative
  // side) pick level 2 n + 1 with n = the number of thresholds 2a, 4a, .. that |x| exceeds; (2 n + 1) a as one rounding of the
  // exact product, which is what a * k.0f is (2a is exact, so the FMA rounds the same real number)
  const float ax = fabsf(x);
  float n = ax > 2 * a ? 1.f : 0.f;
  if (MOD >= 2) n += ax > 4 * a ? 1.f : 0.f;
  if (MOD == 2) n += (x > 0 && ax > 6 * a) ? 1.f : 0.f;         // llr_demapper.cpp:407,427: the negative side never reaches -7a
  if (MOD == 3) {
    n += ax > 6 * a ? 1.f : 0.f;
    n += ax > 8 * a ? 1.f : 0.f;
    n += ax > 10 * a ? 1.f : 0.f;
    n += ax > 12 * a ? 1.f : 0.f;
    n += ax > 14 * a ? 1.f : 0.f;
  }
  const float s = __fmaf_rn(n, 2 * a, a);
  return x > 0 ? s : -s;
}


// levels k*a must be the same floats the reference holds (norm_x_k = NORM * k.0f): computed as a*k in float.
// (|s|^2, |e|^2) of the hard decision on one (derotated) cell: the two terms the reference accumulates
template <int MOD>
__device__ __forceinline__ float2 demap_term(float2 v, float a)
{
  const float sx = slice_axis<MOD>(v.x, a), sy = slice_axis<MOD>(v.y, a);
  const float ex = __fsub_rn(v.x, sx), ey = __fsub_rn(v.y, sy);
  return make_float2(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
}

// Pass 1a (stand-alone stage only): derotate in place, like the reference (llr_demapper.cpp:555-557; no FMA contraction), and
// leave the chunk sums of the statistics terms (see ti_deinterleave_kernel)
template <int MOD>
__global__ void __launch_bounds__(kChunkThreads) demap_prepare_kernel(float2* __restrict__ cells, const DemapBlockDesc* __restrict__ blocks,
                                                                       int rotate, float rc, float rs, float2* __restrict__ partial,
                                                                       int chunks_max)
{
  __shared__ float2 red[kChunkThreads / 32];
  const DemapBlockDesc b = blocks[blockIdx.y];
  float2* c = cells + b.cell_off;
  for (int c0 = blockIdx.x * kSumChunk; c0 < b.n_cells; c0 += gridDim.x * kSumChunk) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kSumChunk / kChunkThreads; ++u) {
      const int k = c0 + threadIdx.x + u * kChunkThreads;
      if (k < b.n_cells) {
        float2 v = c[k];
        if (rotate) {
          v = make_float2(__fsub_rn(__fmul_rn(v.x, rc), __fmul_rn(v.y, rs)), __fadd_rn(__fmul_rn(v.x, rs), __fmul_rn(v.y, rc)));
          c[k] = v;
        }
        const float2 t = demap_term<MOD>(v, kNorm[MOD]);
        acc.x += t.x; acc.y += t.y;
      }
    }
    acc = block_sum2(acc, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * chunks_max + c0 / kSumChunk] = acc;
  }
}

// Pass 1b: sum_s / sum_e exactly as the reference accumulates them -- float, in cell order (the one order-dependent
// reduction of the receiver; every LLR of the TI block is scaled by its result, so a tree sum would flip ~0.1 % of the LLRs
// by one LSB).  The serial recurrence s <- fl(s + t_k) is evaluated EXACTLY in parallel: while s stays inside one binade
// [2^e, 2^(e+1)) it is an integer S (units of ulp = 2^(e-23)) and round-to-nearest-even addition of t >= 0 is
// S += q + [f > 1/2] + [f == 1/2 and S + q odd] with q = floor(t / ulp), f = frac(t / ulp).  Without an exact tie
// (f == 1/2: rare) the increment does not depend on S at all, so
//   demap_sum_chunks_kernel (every chunk of kSumChunk cells of every TI block in parallel) predicts the binade the running sum
//     is in when it reaches the chunk -- from the tree-order chunk sums the previous pass left -- and adds the chunk's
//     increments up as one integer;
//   demap_sum_stitch_kernel (one CTA per TI block and sum) adds the head of the block serially, then walks the chunks: a
//     chunk whose prediction holds (same binade, no tie, no carry out of the binade) is ONE integer addition; the others
//     -- the ~20 chunks in which the sum crosses a power of two, a chunk with a tie, a misprediction -- are redone exactly
//     by the CTA (a scan over (increment if S even, increment if S odd) pairs; the crossing addition itself is a real float
//     addition).
// tools/ordered_sum_model.py is the bit-level model of the arithmetic.
constexpr int kSumSat = 1 << 26;           // increments saturate far above 2^24 (= "left the binade")

struct SumPair { int a0, a1; };            // S + a0 if S is even on entry, S + a1 if odd
__device__ __forceinline__ int sum_sat(int a, int b) { return min(a + b, kSumSat); }
__device__ __forceinline__ SumPair sum_compose(SumPair x, SumPair y)     // x first, then y
{
  SumPair r;
  r.a0 = sum_sat(x.a0, (x.a0 & 1) ? y.a1 : y.a0);
  r.a1 = sum_sat(x.a1, (x.a1 & 1) ? y.a0 : y.a1);
  return r;
}
// q | [f > 1/2] << 30 | [f == 1/2] << 31 of term bits tb against the binade with exponent field es
__device__ __forceinline__ uint32_t sum_elem(int es, uint32_t tb)
{
  if (tb == 0) return 0;
  int et = (tb >> 23) & 0xff;
  uint32_t mt = tb & 0x7fffffu;
  if (et) mt |= 0x800000u; else et = 1;
  const int sh = es - et;
  if (sh < 0) return 1u << 24;              // t >= 2^(e+1): certainly leaves the binade
  if (sh == 0) return mt;
  if (sh > 24) return 0;                    // t < ulp / 2
  const uint32_t q = mt >> sh, rem = mt & ((1u << sh) - 1u), half = 1u << (sh - 1);
  return q | (rem > half ? 1u << 30 : 0u) | (rem == half ? 1u << 31 : 0u);
}
__device__ __forceinline__ int sum_apply(int S, uint32_t w)
{
  const int q = (int)(w & 0x1ffffffu);
  return min(S + q + (int)((w >> 30) & 1u) + (int)((w >> 31) & (uint32_t)(S + q) & 1u), kSumSat);
}

struct SumChunk { int es[2]; int d[2]; };  // per chunk and sum: the binade it was evaluated for (-1: not evaluated), the increment
                                           // (-1: the chunk holds an exact tie)

template <int MOD>
__global__ void __launch_bounds__(kChunkThreads) demap_sum_chunks_kernel(const float2* __restrict__ cells,
                                                                          const DemapBlockDesc* __restrict__ blocks,
                                                                          const float2* __restrict__ partial, int chunks_max,
                                                                          SumChunk* __restrict__ info)
{
  __shared__ float2 red[kChunkThreads / 32];
  __shared__ int ired[kChunkThreads / 32][3];
  __shared__ int s_es[2];
  const DemapBlockDesc b = blocks[blockIdx.y];
  const int n = MOD == 0 ? min(b.n_cells, 2048) : b.n_cells;           // llr_demapper.cpp:185
  const int c = blockIdx.x + 1;                                        // chunk 0 is the serial head of the stitch kernel
  const int c0 = c * kSumChunk;
  if (c0 >= n) return;
  // the running sums when the chunk is reached, to tree-order accuracy: enough to name their binade almost always
  float2 pre = make_float2(0.f, 0.f);
  for (int k = threadIdx.x; k < c; k += kChunkThreads) {
    const float2 t = __ldg(partial + (size_t)blockIdx.y * chunks_max + k);
    pre.x += t.x; pre.y += t.y;
  }
  pre = block_sum2(pre, red);
  if (threadIdx.x == 0) { s_es[0] = (__float_as_uint(pre.x) >> 23) & 0xff; s_es[1] = (__float_as_uint(pre.y) >> 23) & 0xff; }
  __syncthreads();
  const int es0 = s_es[0], es1 = s_es[1];
  const float2* cp = cells + b.cell_off + c0;
  const float a = kNorm[MOD];
  int d0 = 0, d1 = 0;
  uint32_t ties = 0;
#pragma unroll
  for (int u = 0; u < kSumChunk / kChunkThreads; ++u) {
    const int k = threadIdx.x + u * kChunkThreads;
    if (c0 + k < n) {
      const float2 t = demap_term<MOD>(__ldg(cp + k), a);
      const uint32_t w0 = sum_elem(es0, __float_as_uint(t.x)), w1 = sum_elem(es1, __float_as_uint(t.y));
      d0 = sum_sat(d0, (int)(w0 & 0x1ffffffu) + (int)((w0 >> 30) & 1u));
      d1 = sum_sat(d1, (int)(w1 & 0x1ffffffu) + (int)((w1 >> 30) & 1u));
      ties |= ((w0 >> 31) & 1u) | ((w1 >> 30) & 2u);
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    d0 = sum_sat(d0, __shfl_down_sync(0xffffffffu, d0, off));
    d1 = sum_sat(d1, __shfl_down_sync(0xffffffffu, d1, off));
    ties |= __shfl_down_sync(0xffffffffu, ties, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ired[warp][0] = d0; ired[warp][1] = d1; ired[warp][2] = (int)ties; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kChunkThreads / 32; ++w) { d0 = sum_sat(d0, ired[w][0]); d1 = sum_sat(d1, ired[w][1]); ties |= (uint32_t)ired[w][2]; }
    SumChunk o;
    o.es[0] = (es0 == 0 || es0 == 255) ? -1 : es0; o.es[1] = (es1 == 0 || es1 == 255) ? -1 : es1;
    o.d[0] = (ties & 1u) ? -1 : d0; o.d[1] = (ties & 2u) ? -1 : d1;
    info[(size_t)blockIdx.y * chunks_max + c] = o;
  }
}

// Exact evaluation of the chunk's terms buf[0 .. m) (staged in shared memory, skewed: see sum_slot) by the CTA, starting from
// the running sum *sh_s (updated in place): every thread takes a run of kSumE consecutive terms as a (increment if S even,
// increment if S odd) pair, the pairs compose associatively (warp scan + scan of the warp totals); the thread in whose run
// the sum leaves the binade finds the addition that does it, which is then made in real float arithmetic, and the scan
// resumes behind it.  All threads of the CTA call it; it ends with a barrier.
constexpr int kStitchThreads = 256;
constexpr int kSumE = kSumChunk / kStitchThreads;                               // terms per thread
__device__ __forceinline__ int sum_slot(int k) { return k + k / kSumE; }        // the threads' runs start in different banks

__device__ void sum_exact_cta(const float* __restrict__ buf, int m_total, float* sh_s, SumPair* wt, int* sh_cross)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int k0 = 0, pass = 0;
  if (tid == 0) { sh_cross[0] = INT_MAX; sh_cross[1] = INT_MAX; }
  __syncthreads();
  while (k0 < m_total) {
    const float s = *sh_s;
    const uint32_t sb = __float_as_uint(s);
    const int es = (sb >> 23) & 0xff;
    if (es == 0 || es == 255) {                                     // zero / denormal / non-finite sum: one plain addition
      __syncthreads();
      if (tid == 0) *sh_s = __fadd_rn(s, buf[sum_slot(k0)]);
      __syncthreads();
      ++k0;
      continue;
    }
    const int S0 = (int)((sb & 0x7fffffu) | 0x800000u);
    const int lo = max(k0, tid * kSumE), hi = min(m_total, (tid + 1) * kSumE);
    int x0 = 0, x1 = 1;                                             // pseudo-S started even / odd
    for (int k = lo; k < hi; ++k) {
      const uint32_t w = sum_elem(es, __float_as_uint(buf[sum_slot(k)]));
      x0 = sum_apply(x0, w); x1 = sum_apply(x1, w);
    }
    SumPair p = {x0, x1 - 1};
    const SumPair own = p;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      SumPair o;
      o.a0 = __shfl_up_sync(0xffffffffu, p.a0, off); o.a1 = __shfl_up_sync(0xffffffffu, p.a1, off);
      if (lane >= off) p = sum_compose(o, p);
    }
    if (lane == 31) wt[warp] = p;
    SumPair e;                                                      // exclusive inside the warp
    e.a0 = __shfl_up_sync(0xffffffffu, p.a0, 1); e.a1 = __shfl_up_sync(0xffffffffu, p.a1, 1);
    if (lane == 0) e.a0 = e.a1 = 0;
    __syncthreads();
    SumPair q;                                                      // every warp scans the warp totals for itself
    {
      constexpr int NWARP = kStitchThreads / 32;
      SumPair t = {0, 0};
      if (lane < NWARP) t = wt[lane];
#pragma unroll
      for (int off = 1; off < NWARP; off <<= 1) {
        SumPair o;
        o.a0 = __shfl_up_sync(0xffffffffu, t.a0, off); o.a1 = __shfl_up_sync(0xffffffffu, t.a1, off);
        if (lane >= off) t = sum_compose(o, t);
      }
      const int srcl = warp == 0 ? 0 : warp - 1;
      SumPair x;
      x.a0 = __shfl_sync(0xffffffffu, t.a0, srcl); x.a1 = __shfl_sync(0xffffffffu, t.a1, srcl);
      if (warp == 0) x.a0 = x.a1 = 0;
      q = sum_compose(x, e);
    }
    int Sa = sum_sat(S0, (S0 & 1) ? q.a1 : q.a0);                   // S when this thread's run starts
    int my_cross = INT_MAX, Sa_before = Sa;
    const int Sout = sum_sat(Sa, (Sa & 1) ? own.a1 : own.a0);
    if (Sa < (1 << 24)) {
      if (Sout >= (1 << 24)) {                                      // S only grows: at most one thread per pass
        for (int k = lo; k < hi; ++k) {
          const int na = sum_apply(Sa, sum_elem(es, __float_as_uint(buf[sum_slot(k)])));
          if (na >= (1 << 24)) { my_cross = k; Sa_before = Sa; break; }
          Sa = na;
        }
      } else Sa = Sout;
    }
    if (tid == 0) sh_cross[(pass + 1) & 1] = INT_MAX;               // next pass's slot: last read before this pass's first barrier
    if (my_cross != INT_MAX) atomicMin(&sh_cross[pass & 1], my_cross);
    __syncthreads();
    const int cross = sh_cross[pass & 1];
    ++pass;
    if (cross == INT_MAX) {
      if (tid == (m_total - 1) / kSumE) *sh_s = __uint_as_float(((uint32_t)es << 23) | ((uint32_t)Sa & 0x7fffffu));   // owner of the last term holds the total
      k0 = m_total;
    } else {
      if (my_cross == cross)                                        // the crossing addition itself, in real float arithmetic
        *sh_s = __fadd_rn(__uint_as_float(((uint32_t)es << 23) | ((uint32_t)Sa_before & 0x7fffffu)), buf[sum_slot(cross)]);
      k0 = cross + 1;
    }
    __syncthreads();
  }
}

constexpr int kStitchInfo = 1024;          // chunk records staged in shared memory (more chunks than that are read from global memory)

template <int MOD>
__global__ void __launch_bounds__(kStitchThreads) demap_sum_stitch_kernel(const float2* __restrict__ cells,
                                                                           const DemapBlockDesc* __restrict__ blocks,
                                                                           const SumChunk* __restrict__ info, int chunks_max,
                                                                           float* __restrict__ sums)
{
  __shared__ float buf[kSumChunk + kStitchThreads];
  __shared__ int2 inf[kStitchInfo];
  __shared__ SumPair wt[kStitchThreads / 32];
  __shared__ float sh_s;
  __shared__ int sh_cross[2], sh_ch;
  const int which = blockIdx.y, tid = threadIdx.x;
  const DemapBlockDesc b = blocks[blockIdx.x];
  const float2* c = cells + b.cell_off;
  const float a = kNorm[MOD];
  const int n = MOD == 0 ? min(b.n_cells, 2048) : b.n_cells;           // llr_demapper.cpp:185
  const SumChunk* ci = info + (size_t)blockIdx.x * chunks_max;
  const int nch = (n + kSumChunk - 1) / kSumChunk;
  for (int k = 1 + tid; k < min(nch, kStitchInfo); k += kStitchThreads) inf[k] = make_int2(__ldg(&ci[k].es[which]), __ldg(&ci[k].d[which]));
  // the head: the sums double every few cells at first -- plain serial additions, operands staged in shared memory
  const int nh = min(n, kSumChunk);
  for (int k = tid; k < nh; k += kStitchThreads) { const float2 t = demap_term<MOD>(__ldg(c + k), a); buf[k] = which ? t.y : t.x; }
  __syncthreads();
  int ch = 1;
  if (tid == 0) {
    float s = 0.0f;
    int k = 0;
    for (; k + 8 <= nh; k += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = buf[k + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) s = __fadd_rn(s, v[u]);
    }
    for (; k < nh; ++k) s = __fadd_rn(s, buf[k]);
    sh_s = s;
  }
  for (;;) {
    // thread 0 takes the chunks whose prediction holds -- one integer addition each -- up to the next one that must be redone
    if (tid == 0) {
      float s = sh_s;
      for (; ch < nch; ++ch) {
        const int2 rec = ch < kStitchInfo ? inf[ch] : make_int2(__ldg(&ci[ch].es[which]), __ldg(&ci[ch].d[which]));
        const uint32_t sb = __float_as_uint(s);
        const int S = (int)((sb & 0x7fffffu) | 0x800000u);
        if (!((int)((sb >> 23) & 0xff) == rec.x && rec.y >= 0 && S + rec.y < (1 << 24))) break;
        s = __uint_as_float((sb & 0x7f800000u) | ((uint32_t)(S + rec.y) & 0x7fffffu));
      }
      sh_s = s; sh_ch = ch;
    }
    __syncthreads();
    ch = sh_ch;
    if (ch >= nch) break;
    const int k0 = ch * kSumChunk, m = min(n, k0 + kSumChunk) - k0;
    for (int k = tid; k < m; k += kStitchThreads) { const float2 t = demap_term<MOD>(__ldg(c + k0 + k), a); buf[sum_slot(k)] = which ? t.y : t.x; }
    __syncthreads();
    sum_exact_cta(buf, m, &sh_s, wt, sh_cross);
    ++ch;
  }
  if (tid == 0) sums[2 * blockIdx.x + which] = sh_s;
}

// precision = 8.0f * NORM * sum_s / sum_e, left to right (llr_demapper.cpp:722-737); also written out (with the SNR estimate)
// by the first CTA of each TI block
template <int MOD>
__device__ __forceinline__ float demap_precision(const float* __restrict__ sums, const float* __restrict__ precision_in, int blk,
                                                 float* __restrict__ precision, float* __restrict__ snr, bool writer)
{
  const float ss = sums[2 * blk], se = sums[2 * blk + 1];
  float p = __fdiv_rn(__fmul_rn(__fmul_rn(8.0f, kNorm[MOD]), ss), se);
  if (precision_in) p = precision_in[blk];
  if (writer) {
    precision[blk] = p;
    if (snr) snr[blk] = (MOD == 0 ? 10.0f : 20.0f) * log10f(ss / se);
  }
  return p;
}

// (int8_t)(float) as x86-64 gcc compiles it: cvttss2si to int32 (0x80000000 when out of range), low byte.
// SAT (non-reference option T2B200_OPT_DEMAP_SATURATE) clamps to [-128, 127] instead of wrapping.
template <bool SAT>
__device__ __forceinline__ int wrap_i8(float r)
{
  if (SAT) return (int)fminf(fmaxf(r, -128.0f), 127.0f);
  if (!(fabsf(r) < 2147483648.0f)) return 0;
  return (int)(int8_t)(__float2int_rz(r) & 0xff);
}

// ---- K4 pass 2: one CTA per FECFRAME -----------------------------------------------------------
template <int MOD, bool SAT>
__global__ void demap_llr_kernel(const float2* __restrict__ cells, const DemapBlockDesc* __restrict__ blocks,
                                 const float* __restrict__ sums, const float* __restrict__ precision_in,
                                 float* __restrict__ precision, float* __restrict__ snr, const int32_t* __restrict__ address,
                                 int8_t* __restrict__ llr, int cpf, int fec_bits)
{
  extern __shared__ __align__(16) int8_t frame[];
  constexpr int BPC = 2 * (MOD + 1);
  const DemapBlockDesc b = blocks[blockIdx.y];
  const int n_fec = b.n_cells / cpf;
  const float a = kNorm[MOD];
  const float p = demap_precision<MOD>(sums, precision_in, blockIdx.y, precision, snr, blockIdx.x == 0 && threadIdx.x == 0);
  for (int f = blockIdx.x; f < n_fec; f += gridDim.x) {
    const float2* c = cells + b.cell_off + (size_t)f * cpf;
    for (int k = threadIdx.x; k < cpf; k += blockDim.x) {
      const float2 v = __ldg(c + k);
      float xi = v.x, xq = v.y;
#pragma unroll
      for (int l = 0; l < MOD + 1; ++l) {
        float ri, rq;
        if (MOD == 0) {                                  // quantize(): nearbyint, then saturate (llr_demapper.cpp:770-776)
          ri = fminf(fmaxf(rintf(__fmul_rn(xi, p)), -128.0f), 127.0f);
          rq = fminf(fmaxf(rintf(__fmul_rn(xq, p)), -128.0f), 127.0f);
        } else {
          ri = rintf(__fmul_rn(xi, p));                  // _mm256_round_ps(x * precision, NEAREST)
          rq = rintf(__fmul_rn(xq, p));
        }
        const int bi = BPC * k + 2 * l;
        const int ai = MOD == 0 ? bi : __ldg(address + bi), aq = MOD == 0 ? bi + 1 : __ldg(address + bi + 1);
        frame[ai] = (int8_t)wrap_i8<SAT>(ri);
        frame[aq] = (int8_t)wrap_i8<SAT>(rq);
        if (l < MOD) {                                   // next level: |x| - (2^(MOD-l)) * a
          const float t = a * (float)(1 << (MOD - l));
          xi = __fsub_rn(fabsf(xi), t);
          xq = __fsub_rn(fabsf(xq), t);
        }
      }
    }
    __syncthreads();
    int4* dst = reinterpret_cast<int4*>(llr + ((size_t)b.first_fec + f) * fec_bits);
    const int4* src = reinterpret_cast<const int4*>(frame);
    if ((fec_bits & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      // 64 800-bit frames: the finished image leaves the SM as ONE bulk copy of the TMA engine (cp.async.bulk shared -> global);
      // the threads only wait until the engine has read the image out of shared memory
      if (threadIdx.x == 0) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(frame);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(sa), "r"(fec_bits) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else if ((((size_t)b.first_fec + f) * fec_bits) % 16 == 0) {
      for (int k = threadIdx.x; k < fec_bits / 16; k += blockDim.x) dst[k] = src[k];
      for (int k = (fec_bits / 16) * 16 + threadIdx.x; k < fec_bits; k += blockDim.x)
        llr[((size_t)b.first_fec + f) * fec_bits + k] = frame[k];
    } else {                                             // 16200-bit frames: odd frames start 8-byte aligned
      int2* d2 = reinterpret_cast<int2*>(llr + ((size_t)b.first_fec + f) * fec_bits);
      const int2* s2 = reinterpret_cast<const int2*>(frame);
      for (int k = threadIdx.x; k < fec_bits / 8; k += blockDim.x) d2[k] = s2[k];
    }
    __syncthreads();
  }
}

TiDemapState* state(t2b200_ctx* ctx)
{
  if (!ctx->ti) ctx->ti = new TiDemapState();
  return ctx->ti;
}

}  // namespace

void t2_ti_free(t2b200_ctx* ctx)
{
  if (!ctx->ti) return;
  for (auto& kv : ctx->ti->plp) { cudaFree(kv.second.d_perm); cudaFree(kv.second.d_src); }
  for (auto& kv : ctx->ti->addr) cudaFree(kv.second.d_addr);
  delete ctx->ti;
  ctx->ti = nullptr;
}

extern "C" int t2b200_cell_permutation(int n_fec_blocks, int cells_per_fec, int32_t* perm_out)
{
  if (n_fec_blocks <= 0 || cells_per_fec <= 0 || !perm_out) return T2B200_ERR_ARG;
  std::vector<int32_t> p;
  t2_cell_deinterleaver_permutation(n_fec_blocks, cells_per_fec, p);
  std::copy(p.begin(), p.end(), perm_out);
  return T2B200_OK;
}

extern "C" int t2b200_demap_address_table(int fec_type, int mod, int code_rate, int32_t* address_out)
{
  std::vector<int32_t> a;
  if (!address_out || !t2_demap_address_table(fec_type, mod, code_rate, a)) return T2B200_ERR_ARG;
  if (a.empty()) { const int n = fec_type ? 64800 : 16200; for (int i = 0; i < n; ++i) address_out[i] = i; return T2B200_OK; }
  std::copy(a.begin(), a.end(), address_out);
  return T2B200_OK;
}

extern "C" int t2b200_freq_deinterleaver_table(int fft_size, int n_cells, int32_t* h_even_out, int32_t* h_odd_out)
{
  std::vector<int32_t> e, o;
  if (!h_even_out || !h_odd_out || !t2_freq_deinterleaver_tables(fft_size, n_cells, e, o)) return T2B200_ERR_ARG;
  std::copy(e.begin(), e.end(), h_even_out);
  std::copy(o.begin(), o.end(), h_odd_out);
  return T2B200_OK;
}

extern "C" int t2b200_ti_configure(t2b200_ctx* ctx, int plp, int fec_type, int mod, int n_fec_blocks_max,
                                   const int32_t* permutation)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (plp < 0 || plp > 255 || mod < 0 || mod > 3 || (fec_type != 0 && fec_type != 1) || n_fec_blocks_max <= 0) {
    ctx->err = "t2b200_ti_configure: bad argument"; return T2B200_ERR_ARG;
  }
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  TiDemapState* st = state(ctx);
  TiPlp& p = st->plp[plp];
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (p.d_perm) { cudaFree(p.d_perm); p.d_perm = nullptr; }
  if (p.d_src) { cudaFree(p.d_src); p.d_src = nullptr; }
  p.cells_per_fec = t2_cells_per_fec(fec_type, mod);
  p.rows = p.cells_per_fec / 5;
  p.n_fec_max = n_fec_blocks_max;
  std::vector<int32_t> own;
  const size_t n = (size_t)n_fec_blocks_max * p.cells_per_fec;
  if (!permutation) { t2_cell_deinterleaver_permutation(n_fec_blocks_max, p.cells_per_fec, own); permutation = own.data(); }
  T2_CUDA(ctx, cudaMalloc(&p.d_perm, n * sizeof(int32_t)));
  T2_CUDA(ctx, cudaMemcpy(p.d_perm, permutation, n * sizeof(int32_t), cudaMemcpyHostToDevice));
  // inverse, split into (row, column) of the interleaver memory so that the kernel needs no division by `rows`
  std::vector<uint32_t> srcmap(n, 0);
  for (size_t d = 0; d < n; ++d) {
    const int32_t a = permutation[d];
    if (a < 0 || (size_t)a >= n) { ctx->err = "t2b200_ti_configure: permutation entry out of range"; return T2B200_ERR_ARG; }
    srcmap[a] = ((uint32_t)(d % p.rows) << 16) | (uint32_t)(d / p.rows);
  }
  T2_CUDA(ctx, cudaMalloc(&p.d_src, n * sizeof(uint32_t)));
  T2_CUDA(ctx, cudaMemcpy(p.d_src, srcmap.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return T2B200_OK;
}

static int upload_descs(t2b200_ctx* ctx, int slot, const void* h, size_t bytes, void** d)
{
  int rc;
  if ((rc = t2_dev_scratch(ctx, slot, bytes, d))) return rc;
  T2_CUDA(ctx, cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return T2B200_OK;
}

static int sum_chunks_max(int max_cells) { return (max_cells + kSumChunk - 1) / kSumChunk; }
constexpr int kPartialSlot = 11;       // scratch slot of the chunk sums handed from the fused TI pass to the demapper

// fuse_mod >= 0 (frame pipeline): the cells leave derotated (if `rotate`) for that constellation and the chunk sums of the
// statistics terms are left for t2_demap_device(..., prepared = true)
int t2_ti_device(t2b200_ctx* ctx, int plp, const float2* d_in, float2* d_out, const TiBlockDesc* d_desc, int n_ti_blocks,
                 int max_cells, int fuse_mod, int rotate)
{
  if (!ctx->ti || !ctx->ti->plp.count(plp)) { ctx->err = "TI: PLP not configured"; return T2B200_ERR_STATE; }
  const TiPlp& p = ctx->ti->plp[plp];
  const int cm = sum_chunks_max(max_cells);
  dim3 grid(std::max(1, std::min(cm, ctx->sm_count * 8)), n_ti_blocks);
  if (fuse_mod >= 0) {
    void* d_part; int rc;
    if ((rc = t2_dev_scratch(ctx, kPartialSlot, (size_t)n_ti_blocks * cm * sizeof(float2), &d_part))) return rc;
    const float th = -kRot[fuse_mod & 3];
    const float rc_ = (float)cos((double)th), rs_ = (float)sin((double)th);   // llr_demapper.cpp:34-41
#define TI(M) ti_deinterleave_kernel<M><<<grid, kChunkThreads, 0, ctx->stream>>>(d_in, d_out, p.d_src, d_desc, p.rows, p.cells_per_fec, \
                                                                              rotate, rc_, rs_, (float2*)d_part, cm)
    switch (fuse_mod & 3) { case 0: TI(0); break; case 1: TI(1); break; case 2: TI(2); break; default: TI(3); break; }
#undef TI
  } else {
    ti_deinterleave_kernel<-1><<<grid, kChunkThreads, 0, ctx->stream>>>(d_in, d_out, p.d_src, d_desc, p.rows, p.cells_per_fec, 0, 1.0f, 0.0f,
                                                                        nullptr, cm);
  }
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return T2B200_OK;
}

int t2_ti_geometry(t2b200_ctx* ctx, int plp, int* cells_per_fec, int* n_fec_max)
{
  if (!ctx->ti || !ctx->ti->plp.count(plp)) { ctx->err = "TI: PLP not configured"; return T2B200_ERR_STATE; }
  *cells_per_fec = ctx->ti->plp[plp].cells_per_fec; *n_fec_max = ctx->ti->plp[plp].n_fec_max;
  return T2B200_OK;
}

extern "C" int t2b200_ti_deinterleave(t2b200_ctx* ctx, int plp, const float* cells_in, int n_ti_blocks,
                                      const int32_t* n_fec_per_block, float* cells_out)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (!cells_in || !cells_out || n_ti_blocks < 0 || !n_fec_per_block) { ctx->err = "t2b200_ti_deinterleave: bad argument"; return T2B200_ERR_ARG; }
  if (!ctx->ti || !ctx->ti->plp.count(plp)) { ctx->err = "t2b200_ti_deinterleave: PLP not configured"; return T2B200_ERR_STATE; }
  if (n_ti_blocks == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  const TiPlp& p = ctx->ti->plp[plp];
  std::vector<TiBlockDesc> d(n_ti_blocks);
  long long off = 0; int max_cells = 0;
  for (int i = 0; i < n_ti_blocks; ++i) {
    if (n_fec_per_block[i] <= 0 || n_fec_per_block[i] > p.n_fec_max) { ctx->err = "TI block larger than configured"; return T2B200_ERR_ARG; }
    d[i] = {off, off, n_fec_per_block[i]};
    off += (long long)n_fec_per_block[i] * p.cells_per_fec;
    max_cells = std::max(max_cells, n_fec_per_block[i] * p.cells_per_fec);
  }
  int rc; const void* din; void *dout, *ddesc;
  if ((rc = t2_to_device(ctx, 0, cells_in, (size_t)off * 8, &din))) return rc;
  if ((rc = t2_out_device(ctx, 1, cells_out, (size_t)off * 8, &dout))) return rc;
  if ((rc = upload_descs(ctx, 5, d.data(), d.size() * sizeof(TiBlockDesc), &ddesc))) return rc;
  if ((rc = t2_ti_device(ctx, plp, (const float2*)din, (float2*)dout, (const TiBlockDesc*)ddesc, n_ti_blocks, max_cells, -1, 0))) return rc;
  // (the descriptor upload above is a pageable-memory copy: the runtime has consumed the host vector when it returns)
  return t2_finish_out(ctx, cells_out, dout, (size_t)off * 8);
}

template <int MOD>
static int demap_launch(t2b200_ctx* ctx, float2* d_cells, const DemapBlockDesc* d_desc, int n_blocks, int max_cells,
                        int rotation, bool prepared, const int32_t* d_addr, int8_t* d_llr, int cpf, int fec_bits, int max_fec,
                        float* d_prec, float* d_snr, const float* d_prec_in)
{
  void *d_sums, *d_part, *d_info;
  int rc0;
  const int cm = sum_chunks_max(max_cells);
  if ((rc0 = t2_dev_scratch(ctx, 4, (size_t)n_blocks * 2 * sizeof(float), &d_sums))) return rc0;
  if ((rc0 = t2_dev_scratch(ctx, kPartialSlot, (size_t)n_blocks * cm * sizeof(float2), &d_part))) return rc0;
  if ((rc0 = t2_dev_scratch(ctx, 12, (size_t)n_blocks * cm * sizeof(SumChunk), &d_info))) return rc0;
  const int gx = std::max(1, std::min(cm, ctx->sm_count * 8));
  if (!prepared) {                       // stand-alone stage: derotate in place + chunk sums (the frame pipeline's TI pass has done both)
    const float th = -kRot[MOD];
    const float rc = (float)cos((double)th), rs = (float)sin((double)th);   // llr_demapper.cpp:34-41
    demap_prepare_kernel<MOD><<<dim3(gx, n_blocks), kChunkThreads, 0, ctx->stream>>>(d_cells, d_desc, rotation != 0, rc, rs, (float2*)d_part, cm);
    T2_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
  }
  if (cm > 1) {
    demap_sum_chunks_kernel<MOD><<<dim3(cm - 1, n_blocks), kChunkThreads, 0, ctx->stream>>>(d_cells, d_desc, (const float2*)d_part, cm,
                                                                                           (SumChunk*)d_info);
    T2_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
  }
  demap_sum_stitch_kernel<MOD><<<dim3(n_blocks, 2), kStitchThreads, 0, ctx->stream>>>(d_cells, d_desc, (const SumChunk*)d_info, cm, (float*)d_sums);
  T2_CUDA(ctx, cudaGetLastError());
  auto k = ctx->opt_demap_saturate ? demap_llr_kernel<MOD, true> : demap_llr_kernel<MOD, false>;
  const size_t smem = (size_t)((fec_bits + 15) & ~15);
  T2_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<dim3(std::min(max_fec, ctx->sm_count * 3), n_blocks), 512, smem, ctx->stream>>>(d_cells, d_desc, (const float*)d_sums, d_prec_in,
                                                                                     d_prec, d_snr, d_addr, d_llr, cpf, fec_bits);
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches += 2;
  return T2B200_OK;
}

int t2_demap_device(t2b200_ctx* ctx, float2* d_cells, const DemapBlockDesc* d_desc, int n_ti_blocks, int max_cells,
                    int max_fec, int mod, int rotation, bool prepared, int fec_type, int code_rate, int8_t* d_llr,
                    float* d_prec, float* d_snr, const float* d_prec_in)
{
  TiDemapState* st = state(ctx);
  const int cpf = t2_cells_per_fec(fec_type, mod), fec_bits = fec_type ? 64800 : 16200;
  const int cr_class = (fec_type && code_rate == 1) ? 1 : (fec_type && mod == 3 && code_rate == 2) ? 2 : 0;
  const int key = fec_type * 100 + mod * 10 + cr_class;
  if (mod != 0 && !st->addr.count(key)) {
    std::vector<int32_t> a;
    t2_demap_address_table(fec_type, mod, code_rate, a);
    DemapTable t;
    T2_CUDA(ctx, cudaMalloc(&t.d_addr, a.size() * 4));
    T2_CUDA(ctx, cudaMemcpy(t.d_addr, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    st->addr[key] = t;
  }
  const int32_t* daddr = mod ? st->addr[key].d_addr : nullptr;
#define DM(M) demap_launch<M>(ctx, d_cells, d_desc, n_ti_blocks, max_cells, rotation, prepared, daddr, d_llr, cpf, fec_bits, \
                              max_fec, d_prec, d_snr, d_prec_in)
  switch (mod) { case 0: return DM(0); case 1: return DM(1); case 2: return DM(2); default: return DM(3); }
#undef DM
}

extern "C" int t2b200_demap(t2b200_ctx* ctx, float* ti_cells, int n_ti_blocks, const int32_t* n_fec_per_block,
                            int mod, int rotation, int fec_type, int code_rate, int8_t* llr_out,
                            float* snr_out, float* precision_out, const float* precision_in)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (!ti_cells || !llr_out || n_ti_blocks < 0 || !n_fec_per_block || mod < 0 || mod > 3 ||
      (fec_type != 0 && fec_type != 1) || code_rate < 0 || code_rate > 5) { ctx->err = "t2b200_demap: bad argument"; return T2B200_ERR_ARG; }
  if (n_ti_blocks == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  const int cpf = t2_cells_per_fec(fec_type, mod), fec_bits = fec_type ? 64800 : 16200;
  std::vector<DemapBlockDesc> d(n_ti_blocks);
  long long off = 0; int fec = 0, max_cells = 0, max_fec = 0;
  for (int i = 0; i < n_ti_blocks; ++i) {
    if (n_fec_per_block[i] <= 0) { ctx->err = "t2b200_demap: empty TI block"; return T2B200_ERR_ARG; }
    d[i] = {off, n_fec_per_block[i] * cpf, fec};
    off += (long long)n_fec_per_block[i] * cpf; fec += n_fec_per_block[i];
    max_cells = std::max(max_cells, n_fec_per_block[i] * cpf); max_fec = std::max(max_fec, n_fec_per_block[i]);
  }
  int rc; const void* dcells; void *dllr, *ddesc, *dprec, *dsnr; const void* dpin = nullptr;
  const bool cells_on_dev = t2_is_device_ptr(ti_cells);
  if ((rc = t2_to_device(ctx, 0, ti_cells, (size_t)off * 8, &dcells))) return rc;
  if ((rc = t2_out_device(ctx, 1, llr_out, (size_t)fec * fec_bits, &dllr))) return rc;
  if ((rc = upload_descs(ctx, 5, d.data(), d.size() * sizeof(DemapBlockDesc), &ddesc))) return rc;
  if ((rc = t2_dev_scratch(ctx, 6, 8 * (size_t)n_ti_blocks + 16, &dprec))) return rc;
  dsnr = (float*)dprec + n_ti_blocks;
  if (precision_in) {
    void* tmp;
    if ((rc = t2_dev_scratch(ctx, 7, 4 * (size_t)n_ti_blocks, &tmp))) return rc;
    T2_CUDA(ctx, cudaMemcpyAsync(tmp, precision_in, 4 * (size_t)n_ti_blocks,
                                 t2_is_device_ptr(precision_in) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    dpin = tmp;
  }
  if ((rc = t2_demap_device(ctx, (float2*)dcells, (const DemapBlockDesc*)ddesc, n_ti_blocks, max_cells, max_fec, mod, rotation, false,
                            fec_type, code_rate, (int8_t*)dllr, (float*)dprec, (float*)dsnr, (const float*)dpin))) return rc;
  auto copy_small = [&](float* dst, const void* src) -> int {
    if (!dst) return T2B200_OK;
    T2_CUDA(ctx, cudaMemcpyAsync(dst, src, 4 * (size_t)n_ti_blocks,
                                 t2_is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    return T2B200_OK;
  };
  if ((rc = copy_small(precision_out, dprec))) return rc;
  if ((rc = copy_small(snr_out, dsnr))) return rc;
  // the reference derotates its input in place (llr_demapper.cpp:555-557): mirror that for host buffers too
  if (!cells_on_dev && rotation)
    T2_CUDA(ctx, cudaMemcpyAsync(ti_cells, dcells, (size_t)off * 8, cudaMemcpyDeviceToHost, ctx->stream));
  // host outputs must be filled when the call returns; with device buffers everywhere the call stays asynchronous
  const bool any_host = !cells_on_dev || (snr_out && !t2_is_device_ptr(snr_out)) || (precision_out && !t2_is_device_ptr(precision_out));
  if (any_host) T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return t2_finish_out(ctx, llr_out, dllr, (size_t)fec * fec_bits);
}
