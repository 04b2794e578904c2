// K5 + K6: layered offset-min-sum LDPC decoder for every DVB-T2 code, one CTA per PAIR of codewords: the posteriors of
// both codewords sit interleaved in shared memory (one 16-bit unit per bit), every register of the check-node arithmetic
// carries the two codewords as s16x2 halves (DPX instructions, ldpc_pair.h), check-node messages live in an L2-resident
// scratch, and the BCH-parity strip + BB descramble is fused into the epilogue.
//
// Semantics reproduced bit-for-bit (paths relative to the reference's src/DVB_T2):
//   LDPC/layered_decoder.hh:83-110   update(): layers i = 0..q-1, check nodes j = 0..359 SERIALLY
//   LDPC/layered_decoder.hh:65-82    bad(): a check passes only if the product of signs is > 0
//                                    (a zero posterior fails it)
//   LDPC/layered_decoder.hh:168-180  while (bad && --trials >= 0) update
//   LDPC/algorithms.hh:250-291       offset min-sum, beta = 1, int8 saturating, stored message
//                                    clamped to [-32, 31]
//   ldpc_decoder.cpp:262-277         32 codewords in lock-step; hard bit = posterior < 0
//   bch_decoder.cpp:50-61,139-142    PRBS 1+x^14+x^15 (0x4A80), out[i] = in[i] ^ prbs[i], i < K_bch
//
// B200 design (DESIGN.md "K5"):
//   * two codewords per thread: one LDS.U16 / STS.U16, one address computation and one VIADDMNMX.S16x2 / VIMNMX.S16x2
//     serve both -- about a third of the instructions per edge of a one-codeword-per-thread decoder.
//   * posteriors 2 x int8[N] (<= 129.6 KB) stay in shared memory for the whole decode; the stored messages (one byte per
//     edge and codeword, already negated and clamped: ldpc_pair.h) live in a global scratch that only the owning thread
//     touches (L2-resident, prefetched a layer ahead) -> HBM traffic = the compulsory N bytes in, K out.
//   * one CTA of 384 threads per SM for 64 800-bit codes (<= 128 registers), two for 16 200-bit codes: a quarter of the
//     register file, ~80 KB of shared memory and 1 280 threads of every SM stay free for the streaming kernels of
//     another stream.
//   * the quasi-cyclic structure makes every edge of a layer a contiguous (rotated) run of 360 posteriors: thread j of
//     the CTA owns check node (i, j), so all shared-memory traffic is conflict-free and contiguous across a warp; no
//     position table is read, addresses come from q * CNL (base, shift) pairs in the kernel-parameter constant bank.
//   * the reference's serial j order matters only where two check nodes of one layer share a bit; such edges are the
//     first slots of their layer (ldpc_schedule.cpp).  A layer with one shared pair is resolved by walking its dependency
//     runs in registers, the others level by level along the precomputed dependency depth; everything else is one
//     parallel step per layer.
//   * lock-step groups of 32 (reference batch semantics) are 16 co-resident CTAs that exchange their parity verdict
//     through one global word per iteration; groups are claimed from an atomic queue.
#include "stages.h"
#include "ldpc_schedule.h"
#include "ldpc_pair.h"
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace {

using namespace t2pair;

constexpr int kThreads = 384;          // 360 check nodes of a layer + 24 idle lanes (12 warps)
constexpr int kSyncStride = 64;        // group-sync words per group (max_trials + 1 <= 64)
constexpr int kMaxLayers = 90;         // q of the rate-1/2 normal code
constexpr int kMaxEdgeWords = 648;     // max q * CNL over all codes (rate 3/5 normal: 72 * 9)

// Passed by value: the whole schedule lives in the kernel-parameter constant bank, so the per-layer
// (group base, shift) pairs are fetched through the constant path, not through the LSU.
struct LdpcParams {
  const int8_t* llr; uint8_t* bits; int32_t* trials_left; int32_t* iters; int8_t* post_out;
  const uint8_t* level; const uint8_t* prbs; unsigned* gsync;
  unsigned* gqueue;                   // [0] next unclaimed group; [1 + slot * (n_groups + 1) + round] group (+1) claimed by a slot's lane 0
  uint32_t* cn_state;                 // [grid][q][NSW][360] stored messages (L2-resident scratch, thread-private)
  unsigned* err_flag;                 // set when a lock-step wait timed out (the host turns it into T2B200_ERR_CUDA)
  int n_cw, group_lanes, max_trials; unsigned flags;
  int N, K, q, k_out;
  // CN (i,j) data edge c reads posterior 360 eg + (j + es2 / 2) mod 360
  uint16_t eg[kMaxEdgeWords];         // [q][CNL]: bit-group
  uint16_t es2[kMaxEdgeWords];        // [q][CNL]: 2 x cyclic shift (byte offset in the interleaved posteriors)
  int16_t cidx[kMaxLayers];           // row of level[] for layers with shared bits
  uint8_t cnt[kMaxLayers], nlev[kMaxLayers];
  uint8_t ns[kMaxLayers];             // per layer: slots 0 .. ns-1 read a bit another check node of the layer also uses
  uint8_t walked[kMaxLayers];         // per layer: one shared pair whose dependency runs are long enough to be walked
};
static_assert(sizeof(LdpcParams) <= 4000, "kernel parameter block");

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// the stored messages: thread-private, re-read every iteration -> cached in L2 only, lines marked evict-last so that the
// streaming traffic (LLRs in, bits out, the other stream's kernels) does not push them out to HBM
__device__ __forceinline__ uint64_t state_policy()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint32_t state_ld(const uint32_t* p, uint64_t pol)
{
  uint32_t v;
  asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void state_st(uint32_t* p, uint32_t v, uint64_t pol)
{
  asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}
constexpr unsigned long long kWaitBudgetNs = 20ull * 1000 * 1000 * 1000;   // lock-step waits give up after 20 s (never seen: the lanes are co-resident)

// Poll *w until pred(value); on time-out raise the error flag and return `fallback` so that the CTA leaves cleanly.
template <class Pred>
__device__ __forceinline__ unsigned wait_word(const unsigned* w, Pred pred, unsigned* err_flag, unsigned fallback)
{
  unsigned v = ld_acquire(w);
  if (pred(v)) return v;
  const unsigned long long t0 = global_ns();
  for (unsigned polls = 0;; ++polls) {
    __nanosleep(64);
    v = ld_acquire(w);
    if (pred(v)) return v;
    if ((polls & 1023u) == 1023u && global_ns() - t0 > kWaitBudgetNs) { atomicExch(err_flag, 1u); return fallback; }
  }
}

// 32 hard-decision bits starting at bit position P of the packed sign plane
__device__ __forceinline__ uint32_t bits32(const uint32_t* hb, int P)
{
  return __funnelshift_r(hb[P >> 5], hb[(P >> 5) + 1], P & 31);
}

// four sign bits of the four bytes of x, byte 0 first in bit 0
__device__ __forceinline__ uint32_t sign_nibble(uint32_t x) { return ((((x >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu; }
__device__ __forceinline__ uint32_t has_zero_byte(uint32_t x) { return (x - 0x01010101u) & ~x & 0x80808080u; }

// bad() (LDPC/layered_decoder.hh:65-82) of both codewords on a snapshot of the posteriors.  The sign planes hbA / hbB (one
// bit per posterior) are packed first; then each thread XORs, for 32 check nodes at a time, the rotated 32-bit windows of
// a sign plane that the layer's edges select.  A zero posterior fails its checks (vsign gives 0), and every bit takes part
// in at least one check.  Returns bit 0: codeword A fails, bit 1: codeword B fails (per thread; the caller reduces).
template <int CNL>
__device__ __forceinline__ int syndrome_bad(const uint16_t* __restrict__ post, uint32_t* __restrict__ hbA,
                                            uint32_t* __restrict__ hbB, const LdpcParams& p)
{
  const int tid = threadIdx.x;
  const int nwords = (p.N + 31) >> 5;
  // one thread packs the signs of 8 posteriors of each codeword (one 16-byte load) into one byte of each plane
  const uint4* pw = reinterpret_cast<const uint4*>(post);
  uint8_t* hA = reinterpret_cast<uint8_t*>(hbA);
  uint8_t* hB = reinterpret_cast<uint8_t*>(hbB);
  uint32_t zeroA = 0, zeroB = 0;
  for (int k = tid; k < p.N / 8; k += kThreads) {                 // N is a multiple of 8
    const uint4 w = pw[k];
    const uint32_t a0 = __byte_perm(w.x, w.y, 0x6420), a1 = __byte_perm(w.z, w.w, 0x6420);
    const uint32_t b0 = __byte_perm(w.x, w.y, 0x7531), b1 = __byte_perm(w.z, w.w, 0x7531);
    zeroA |= has_zero_byte(a0) | has_zero_byte(a1);
    zeroB |= has_zero_byte(b0) | has_zero_byte(b1);
    hA[k] = (uint8_t)(sign_nibble(a0) | (sign_nibble(a1) << 4));
    hB[k] = (uint8_t)(sign_nibble(b0) | (sign_nibble(b1) << 4));
  }
  for (int k = p.N / 8 + tid; k < 4 * (nwords + 2); k += kThreads) { hA[k] = 0; hB[k] = 0; }   // the windows read up to two words past the end
  int bad = (zeroA ? 1 : 0) | (zeroB ? 2 : 0);
  __syncthreads();
  // word-check (layer i, word w): parity of 32 check nodes at once
  auto word_check = [&](const uint32_t* hb, int t) {
    const int i = t / 12, w = t - 12 * i;
    const int cnt = p.cnt[i];
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < CNL; ++c)
      if (c < cnt) {
        const int sh = (int)p.es2[i * CNL + c] >> 1, base = 360 * (int)p.eg[i * CNL + c];
        int o = 32 * w + sh;
        if (o >= 360) o -= 360;
        uint32_t r = bits32(hb, base + o);
        if (o > 328) {                                    // the 32-bit window wraps inside the 360-bit group
          const int n1 = 360 - o;
          r = (r & ((1u << n1) - 1u)) | (bits32(hb, base) << n1);
        }
        x ^= r;
      }
    x ^= bits32(hb, p.K + 360 * i + 32 * w);
    if (i) x ^= bits32(hb, p.K + 360 * (i - 1) + 32 * w);
    else {
      uint32_t r = bits32(hb, p.K + 360 * (p.q - 1) + 32 * w - 1);
      if (w == 0) r &= ~1u;                               // check node (0,0) has a single parity edge
      x ^= r;
    }
    if (w == 11) x &= 0xffu;
    return (int)(x != 0);
  };
  const int n_tasks = p.q * 12;
  for (int t = tid; t < n_tasks; t += kThreads) bad |= word_check(hbA, t) | (word_check(hbB, t) << 1);
  return bad;
}

// A cheap sufficient test for "bad": the posteriors the LAST layer touched are final for the iteration, so the sign parity
// of a check node of that layer is its parity check.  Thread j tests check node (q-1, j); bit 0 / 1: the check of codeword
// A / B fails.  (A zero posterior also fails a check: that, and every other layer, is left to syndrome_bad when this test
// finds nothing.)
template <int CNL>
__device__ __forceinline__ int last_layer_bad(const uint16_t* __restrict__ post, const LdpcParams& p)
{
  const int tid = threadIdx.x;
  if (tid >= 360) return 0;
  const int i = p.q - 1, cnt = p.cnt[i];
  uint32_t x = post[p.K + 360 * i + tid];
  x ^= i ? post[p.K + 360 * (i - 1) + tid] : (tid ? post[p.K + 360 * (p.q - 1) + tid - 1] : 0u);
#pragma unroll
  for (int c = 0; c < CNL; ++c)
    if (c < cnt) {
      int o = tid + ((int)p.es2[i * CNL + c] >> 1);
      if (o >= 360) o -= 360;
      x ^= post[360 * (int)p.eg[i * CNL + c] + o];
    }
  return (int)((x >> 7) & 1u) | (int)((x >> 14) & 2u);
}

// Register budget: one CTA per SM for the 64 800-bit codes (MINB = 1: 128 registers, so that a quarter of the register file
// stays free), two CTAs per SM for the 16 200-bit codes (MINB = 2: 80 registers).
template <int CNL, int MINB> constexpr int kLdpcRegs = MINB == 2 ? 80 : CNL >= 16 ? 160 : 128;   // (16+ slots need more than 128)

template <int CNL, int MINB>
__global__ void __maxnreg__((kLdpcRegs<CNL, MINB>)) ldpc_decode_kernel(const __grid_constant__ LdpcParams p)
{
  using LY = CnLayout<CNL>;
  constexpr int NSW = LY::NSW;
  extern __shared__ __align__(16) unsigned char smem[];
  uint16_t* post = reinterpret_cast<uint16_t*>(smem);                       // [N]: low byte codeword A, high byte codeword B
  const int plane_words = (((p.N + 31) >> 5) + 3) & ~1;
  uint32_t* hbA = reinterpret_cast<uint32_t*>(smem + ((2 * p.N + 15) & ~15));
  uint32_t* hbB = hbA + plane_words;
  uint32_t* walk = hbB + plane_words;                                       // [360][kWalkWords]: parked check-node state of the chain walk
  // Stored messages are private to the thread that owns check node (i, tid): they live in an L2-resident global scratch
  // (ld/st.cg, next layer's words prefetched a layer ahead), [layer][word][360] per resident CTA.
  uint32_t* const state = p.cn_state + (size_t)blockIdx.x * (size_t)(p.q * NSW * 360) + threadIdx.x;
  __shared__ int s_flag, s_group;
  const uint64_t pol = state_policy();

  const int tid = threadIdx.x;
  const int GL = p.group_lanes;                  // codewords per lock-step group: 32, or 1 (every codeword on its own)
  const int GC = GL > 1 ? GL / 2 : 1;            // CTAs per group
  const int GW = GL > 1 ? GL : 2;                // codewords a group of CTAs takes per claim
  const int lane = blockIdx.x % GC, slot = blockIdx.x / GC;
  const int n_groups = (p.n_cw + GW - 1) / GW;
  const bool lockstep = GL > 1;

  // Groups are claimed dynamically (an atomic counter; lane 0 of a slot claims, the slot's other lanes pick the claim
  // up from the queue word of that round): lock-step groups differ in iteration count, a static round-robin would leave
  // slots idle at the end.
  for (int round = 0;; ++round) {
    if (tid == 0) {
      int gg;
      if (lane == 0) {
        // last group first: a trailing partial group costs a full group's time, so it should not be the one left over
        const int claim = (int)atomicAdd(p.gqueue, 1u);
        gg = claim < n_groups ? n_groups - 1 - claim : n_groups;
        if (GC > 1) { __threadfence(); atomicExch(p.gqueue + 1 + (size_t)slot * (n_groups + 1) + round, (unsigned)gg + 1u); }
      } else {
        const unsigned v = wait_word(p.gqueue + 1 + (size_t)slot * (n_groups + 1) + round, [](unsigned x) { return x != 0u; },
                                     p.err_flag, (unsigned)n_groups + 1u);
        gg = (int)v - 1;
      }
      s_group = gg;
    }
    __syncthreads();
    const int g = s_group;
    __syncthreads();
    if (g >= n_groups) break;
    const int cws_here = min(GW, p.n_cw - g * GW);          // codewords of this group
    const int ctas_here = (cws_here + 1) >> 1;
    if (lane >= ctas_here) continue;
    const int cwA = g * GW + 2 * lane;
    const bool haveB = 2 * lane + 1 < cws_here;
    const int cwB = haveB ? cwA + 1 : cwA;                  // an odd tail decodes its last codeword twice (and writes it once)

    // ---- load the two codewords' channel LLRs, interleaved (reset(): the check-node words are not cleared in memory --
    //      the first update() reads them as zero) ----
    {
      const int2* srcA = reinterpret_cast<const int2*>(p.llr + (size_t)cwA * p.N);
      const int2* srcB = reinterpret_cast<const int2*>(p.llr + (size_t)cwB * p.N);
      uint4* dst = reinterpret_cast<uint4*>(post);
      for (int k = tid; k < p.N / 8; k += kThreads) {
        const int2 a = __ldg(srcA + k), b = __ldg(srcB + k);
        uint4 o;
        o.x = __byte_perm((uint32_t)a.x, (uint32_t)b.x, 0x5140); o.y = __byte_perm((uint32_t)a.x, (uint32_t)b.x, 0x7362);
        o.z = __byte_perm((uint32_t)a.y, (uint32_t)b.y, 0x5140); o.w = __byte_perm((uint32_t)a.y, (uint32_t)b.y, 0x7362);
        dst[k] = o;
      }
    }
    __syncthreads();

    // epilogue of one codeword: hard decision (+ BCH strip, BB descramble), posteriors, status
    auto write_out = [&](int which, int cw, int trials, int iters) {
      const bool descr = p.flags & T2B200_LDPC_BCH_DESCRAMBLE;
      const uint32_t sel = which ? 0x7531u : 0x6420u;
      if (p.bits) {
        if (p.flags & T2B200_LDPC_PACK_BITS) {
          uint8_t* out = p.bits + (size_t)cw * (p.k_out / 8);
          const uint4* pw = reinterpret_cast<const uint4*>(post);
          const uint2* pr = reinterpret_cast<const uint2*>(p.prbs);
          for (int b = tid; b < p.k_out / 8; b += kThreads) {
            const uint4 w = pw[b];
            uint32_t lo = (__byte_perm(w.x, w.y, sel) >> 7) & 0x01010101u, hi = (__byte_perm(w.z, w.w, sel) >> 7) & 0x01010101u;
            if (descr) { const uint2 s = __ldg(pr + b); lo ^= s.x; hi ^= s.y; }
            out[b] = (uint8_t)((((lo * 0x08040201u) >> 24) << 4) | ((hi * 0x08040201u) >> 24));     // first bit in the MSB
          }
        } else {
          uint32_t* out = reinterpret_cast<uint32_t*>(p.bits + (size_t)cw * p.k_out);
          const uint2* pw = reinterpret_cast<const uint2*>(post);
          const uint32_t* pr = reinterpret_cast<const uint32_t*>(p.prbs);
          for (int k = tid; k < p.k_out / 4; k += kThreads) {
            const uint2 w = pw[k];
            uint32_t v = (__byte_perm(w.x, w.y, sel) >> 7) & 0x01010101u;       // ldpc_decoder.cpp:270-277
            if (descr) v ^= __ldg(pr + k);                                       // bch_decoder.cpp:139-142
            out[k] = v;
          }
        }
      }
      if (p.post_out) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.post_out + (size_t)cw * p.N);
        const uint2* pw = reinterpret_cast<const uint2*>(post);
        for (int k = tid; k < p.N / 4; k += kThreads) { const uint2 w = pw[k]; dst[k] = __byte_perm(w.x, w.y, sel); }
      }
      if (tid == 0) {
        if (p.trials_left) p.trials_left[cw] = trials;
        if (p.iters) p.iters[cw] = iters;
      }
    };

    int trials = p.max_trials, iters = 0;
    int done = haveB ? 0 : 2;            // per-codeword mode: bit 0 / 1 set once codeword A / B has stopped and been written
    for (;;) {
      int badA = 0, badB = 0;
      if (iters > 0) {                                     // (the layer loop ended with a barrier)
        const int quick = last_layer_bad<CNL>(post, p);
        badA = __syncthreads_or(quick & 1); badB = __syncthreads_or(quick & 2);
      }
      if (lockstep ? !(badA | badB) : !(badA && badB)) {
        const int mine = syndrome_bad<CNL>(post, hbA, hbB, p);
        badA = __syncthreads_or(mine & 1); badB = __syncthreads_or(mine & 2);
      }
      int go;
      if (lockstep) {
        if (tid == 0) {
          unsigned* w = p.gsync + (size_t)g * kSyncStride + iters;
          atomicAdd(w, 1u | ((badA | badB) ? 0x10000u : 0u));
          const unsigned want = (unsigned)ctas_here;
          const unsigned v = wait_word(w, [want](unsigned x) { return (x & 0xffffu) == want; }, p.err_flag, 0u);
          s_flag = (v >> 16) != 0;
        }
        __syncthreads();
        go = s_flag;
      } else {
        // every codeword stops on its own: one that has converged (or is out of trials together with its neighbour) is
        // written the moment it stops; the pair keeps iterating for the other one
        const bool stopA = !(badA && trials > 0), stopB = !(badB && trials > 0);
        if (stopA && !(done & 1)) { write_out(0, cwA, badA ? -1 : trials, iters); done |= 1; }
        if (stopB && !(done & 2)) { write_out(1, cwB, badB ? -1 : trials, iters); done |= 2; }
        go = done != 3;
      }
      if (!(go && --trials >= 0)) break;
      // ---- one update() ----
      const bool owner = tid < 360;
      const bool stored = iters > 0 && owner;                    // first pass: all messages are zero, nothing to read
      uint32_t nw[NSW];
#pragma unroll
      for (int k = 0; k < NSW; ++k) nw[k] = stored ? state_ld(state + k * 360, pol) : 0u;
      for (int i = 0; i < p.q; ++i) {
        uint32_t* const sp = state + (size_t)i * (NSW * 360);
        uint32_t w[NSW];
#pragma unroll
        for (int k = 0; k < NSW; ++k) w[k] = nw[k];
        if (stored && i + 1 < p.q) {
#pragma unroll
          for (int k = 0; k < NSW; ++k) nw[k] = state_ld(sp + (NSW + k) * 360, pol);
        }
        const int cnt = p.cnt[i];
        const int nl = p.nlev[i];
        const uint16_t* eb = p.eg + i * CNL;
        const uint16_t* es = p.es2 + i * CNL;
        CheckNodePair<CNL> cn;
        if (nl == 1) {
          if (owner) {
            cn.begin(post, w);
            if (cnt == CNL) {
              cn.template load<true>(eb, es, 0, cnt, i, tid, p.K, p.q);
              cn.template store<true>(0, cnt, i, tid, w);
            } else {
              cn.template load<false>(eb, es, 0, cnt, i, tid, p.K, p.q);
              cn.template store<false>(0, cnt, i, tid, w);
            }
#pragma unroll
            for (int k = 0; k < NSW; ++k) state_st(sp + k * 360, w[k], pol);
          }
          __syncthreads();
        } else {
          // Two check nodes of this layer use the same bit: the reference runs j = 0..359 serially, so
          // the smaller j must finish that bit first.  Private edges go in parallel (before / after),
          // shared edges are resolved along the dependency chains.
          const int ns = p.ns[i];
          if (owner) {
            cn.begin(post, w);
            cn.template load<false>(eb, es, ns, cnt, i, tid, p.K, p.q);
          }
          if (p.walked[i]) {
            // One pair of slots reads the same bit-group (the common case): slot 1 of check node j is slot 0 of check node
            // j + step, so the check nodes form runs j0, j0 + step, j0 + 2 step, ... (j0 < step) in which each one hands ONE
            // updated posterior to the next, and the last one of a run also shares a bit with the first one of another run.
            // Instead of one barrier per level of these dependency chains, the thread of a run's first check node WALKS its
            // run: first every run head is done (they depend on nothing), then -- one barrier later -- each walker goes down
            // its run with the handed-over posterior in a register; what it needs from the other check nodes is parked in
            // shared memory.
            const int step = mod360(((int)es[1] >> 1) + 360 - ((int)es[0] >> 1));
            const bool head = tid < step, sink = tid + step >= 360;
            uint32_t nI = 0, vO = 0, carry = 0;
            post_ref rI = 0, rO = 0;
            if (owner) {
              nI = cn.template stored_neg<0>();
              rI = cn.template slot_ref<0>(eb, es, tid); rO = cn.template slot_ref<1>(eb, es, tid);
            }
            if (head) {                                                          // run heads depend on nothing: both shared edges now
              cn.template edge_in_value<0>(sat8_add(unpack_post(post_ld(rI)), nI));
              cn.template edge_in_value<1>(sat8_add(unpack_post(post_ld(rO)), cn.template stored_neg<1>()));
              const CnOut o = cn.minima();
              cn.template edge_out<0>(o, true, true);
              carry = cn.template edge_out<1>(o, true, true);
            }
            __syncthreads();
            if (owner && !head) {                                                // what the walker needs from the others
              vO = sat8_add(unpack_post(post_ld(rO)), cn.template stored_neg<1>());   // (a run's last check node reads the bit a head just wrote)
              const CnMags m = cn.mags();                                        // of the private edges only
              uint4 a;
              a.x = m.A0; a.y = m.NA0; a.z = cn.sx; a.w = vO;
              *reinterpret_cast<uint4*>(walk + tid * kWalkWords) = a;
              walk[tid * kWalkWords + 4] = nI;
            }
            __syncthreads();
            if (head) {                                                          // the serial part: one posterior handed down the run
              int j = tid + step;
              const uint32_t* wj = walk + (j < 360 ? j : tid) * kWalkWords;
              uint4 a = *reinterpret_cast<const uint4*>(wj);
              uint32_t ni = wj[4];
              while (j < 360) {                                                  // (the next check node's words are fetched a step ahead)
                const int jn = j + step;
                const uint32_t* wn = walk + (jn < 360 ? jn : tid) * kWalkWords;
                const uint4 an = *reinterpret_cast<const uint4*>(wn);
                const uint32_t nin = wn[4];
                walk[j * kWalkWords + 5] = carry;                                // the posterior check node j starts from
                carry = walk_carry(carry, a.x, a.y, a.z, a.w, ni);
                a = an; ni = nin; j = jn;
              }
            }
            __syncthreads();
            if (owner && !head) {                                                // ... and everybody finishes its own check node
              cn.template edge_in_value<0>(sat8_add(walk[tid * kWalkWords + 5], nI));
              cn.template edge_in_value<1>(vO);
              const CnOut o = cn.minima();
              cn.template edge_out<0>(o, true, true);
              cn.template edge_out<1>(o, true, sink);                            // elsewhere the successor writes this bit's final value
            }
          } else {
            int mylev = 0;
            if (owner) mylev = __ldg(p.level + (int)p.cidx[i] * 360 + tid);
            for (int l = 1; l <= nl; ++l) {
              if (mylev == l) {
                cn.template shared_load<0>(eb, es, ns, tid);
                const CnOut o = cn.minima();
                cn.template shared_store<0>(o, ns);
              }
              __syncthreads();
            }
          }
          if (owner) {
            cn.template store<false>(ns, cnt, i, tid, w);
#pragma unroll
            for (int k = 0; k < NSW; ++k) state_st(sp + k * 360, w[k], pol);
          }
          __syncthreads();
        }
      }
      ++iters;
    }

    if (lockstep) {
      // ldpc_decoder.cpp:262-277: every lane of the group reports the group's trial count
      write_out(0, cwA, trials, iters);
      if (haveB) write_out(1, cwB, trials, iters);
    }
    __syncthreads();
  }
}

template <int CNL, int MINB>
cudaError_t launch(const LdpcParams& p, int grid, size_t smem, cudaStream_t st, bool cooperative)
{
  auto k = ldpc_decode_kernel<CNL, MINB>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // largest shared-memory carve-out: what the resident decoders leave over can hold the other stream's kernels
  T2_CARVEOUT(k);
  if (p.group_lanes > 1 && cooperative) {     // lock-step lanes spin on each other: they must be co-resident
    void* args[] = {(void*)&p};
    return cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(kThreads), args, smem, st);
  }
  k<<<grid, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int CNL, int MINB>
cudaError_t occupancy(size_t smem, int* blocks_per_sm)
{
  auto k = ldpc_decode_kernel<CNL, MINB>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, kThreads, smem);
}

// smallest instantiated CNL >= cnl_max.  The 64 800-bit codes use buckets 5, 8, 9, 12, 16, 20 with one CTA per SM
// (MINB = 1), the 16 200-bit codes 4, 5, 7, 8, 11 with two (MINB = 2; 17 with one).
const int kCnlBuckets[] = {4, 5, 7, 8, 9, 11, 12, 13, 16, 17, 20};

#define DISPATCH(CALL)                                                                    \
  if (minb == 1) switch (cnl) {                                                           \
    case 5:  return CALL(5, 1);   case 8:  return CALL(8, 1);   case 9:  return CALL(9, 1);  \
    case 12: return CALL(12, 1);  case 16: return CALL(16, 1);  case 20: return CALL(20, 1); \
    case 17: return CALL(17, 1);                                                          \
    default: return cudaErrorInvalidValue;                                                \
  }                                                                                       \
  switch (cnl) {                                                                          \
    case 4:  return CALL(4, 2);   case 5:  return CALL(5, 2);   case 7:  return CALL(7, 2);  \
    case 8:  return CALL(8, 2);   case 11: return CALL(11, 2);                               \
    default: return cudaErrorInvalidValue;                                                \
  }

cudaError_t launch_dispatch(int cnl, int minb, const LdpcParams& p, int grid, size_t smem, cudaStream_t st, bool cooperative)
{
#define CALL_L(C, B) launch<C, B>(p, grid, smem, st, cooperative)
  DISPATCH(CALL_L)
}
cudaError_t occupancy_dispatch(int cnl, int minb, size_t smem, int* bps)
{
#define CALL_O(C, B) occupancy<C, B>(smem, bps)
  DISPATCH(CALL_O)
}

}  // namespace

struct LdpcDeviceCode {
  LdpcSchedule s;
  int cnl = 0;              // instantiated bucket
  int minb = 1;             // co-resident CTAs per SM the kernel is built for (1: 64 800-bit codes, 2: 16 200-bit codes)
  size_t state_bytes = 4, smem = 0;     // stored messages of one PAIR of codewords per check node; shared memory per CTA
  int blocks_per_sm = 0;
  uint8_t* d_level = nullptr;
  LdpcParams proto;         // schedule part of the kernel parameters, filled once
};

static int get_code(t2b200_ctx* ctx, int code, LdpcDeviceCode** out)
{
  auto it = ctx->ldpc.find(code);
  if (it != ctx->ldpc.end()) { *out = it->second; return T2B200_OK; }
  LdpcDeviceCode* d = new LdpcDeviceCode();
  if (!t2_build_ldpc_schedule(code, d->s)) { delete d; ctx->err = "unknown LDPC code id"; return T2B200_ERR_ARG; }
  const LdpcSchedule& s = d->s;
  d->cnl = 0;
  for (int b : kCnlBuckets) if (b >= s.cnl_max) { d->cnl = b; break; }
  if (!d->cnl || s.q > kMaxLayers || s.q * d->cnl > kMaxEdgeWords) {
    delete d; ctx->err = "LDPC code geometry not instantiated"; return T2B200_ERR_ARG;
  }
  d->state_bytes = 4 * (size_t)((d->cnl + 3) / 2);          // CnLayout<CNL>::NSW words per check node of the pair
  // shared memory: interleaved posteriors of the pair | two packed sign planes (+ padding words, rounded) | the parking
  // area of the chain walk; check-node words are in global scratch
  d->smem = (size_t)((2 * s.N + 15) & ~15) + 2 * (size_t)((((s.N + 31) >> 5) + 3) & ~1) * 4 + 360 * kWalkWords * 4;
  d->minb = (s.N > 16200 || d->cnl >= 17) ? 1 : 2;          // (the 17-slot check node of short r5/6 does not fit 80 registers)
  LdpcParams& p = d->proto;
  memset(&p, 0, sizeof(p));
  p.N = s.N; p.K = s.K; p.q = s.q;
  for (int i = 0; i < s.q; ++i) {
    for (int c = 0; c < s.cnl_max; ++c) {
      const uint32_t e = s.edge[(size_t)i * s.cnl_max + c];
      const int shift = e ? 360 - (int)(e >> 16) : 0;              // edge word: 360*g + shift | (360 - shift) << 16
      p.es2[i * d->cnl + c] = (uint16_t)(2 * shift);
      p.eg[i * d->cnl + c] = (uint16_t)(((e & 0xffffu) - shift) / 360);
    }
    p.ns[i] = s.ns[i]; p.cidx[i] = s.conflict_index[i]; p.cnt[i] = s.cnt[i]; p.nlev[i] = s.nlev[i];
    if (s.ns[i] > 10) { delete d; ctx->err = "LDPC layer with more than 10 shared edges"; return T2B200_ERR_ARG; }
    // (measured: the walk beats the level passes even for chains of three -- profiles/r02_ldpc_walk_threshold.txt)
    p.walked[i] = s.ns[i] == 2;
  }
  std::vector<uint8_t> level = s.level; if (level.empty()) level.resize(360, 1);
  cudaError_t ce = cudaMalloc(&d->d_level, level.size());
  if (ce == cudaSuccess) ce = cudaMemcpy(d->d_level, level.data(), level.size(), cudaMemcpyHostToDevice);
  p.level = d->d_level;
  if (ce == cudaSuccess) ce = occupancy_dispatch(d->cnl, d->minb, d->smem, &d->blocks_per_sm);
  if (ce != cudaSuccess || d->blocks_per_sm < 1) {
    ctx->err = ce != cudaSuccess ? std::string("LDPC code set-up: ") + cudaGetErrorString(ce) : "LDPC kernel does not fit on an SM";
    cudaFree(d->d_level); delete d;
    return T2B200_ERR_CUDA;
  }
  ctx->ldpc[code] = d;
  *out = d;
  return T2B200_OK;
}

void t2_ldpc_free(t2b200_ctx* ctx)
{
  for (auto& kv : ctx->ldpc) {
    cudaFree(kv.second->d_level);
    delete kv.second;
  }
  ctx->ldpc.clear();
}

static int ensure_prbs(t2b200_ctx* ctx)
{
  if (ctx->d_prbs) return T2B200_OK;
  uint8_t h[54000];
  int sr = 0x4A80;                                   // bch_decoder.cpp:50-61
  for (int i = 0; i < 54000; i++) {
    uint8_t b = ((sr) ^ (sr >> 1)) & 1;
    h[i] = b; sr >>= 1; if (b) sr |= 0x4000;
  }
  T2_CUDA(ctx, cudaMalloc(&ctx->d_prbs, 54000));
  T2_CUDA(ctx, cudaMemcpy(ctx->d_prbs, h, 54000, cudaMemcpyHostToDevice));
  return T2B200_OK;
}

extern "C" int t2b200_ldpc_code_id(int fec_type, int code_rate)
{
  if (code_rate < 0 || code_rate > 5 || (fec_type != 0 && fec_type != 1)) return -1;
  return (fec_type == T2B200_FEC_NORMAL ? 0 : 6) + code_rate;
}
extern "C" int t2b200_ldpc_n(int code) { auto t = t2_ldpc_code_data(code); return t ? t->N : 0; }
extern "C" int t2b200_ldpc_k(int code) { auto t = t2_ldpc_code_data(code); return t ? t->K : 0; }
extern "C" int t2b200_ldpc_k_bch(int code) { return t2_ldpc_k_bch(code); }

// rows (of kSyncStride words) of the zeroed sync area one launch uses: one per group + the claim queue
static size_t ldpc_sync_rows(const LdpcDeviceCode* d, int sm_count, int n_cw, unsigned flags)
{
  if (!(flags & T2B200_LDPC_GROUP32)) return 1;                   // the claim counter only
  const size_t n_groups = ((size_t)n_cw + 31) / 32;
  const size_t slots = std::max<size_t>(1, std::min<size_t>((size_t)d->blocks_per_sm * sm_count / 16, n_groups));
  const size_t queue_rows = (1 + slots * (n_groups + 1) + kSyncStride - 1) / kSyncStride;
  return n_groups + queue_rows;
}

// Launch the decoder on device-resident buffers (all pointers device memory or null).
// sync_off: first row of the sync area this launch may use (rows are zeroed by the caller).
static int ldpc_launch(t2b200_ctx* ctx, LdpcDeviceCode* d, const int8_t* d_llr, int n_cw, uint8_t* d_bits,
                       int32_t* d_tr, int32_t* d_it, int8_t* d_post, int max_trials, unsigned flags, int k_out,
                       size_t sync_off, cudaStream_t st)
{
  LdpcParams p = d->proto;
  p.llr = d_llr; p.bits = d_bits; p.trials_left = d_tr; p.iters = d_it; p.post_out = d_post;
  p.n_cw = n_cw; p.max_trials = max_trials; p.flags = flags;
  p.group_lanes = (flags & T2B200_LDPC_GROUP32) ? 32 : 1;
  p.k_out = k_out;
  p.prbs = ctx->d_prbs;
  const int capacity = d->blocks_per_sm * ctx->sm_count;            // resident CTAs, one pair of codewords each
  int grid;
  if (p.group_lanes > 1) {
    const int n_groups = (n_cw + 31) / 32;
    int slots = std::min(capacity / 16, n_groups);                  // a lock-step group of 32 codewords = 16 CTAs
    if (ctx->ldpc_slots_cap > 0) slots = std::min(slots, ctx->ldpc_slots_cap);
    if (slots < 1) { ctx->err = "GPU cannot co-schedule one 32-lane group"; return T2B200_ERR_CUDA; }
    grid = slots * 16;
    p.gsync = ctx->d_group_sync + sync_off * kSyncStride;
    p.gqueue = p.gsync + (size_t)n_groups * kSyncStride;
  } else {
    grid = std::min(capacity, (n_cw + 1) / 2);
    p.gqueue = ctx->d_group_sync + sync_off * kSyncStride;
  }
  p.err_flag = ctx->d_err_flag;
  {
    // one row of check-node words per resident CTA; the kernel clears its row per codeword
    void* cs; int rc;
    if ((rc = t2_dev_scratch(ctx, 8, (size_t)capacity * d->s.R * d->state_bytes, &cs))) return rc;
    p.cn_state = (uint32_t*)cs;
  }
  // With T2B200_OPT_LDPC_PLAIN_LAUNCH the grid (never larger than what fits the GPU) goes out as an ordinary launch: the
  // head of the next decode on ANOTHER stream then starts on the SMs the tail of this one has left.
  T2_CUDA(ctx, launch_dispatch(d->cnl, d->minb, p, grid, d->smem, st, !ctx->opt_ldpc_plain_launch));
  ctx->launches++;
  return T2B200_OK;
}

static int ensure_group_sync(t2b200_ctx* ctx, size_t rows, cudaStream_t st)
{
  const size_t need = rows * kSyncStride * sizeof(unsigned);
  if (ctx->group_sync_cap < need) {
    T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_group_sync) cudaFree(ctx->d_group_sync);
    ctx->d_group_sync = nullptr; ctx->group_sync_cap = 0;
    T2_CUDA(ctx, cudaMalloc(&ctx->d_group_sync, need + need / 2));
    ctx->group_sync_cap = need + need / 2;
  }
  T2_CUDA(ctx, cudaMemsetAsync(ctx->d_group_sync, 0, need, st));
  return T2B200_OK;
}

// Host-buffer path: the batch is cut into chunks that flow through a 3-stage pipeline
// (H2D on one copy stream | decode on the context stream | D2H on another copy stream) over
// double-buffered device scratch, so PCIe transfers hide behind the decode of the neighbour chunks.
static int ldpc_decode_pipelined(t2b200_ctx* ctx, LdpcDeviceCode* d, const int8_t* llr, int n_cw, uint8_t* bits_out,
                                 int32_t* trials_left, int32_t* iterations, int max_trials, unsigned flags,
                                 int k_out, size_t out_row)
{
  const int N = d->s.N;
  const int chunk = 512;
  const int n_chunks = (n_cw + chunk - 1) / chunk;
  int rc;
  if (!ctx->s_in) {
    T2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    T2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      T2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
      T2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
      T2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
  }
  void *b_in, *b_out, *b_tr, *b_it;
  if ((rc = t2_dev_scratch(ctx, 0, 2 * (size_t)chunk * N, &b_in))) return rc;
  if ((rc = t2_dev_scratch(ctx, 1, 2 * (size_t)chunk * out_row, &b_out))) return rc;
  if ((rc = t2_dev_scratch(ctx, 2, 4 * (size_t)n_cw, &b_tr))) return rc;
  if ((rc = t2_dev_scratch(ctx, 3, 4 * (size_t)n_cw, &b_it))) return rc;
  {
    size_t rows = 0;
    for (int c = 0; c < n_chunks; ++c) rows += ldpc_sync_rows(d, ctx->sm_count, std::min(chunk, n_cw - c * chunk), flags);
    if ((rc = ensure_group_sync(ctx, rows, ctx->stream))) return rc;
  }
  // the copy streams must not start before earlier work on the context stream (memset above, previous calls)
  T2_CUDA(ctx, cudaEventRecord(ctx->ev_k[0], ctx->stream));
  T2_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_k[0], 0));
  T2_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[0], 0));
  size_t sync_off = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const int b = c & 1, c0 = c * chunk, n = std::min(chunk, n_cw - c0);
    int8_t* din = (int8_t*)b_in + (size_t)b * chunk * N;
    uint8_t* dout = (uint8_t*)b_out + (size_t)b * chunk * out_row;
    if (c >= 2) T2_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_k[b], 0));      // decode c-2 consumed din
    T2_CUDA(ctx, cudaMemcpyAsync(din, llr + (size_t)c0 * N, (size_t)n * N, cudaMemcpyHostToDevice, ctx->s_in));
    T2_CUDA(ctx, cudaEventRecord(ctx->ev_in[b], ctx->s_in));
    T2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
    if (c >= 2) T2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));  // D2H c-2 drained dout
    if ((rc = ldpc_launch(ctx, d, din, n, bits_out ? dout : nullptr, (int32_t*)b_tr + c0, (int32_t*)b_it + c0, nullptr,
                          max_trials, flags, k_out, sync_off, ctx->stream))) return rc;
    sync_off += ldpc_sync_rows(d, ctx->sm_count, n, flags);
    T2_CUDA(ctx, cudaEventRecord(ctx->ev_k[b], ctx->stream));
    if (bits_out) {
      T2_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[b], 0));
      T2_CUDA(ctx, cudaMemcpyAsync(bits_out + (size_t)c0 * out_row, dout, (size_t)n * out_row, cudaMemcpyDeviceToHost, ctx->s_out));
      T2_CUDA(ctx, cudaEventRecord(ctx->ev_out[b], ctx->s_out));
    }
  }
  if (trials_left) T2_CUDA(ctx, cudaMemcpyAsync(trials_left, b_tr, 4 * (size_t)n_cw, cudaMemcpyDeviceToHost, ctx->stream));
  if (iterations) T2_CUDA(ctx, cudaMemcpyAsync(iterations, b_it, 4 * (size_t)n_cw, cudaMemcpyDeviceToHost, ctx->stream));
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return T2B200_OK;
}

extern "C" int t2b200_ldpc_decode(t2b200_ctx* ctx, int code, const int8_t* llr, int n_cw, uint8_t* bits_out,
                                  int32_t* trials_left, int32_t* iterations, int8_t* post_out,
                                  int max_trials, unsigned flags)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (!llr || n_cw < 0 || max_trials <= 0 || max_trials > 60) { ctx->err = "t2b200_ldpc_decode: bad argument"; return T2B200_ERR_ARG; }
  if ((flags & T2B200_LDPC_WANT_POST) && !post_out) { ctx->err = "WANT_POST without post_out"; return T2B200_ERR_ARG; }
  if (n_cw == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  LdpcDeviceCode* d = nullptr;
  int rc = get_code(ctx, code, &d);
  if (rc) return rc;
  const LdpcSchedule& s = d->s;
  int k_out = s.K;
  if (flags & T2B200_LDPC_BCH_DESCRAMBLE) {
    k_out = t2_ldpc_k_bch(code);
    if (!k_out) { ctx->err = "code has no BCH geometry"; return T2B200_ERR_ARG; }
    if ((rc = ensure_prbs(ctx))) return rc;
  }
  const size_t out_row = (flags & T2B200_LDPC_PACK_BITS) ? (size_t)k_out / 8 : (size_t)k_out;

  const bool host_in = !t2_is_device_ptr(llr);
  const bool host_out = !bits_out || !t2_is_device_ptr(bits_out);
  if (host_in && host_out && n_cw > 512 && !(flags & T2B200_LDPC_WANT_POST) &&
      (!trials_left || !t2_is_device_ptr(trials_left)) && (!iterations || !t2_is_device_ptr(iterations)))
    return ldpc_decode_pipelined(ctx, d, llr, n_cw, bits_out, trials_left, iterations, max_trials, flags, k_out, out_row);

  const void* d_llr; void *d_bits = nullptr, *d_tr = nullptr, *d_it = nullptr, *d_post = nullptr;
  if ((rc = t2_to_device(ctx, 0, llr, (size_t)n_cw * s.N, &d_llr))) return rc;
  if (bits_out && (rc = t2_out_device(ctx, 1, bits_out, out_row * n_cw, &d_bits))) return rc;
  if (trials_left && (rc = t2_out_device(ctx, 2, trials_left, 4 * (size_t)n_cw, &d_tr))) return rc;
  if (iterations && (rc = t2_out_device(ctx, 3, iterations, 4 * (size_t)n_cw, &d_it))) return rc;
  if ((flags & T2B200_LDPC_WANT_POST) && (rc = t2_out_device(ctx, 4, post_out, (size_t)n_cw * s.N, &d_post))) return rc;
  if ((rc = ensure_group_sync(ctx, ldpc_sync_rows(d, ctx->sm_count, n_cw, flags), ctx->stream))) return rc;
  if ((rc = ldpc_launch(ctx, d, (const int8_t*)d_llr, n_cw, (uint8_t*)d_bits, (int32_t*)d_tr, (int32_t*)d_it,
                        (int8_t*)d_post, max_trials, flags, k_out, 0, ctx->stream))) return rc;
  if (bits_out && (rc = t2_finish_out(ctx, bits_out, d_bits, out_row * n_cw))) return rc;
  if (trials_left && (rc = t2_finish_out(ctx, trials_left, d_tr, 4 * (size_t)n_cw))) return rc;
  if (iterations && (rc = t2_finish_out(ctx, iterations, d_it, 4 * (size_t)n_cw))) return rc;
  if (d_post && (rc = t2_finish_out(ctx, post_out, d_post, (size_t)n_cw * s.N))) return rc;
  return T2B200_OK;
}

// device-level decode for the frame pipeline: everything resident, asynchronous
int t2_ldpc_device(t2b200_ctx* ctx, int code, const int8_t* d_llr, int n_cw, uint8_t* d_bits, int32_t* d_trials,
                   int32_t* d_iters, int max_trials, unsigned flags)
{
  if (max_trials <= 0 || max_trials > 60) { ctx->err = "LDPC: max_trials must be in 1..60"; return T2B200_ERR_ARG; }
  LdpcDeviceCode* d = nullptr;
  int rc = get_code(ctx, code, &d);
  if (rc) return rc;
  int k_out = d->s.K;
  if (flags & T2B200_LDPC_BCH_DESCRAMBLE) {
    k_out = t2_ldpc_k_bch(code);
    if (!k_out) { ctx->err = "code has no BCH geometry"; return T2B200_ERR_ARG; }
    if ((rc = ensure_prbs(ctx))) return rc;
  }
  if (n_cw == 0) return T2B200_OK;
  if ((rc = ensure_group_sync(ctx, ldpc_sync_rows(d, ctx->sm_count, n_cw, flags), ctx->stream))) return rc;
  return ldpc_launch(ctx, d, d_llr, n_cw, d_bits, d_trials, d_iters, nullptr, max_trials, flags, k_out, 0, ctx->stream);
}

// ---- K6 stand-alone -------------------------------------------------------------------------
__global__ void bch_descramble_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                      const uint8_t* __restrict__ prbs, int n_words, int k_ldpc, int k_bch)
{
  const int per = k_bch / 4;
  const size_t total = (size_t)n_words * per;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int w = (int)(t / per), k = (int)(t % per);
    uint32_t v = reinterpret_cast<const uint32_t*>(in + (size_t)w * k_ldpc)[k];
    reinterpret_cast<uint32_t*>(out + (size_t)w * k_bch)[k] = v ^ __ldg(reinterpret_cast<const uint32_t*>(prbs) + k);
  }
}

// device-level strip + descramble: d_in uint8[n_words][K_ldpc] -> d_out uint8[n_words][K_bch]
int t2_bch_descramble_device(t2b200_ctx* ctx, int code, const uint8_t* d_in, int n_words, uint8_t* d_out)
{
  const int k_ldpc = t2b200_ldpc_k(code), k_bch = t2_ldpc_k_bch(code);
  if (!k_bch) { ctx->err = "code has no BCH geometry"; return T2B200_ERR_ARG; }
  int rc;
  if ((rc = ensure_prbs(ctx))) return rc;
  if (n_words == 0) return T2B200_OK;
  size_t total = (size_t)n_words * (k_bch / 4);
  int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 8);
  bch_descramble_kernel<<<grid, 256, 0, ctx->stream>>>(d_in, d_out, ctx->d_prbs, n_words, k_ldpc, k_bch);
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return T2B200_OK;
}

extern "C" int t2b200_bch_descramble(t2b200_ctx* ctx, int code, const uint8_t* bits_in, int n_words, uint8_t* bits_out)
{
  if (!ctx) return T2B200_ERR_ARG;
  const int k_ldpc = t2b200_ldpc_k(code), k_bch = t2_ldpc_k_bch(code);
  if (!bits_in || !bits_out || n_words < 0 || !k_bch) { ctx->err = "t2b200_bch_descramble: bad argument"; return T2B200_ERR_ARG; }
  if (n_words == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  const void* d_in; void* d_out;
  if ((rc = t2_to_device(ctx, 0, bits_in, (size_t)n_words * k_ldpc, &d_in))) return rc;
  if ((rc = t2_out_device(ctx, 1, bits_out, (size_t)n_words * k_bch, &d_out))) return rc;
  if ((rc = t2_bch_descramble_device(ctx, code, (const uint8_t*)d_in, n_words, (uint8_t*)d_out))) return rc;
  return t2_finish_out(ctx, bits_out, d_out, (size_t)n_words * k_bch);
}
