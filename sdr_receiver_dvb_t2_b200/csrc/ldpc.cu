// K5 + K6: layered offset-min-sum LDPC decoder for every DVB-T2 code, one CTA per codeword, two CTAs
// per SM: posteriors in shared memory, check-node messages in an L2-resident scratch, with the
// BCH-parity strip + BB descramble fused into the epilogue.
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
//   * posteriors int8[N] (<= 64.8 KB) stay in shared memory for the whole decode; the min-sum messages of a
//     check node are fully determined by (two clamped minima, arg-min slot, output signs) and are kept as
//     8-16 bytes per check node in a global scratch that only the owning thread touches (L2-resident,
//     prefetched a layer ahead) -> two codewords per SM, HBM traffic = the compulsory N bytes in, K out.
//   * 64 registers per thread + the largest shared-memory carve-out: what two resident decoders leave
//     free on an SM runs the streaming kernels of another stream (see kLdpcRegs).
//   * the quasi-cyclic structure makes every edge of a layer a contiguous (rotated) run of 360
//     posteriors: thread j of the CTA owns check node (i, j), so all shared-memory traffic is
//     conflict-free byte-contiguous across a warp; no position table is read, addresses come from
//     q * CNL (base, shift) pairs.
//   * the reference's serial j order matters only where two check nodes of one layer share a bit;
//     the host precomputes the dependency depth of every check node in such layers and the CTA runs
//     them level by level (ldpc_schedule.cpp), everything else is one parallel step per layer.
//   * lock-step groups of 32 (reference batch semantics) are 32 co-resident CTAs that exchange their
//     parity verdict through one global word per iteration; groups are claimed from an atomic queue.
#include "stages.h"
#include "ldpc_schedule.h"
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace {

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
  uint32_t* cn_state;                 // [grid][NS][R] packed check-node words (L2-resident scratch, thread-private)
  int n_cw, group_lanes, max_trials; unsigned flags;
  int N, K, q, k_out;
  // CN (i,j) data edge c reads posterior eb + (j + es) mod 360
  uint16_t eb[kMaxEdgeWords];         // [q][CNL]: 360 * bit-group
  uint16_t es[kMaxEdgeWords];         // [q][CNL]: cyclic shift
  uint32_t shared[kMaxLayers];        // per layer: data-edge slots whose bit another CN of the layer also uses
  int16_t cidx[kMaxLayers];           // row of level[] for layers with shared bits
  uint8_t cnt[kMaxLayers], nlev[kMaxLayers];
};
static_assert(sizeof(LdpcParams) <= 4000, "kernel parameter block");

// Check-node word layout.  The min-sum messages of a check node are fully determined by (m0, m1, arg-min slot,
// output signs): message of slot c = sign_c * (c == arg-min ? m1 : m0).  They are kept as one 2-bit code per slot
// (bit 0: sign negative, bit 1: slot is the arg-min), one code per NIBBLE, so that a single PRMT (byte permute)
// looks the messages of four slots up in a 4-entry byte table {+m0, -m0, +m1, -m1} and a second PRMT extracts one
// of them sign-extended: 1.25 ALU instructions per edge instead of a compare/select ladder.
template <int CNL> struct CnLayout {
  static constexpr int SLOTS = CNL + 2;
  static constexpr int NW = (SLOTS + 7) / 8;                 // code words, 8 nibbles each
  static constexpr int TAIL = SLOTS - 8 * (NW - 1);          // nibbles in use in the last code word
  static constexpr bool MPACK = TAIL <= 5;                   // m0 | m1 (6 + 6 bits) share the last code word
  static constexpr int NS = NW + (MPACK ? 0 : 1);            // 32-bit words per check node
};

// min/max through PTX so that the optimiser cannot range-narrow the int8-valued data into packed
// 16-bit lanes (it then spends more PRMT pack/unpack instructions than it saves)
__device__ __forceinline__ int smin(int a, int b) { int r; asm("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int smax(int a, int b) { int r; asm("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// byte i of x, sign-extended (PRMT with the sign-replicate selector mode)
template <int I> __device__ __forceinline__ int sext_byte(uint32_t x)
{
  // (__byte_perm masks bit 3 of the selector nibbles; the PTX instruction has the sign-replicate mode)
  int r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "n"(I | ((8 | I) << 4) | ((8 | I) << 8) | ((8 | I) << 12)));
  return r;
}
__device__ __forceinline__ uint32_t pack4(int b0, int b1, int b2, int b3)
{
  return ((uint32_t)b0 & 0xffu) | (((uint32_t)b1 & 0xffu) << 8) | (((uint32_t)b2 & 0xffu) << 16) | ((uint32_t)b3 << 24);
}

enum SlotMode { ALL_SLOTS, PREDICATED, BRANCHED };
constexpr int kKeyIdle = 1 << 20;

// One check node (LDPC/layered_decoder.hh:87-107 + algorithms.hh:250-291), split so that the edges
// of a layer that are private to the check node and the edges it shares with another check node of
// the same layer can be read / written at different times.
template <int CNL>
struct CheckNode {
  using LY = CnLayout<CNL>;
  static constexpr int SLOTS = LY::SLOTS, NW = LY::NW, NG = (SLOTS + 3) / 4;
  int inp[SLOTS], adr[SLOTS];
  int key0, key1, sx;
  uint32_t tin;             // bytes {-clamp(+m0), -clamp(-m0), -clamp(+m1), -clamp(-m1)} of the PREVIOUS iteration:
                            // minus the stored message (clamped to [-32, 31]) by code
  uint32_t cw[NW];          // previous iteration's codes
  uint32_t nin[NG];         // minus stored message of slots 4g .. 4g+3, one byte each
  uint32_t ncw[NW];         // codes being built for this iteration
  int8_t* post;

  __device__ __forceinline__ void begin(int8_t* post_, const uint32_t (&w)[LY::NS]) {
    post = post_;
    const uint32_t mw = w[LY::NS - 1];
    const int m0c = LY::MPACK ? (int)((mw >> 20) & 63u) : (int)(mw & 63u);
    const int m1c = LY::MPACK ? (int)(mw >> 26) : (int)((mw >> 6) & 63u);
    tin = pack4(-smin(m0c, 31), m0c, -smin(m1c, 31), m1c);
#pragma unroll
    for (int k = 0; k < NW; ++k) { cw[k] = w[k]; ncw[k] = 0; }
#pragma unroll
    for (int g = 0; g < NG; ++g) nin[g] = __byte_perm(tin, 0u, (g & 1) ? (cw[g >> 1] >> 16) : cw[g >> 1]);
    key0 = kKeyIdle; key1 = kKeyIdle; sx = 0;
  }
  template <int SLOT> __device__ __forceinline__ int stored_neg() const { return sext_byte<SLOT & 3>(nin[SLOT >> 2]); }
  // the same for a run-time slot number (shared-edge fast path)
  __device__ __forceinline__ int stored_neg_rt(int slot) const {
    // (selects between computed values, not between array elements: keeps cw[] in registers)
    const int sh4 = 4 * (slot & 7);
    uint32_t code = (cw[0] >> sh4) & 3u;
#pragma unroll
    for (int k = 1; k < NW; ++k) { const uint32_t ck = (cw[k] >> sh4) & 3u; code = (slot >> 3) == k ? ck : code; }
    return (int)(int8_t)(tin >> (8 * code));
  }
  template <int SLOT> __device__ __forceinline__ void edge_in(int a, bool active) {
    const int pv = post[a];
    const int v = smax(__viaddmin_s32(pv, stored_neg<SLOT>(), 127), -128);   // vqsub(posterior, stored message)
    // vqabs + unsigned vqsub of beta = 1 are monotone, so they are applied to the two minima only (store());
    // key = (|v| - 1) * 32 + slot orders the edges the same way
    int key = abs(v) * 32 + (SLOT - 32);
    if (!active) key = kKeyIdle;
    if (active) { inp[SLOT] = v; adr[SLOT] = a; }
    key1 = smin(key1, smax(key0, key));
    key0 = smin(key0, key);
    sx ^= active ? v : 0;
  }
  // m0 / m1 / arg-min of everything seen so far
  __device__ __forceinline__ void minima(int& m0, int& m1, int& idn) const {
    m0 = smin(smax(key0 >> 5, 0), 126); m1 = smin(smax(key1 >> 5, 0), 126); idn = key0 & 31;
  }
  template <int SLOT> __device__ __forceinline__ void mark_sign(bool active) {
    if (active && ((sx ^ inp[SLOT]) < 0)) ncw[SLOT >> 3] |= 1u << (4 * (SLOT & 7));   // sign of the product of the other links
  }
  // Data slots: ALL_SLOTS  - every c < CNL is an edge (regular layer, nothing shared);
  //             PREDICATED - slot c takes part iff c < cnt and bit c of mask; inactive slots are
  //                          computed and discarded so the loads still issue back to back;
  //             BRANCHED   - same condition by (warp-uniform) branches: few active slots.
  template <SlotMode MODE, int C>
  __device__ __forceinline__ void load_from(const uint16_t* eb, const uint16_t* es, int cnt, uint32_t mask, int j) {
    if constexpr (C < CNL) {
      const bool on = MODE == ALL_SLOTS ? true : (C < cnt && ((mask >> C) & 1u));
      if (!(MODE == BRANCHED && !on)) {
        const unsigned t = (unsigned)j + (unsigned)es[C];
        const int a = (int)__viaddmin_u32(t, 0xfffffe98u, t) + (int)eb[C];      // eb + (j + es) mod 360
        edge_in<C>(a, on);
      }
      load_from<MODE, C + 1>(eb, es, cnt, mask, j);
    }
  }
  template <SlotMode MODE>
  __device__ __forceinline__ void load(const uint16_t* eb, const uint16_t* es, int cnt, uint32_t mask,
                                       bool with_parity, int i, int j, int K, int q) {
    load_from<MODE, 0>(eb, es, cnt, mask, j);
    if (with_parity) {
      edge_in<CNL>(K + 360 * i + j, true);
      const bool hasB = (i | j) != 0;
      edge_in<CNL + 1>(i ? K + 360 * (i - 1) + j : K + 360 * (q - 1) + (hasB ? j - 1 : 0), hasB);
    }
  }
  template <SlotMode MODE, int C>
  __device__ __forceinline__ void sign_from(int cnt, uint32_t mask) {
    if constexpr (C < CNL) {
      const bool on = MODE == ALL_SLOTS ? true : (C < cnt && ((mask >> C) & 1u));
      if (!(MODE == BRANCHED && !on)) mark_sign<C>(on);
      sign_from<MODE, C + 1>(cnt, mask);
    }
  }
  template <int SLOT> __device__ __forceinline__ void edge_out(const uint32_t (&nout)[NG], bool active) {
    const int o = sext_byte<SLOT & 3>(nout[SLOT >> 2]);       // other(mags[i], mins[0], mins[1]) with the sign
    if (active) post[adr[SLOT]] = (int8_t)smax(__viaddmin_s32(inp[SLOT], o, 127), -128);   // vqadd
  }
  template <SlotMode MODE, int C>
  __device__ __forceinline__ void out_from(const uint32_t (&nout)[NG], int cnt, uint32_t mask) {
    if constexpr (C < CNL) {
      const bool on = MODE == ALL_SLOTS ? true : (C < cnt && ((mask >> C) & 1u));
      if (!(MODE == BRANCHED && !on)) edge_out<C>(nout, on);
      out_from<MODE, C + 1>(nout, cnt, mask);
    }
  }
  // sign codes of the shared slots resolved earlier (run-time slot numbers), merged with compile-time shifts
  template <int C> __device__ __forceinline__ void merge_shared(uint32_t shared_neg) {
    if constexpr (C < CNL) {
      if ((shared_neg >> C) & 1u) ncw[C >> 3] |= 1u << (4 * (C & 7));
      merge_shared<C + 1>(shared_neg);
    }
  }
  // Write the private edges back and finish the check-node word.  shared_neg: bit c set when shared slot c
  // (already written by shared_out) carried a negative output sign.
  template <SlotMode MODE>
  __device__ __forceinline__ void store(int cnt, uint32_t mask, int i, int j, uint32_t shared_neg, uint32_t (&w)[LY::NS]) {
    int m0, m1, idn;
    minima(m0, m1, idn);
    sign_from<MODE, 0>(cnt, mask);
    mark_sign<CNL>(true);
    mark_sign<CNL + 1>((i | j) != 0);
    if (MODE != ALL_SLOTS) merge_shared<0>(shared_neg);
    {
      const uint32_t bit = 2u << (4 * (idn & 7));
#pragma unroll
      for (int k = 0; k < NW; ++k) ncw[k] |= (idn >> 3) == k ? bit : 0u;
    }
    const uint32_t tout = pack4(m0, -m0, m1, -m1);
    uint32_t nout[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) nout[g] = __byte_perm(tout, 0u, (g & 1) ? (ncw[g >> 1] >> 16) : ncw[g >> 1]);
    out_from<MODE, 0>(nout, cnt, mask);
    edge_out<CNL>(nout, true);
    edge_out<CNL + 1>(nout, (i | j) != 0);
    const uint32_t mm = (uint32_t)smin(m0, 32) | ((uint32_t)smin(m1, 32) << 6);
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = ncw[k];
    if (LY::MPACK) w[NW - 1] |= mm << 20; else w[LY::NS - 1] = mm;
  }
  // ---- shared-edge fast path: the slot number is a run-time (warp-uniform) value ----
  __device__ __forceinline__ int shared_in(int slot, int a, int nbl) {
    const int v = smax(__viaddmin_s32((int)post[a], nbl, 127), -128);
    const int key = abs(v) * 32 + (slot - 32);
    key1 = smin(key1, smax(key0, key));
    key0 = smin(key0, key);
    sx ^= v;
    return v;
  }
  // returns 1 when the output sign is negative
  __device__ __forceinline__ uint32_t shared_out(int slot, int a, int v, int m0, int m1, int idn) {
    const bool neg = ((sx ^ v) < 0);
    int o = neg ? -m0 : m0;
    if (slot == idn) o = neg ? -m1 : m1;
    post[a] = (int8_t)smax(__viaddmin_s32(v, o, 127), -128);
    return neg ? 1u : 0u;
  }
  // generic (BRANCHED) shared slots: load / write slot C if it is in `mask`
  template <int C> __device__ __forceinline__ void shared_load_generic(const uint16_t* eb, const uint16_t* es, uint32_t mask, int j) {
    if constexpr (C < CNL) {
      if ((mask >> C) & 1u) {
        const unsigned t = (unsigned)j + (unsigned)es[C];
        const int a = (int)__viaddmin_u32(t, 0xfffffe98u, t) + (int)eb[C];
        edge_in<C>(a, true);
      }
      shared_load_generic<C + 1>(eb, es, mask, j);
    }
  }
  template <int C> __device__ __forceinline__ uint32_t shared_store_generic(uint32_t mask, int m0, int m1, int idn) {
    if constexpr (C < CNL) {
      uint32_t r = 0;
      if ((mask >> C) & 1u) r = shared_out(C, adr[C], inp[C], m0, m1, idn) << C;
      return r | shared_store_generic<C + 1>(mask, m0, m1, idn);
    } else return 0u;
  }
};

constexpr unsigned kSpinLimit = 1u << 22;     // ~64 ns per poll: a few seconds, then the kernel traps instead of hanging the GPU

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 32 hard-decision bits starting at bit position P of the packed sign plane
__device__ __forceinline__ uint32_t bits32(const uint32_t* hb, int P)
{
  return __funnelshift_r(hb[P >> 5], hb[(P >> 5) + 1], P & 31);
}

// bad() (LDPC/layered_decoder.hh:65-82) on a snapshot of the posteriors.  The sign plane hb[] (one bit
// per posterior) is packed 128 posteriors per warp step; then each thread XORs, for 32 check nodes at a
// time, the rotated 32-bit windows of the sign plane that the layer's edges select.  A zero posterior
// fails its checks (vsign gives 0), and every bit takes part in at least one check.
template <int CNL>
__device__ __forceinline__ int syndrome_bad(const int8_t* __restrict__ post, uint32_t* __restrict__ hb,
                                            const LdpcParams& p)
{
  const int tid = threadIdx.x;
  const int nwords = (p.N + 31) >> 5;
  // one thread packs the signs of 8 posteriors (one 8-byte load) into one byte of the plane
  const uint2* pw = reinterpret_cast<const uint2*>(post);
  uint8_t* hbb = reinterpret_cast<uint8_t*>(hb);
  uint32_t anyzero = 0;
  for (int k = tid; k < p.N / 8; k += kThreads) {                 // N is a multiple of 8
    const uint2 w = pw[k];
    anyzero |= ((w.x - 0x01010101u) & ~w.x & 0x80808080u) | ((w.y - 0x01010101u) & ~w.y & 0x80808080u);   // some byte == 0
    const uint32_t lo = ((((w.x >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu;
    const uint32_t hi = ((((w.y >> 7) & 0x01010101u) * 0x01020408u) >> 20) & 0xf0u;
    hbb[k] = (uint8_t)(lo | hi);
  }
  for (int k = p.N / 8 + tid; k < 4 * (nwords + 2); k += kThreads) hbb[k] = 0;   // the windows below read up to two words past the end
  int bad = anyzero != 0;
  __syncthreads();
  // word-check (layer i, word w): parity of 32 check nodes at once
  auto word_check = [&](int t) {
    const int i = t / 12, w = t - 12 * i;
    const int cnt = p.cnt[i];
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < CNL; ++c)
      if (c < cnt) {
        const int sh = (int)p.es[i * CNL + c], base = (int)p.eb[i * CNL + c];
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
  if (tid < n_tasks) bad |= word_check(tid);
  if (n_tasks > kThreads) {                                   // (uniform condition: every thread reaches the barrier)
    // a codeword that still fails among the first 384 word-checks (the usual case until the last iterations) needs no more
    if (__syncthreads_or(bad)) return 1;
    for (int t = tid + kThreads; t < n_tasks; t += kThreads) bad |= word_check(t);
  }
  return bad;
}


// Register budget: 64 per thread for the codes whose check nodes fit (two CTAs = 49 152 registers, 146 KB of shared memory
// and 768 threads per SM), so that a quarter of every SM's register file, 81 KB of shared memory and 1 280 threads stay
// free: the streaming kernels of the OTHER stream (FFT, equaliser, de-interleaver, demapper: all <= 16 384 registers per
// CTA) run underneath a resident decoder instead of waiting for it.
template <int CNL> constexpr int kLdpcRegs = CNL <= 13 ? 64 : 80;

template <int CNL, int MINB>
__global__ void __maxnreg__(kLdpcRegs<CNL>) ldpc_decode_kernel(const __grid_constant__ LdpcParams p)
{
  using LY = CnLayout<CNL>;
  constexpr int NS = LY::NS;
  extern __shared__ __align__(16) unsigned char smem[];
  int8_t* post = reinterpret_cast<int8_t*>(smem);
  uint32_t* hb = reinterpret_cast<uint32_t*>(smem + ((p.N + 15) & ~15));
  const int R = p.N - p.K;
  // Check-node words are private to the thread that owns check node (i, tid): they live in an L2-resident
  // global scratch (ld/st.cg, next layer's words prefetched a layer ahead) so that shared memory only holds the
  // posteriors and two codewords fit on one SM.  NS planes of R words per resident CTA.
  uint32_t* state = p.cn_state + (size_t)blockIdx.x * NS * R;
  __shared__ int s_flag;

  const int tid = threadIdx.x;
  const int GL = p.group_lanes;
  const int lane = blockIdx.x % GL, slot = blockIdx.x / GL;
  const int n_groups = (p.n_cw + GL - 1) / GL;

  // Groups are claimed dynamically (an atomic counter; lane 0 of a slot claims, the slot's other lanes pick the claim
  // up from the queue word of that round): lock-step groups differ in iteration count, a static round-robin would leave
  // slots idle at the end.
  __shared__ int s_group;
  for (int round = 0;; ++round) {
    if (tid == 0) {
      unsigned* w = p.gqueue + 1 + (size_t)slot * (n_groups + 1) + round;
      int gg;
      if (lane == 0) {
        // last group first: a trailing partial group costs a full group's time, so it should not be the one left over
        const int claim = (int)atomicAdd(p.gqueue, 1u);
        gg = claim < n_groups ? n_groups - 1 - claim : n_groups;
        if (GL > 1) { __threadfence(); atomicExch(w, (unsigned)gg + 1u); }
      } else {
        unsigned v, polls = 0;
        while ((v = ld_acquire(w)) == 0u) { __nanosleep(64); if (++polls > kSpinLimit) __trap(); }
        gg = (int)v - 1;
      }
      s_group = gg;
    }
    __syncthreads();
    const int g = s_group;
    __syncthreads();
    if (g >= n_groups) break;
    const int lanes_here = min(GL, p.n_cw - g * GL);
    if (lane >= lanes_here) continue;
    const int cw = g * GL + lane;

    // ---- load the codeword's channel LLRs, clear the check-node state (reset()) ----
    {
      const int2* src = reinterpret_cast<const int2*>(p.llr + (size_t)cw * p.N);
      int2* dst = reinterpret_cast<int2*>(post);
      for (int k = tid; k < p.N / 8; k += kThreads) dst[k] = __ldg(src + k);
      // (reset(): the check-node words are not cleared in memory -- the first update() reads them as zero)
    }
    __syncthreads();

    int trials = p.max_trials, iters = 0;
    for (;;) {
      const int lane_bad = __syncthreads_or(syndrome_bad<CNL>(post, hb, p));
      int group_bad = lane_bad;
      if (GL > 1) {
        if (tid == 0) {
          unsigned* w = p.gsync + (size_t)g * kSyncStride + iters;
          atomicAdd(w, 1u | (lane_bad ? 0x10000u : 0u));
          unsigned v, polls = 0;
          while (((v = ld_acquire(w)) & 0xffffu) != (unsigned)lanes_here) { __nanosleep(64); if (++polls > kSpinLimit) __trap(); }
          s_flag = (v >> 16) != 0;
        }
        __syncthreads();
        group_bad = s_flag;
      }
      if (!(group_bad && --trials >= 0)) break;
      // ---- one update() ----
      const bool stored = iters > 0 && tid < 360;                // first pass: all messages are zero, nothing to read
      uint32_t w_next[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) w_next[k] = stored ? __ldcg(state + k * R + tid) : 0u;
      for (int i = 0; i < p.q; ++i) {
        uint32_t w_cur[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) w_cur[k] = w_next[k];
        if (stored && i + 1 < p.q) {
#pragma unroll
          for (int k = 0; k < NS; ++k) w_next[k] = __ldcg(state + k * R + (i + 1) * 360 + tid);
        }
        const int cnt = p.cnt[i];
        const int nl = p.nlev[i];
        const uint16_t* eb = p.eb + i * CNL;
        const uint16_t* es = p.es + i * CNL;
        CheckNode<CNL> cn;
        if (nl == 1) {
          if (tid < 360) {
            cn.begin(post, w_cur);
            if (cnt == CNL) {
              cn.template load<ALL_SLOTS>(eb, es, cnt, ~0u, true, i, tid, p.K, p.q);
              cn.template store<ALL_SLOTS>(cnt, ~0u, i, tid, 0u, w_cur);
            } else {
              cn.template load<PREDICATED>(eb, es, cnt, ~0u, true, i, tid, p.K, p.q);
              cn.template store<PREDICATED>(cnt, ~0u, i, tid, 0u, w_cur);
            }
#pragma unroll
            for (int k = 0; k < NS; ++k) __stcg(state + k * R + i * 360 + tid, w_cur[k]);
          }
          __syncthreads();
        } else {
          // Two check nodes of this layer use the same bit: the reference runs j = 0..359 serially, so
          // the smaller j must finish that bit first.  Private edges go in parallel (before / after),
          // shared edges are resolved level by level along the dependency chains.
          const uint32_t sh = p.shared[i];
          int mylev = 0;
          uint32_t shared_neg = 0;
          if (tid < 360) {
            mylev = __ldg(p.level + (int)p.cidx[i] * 360 + tid);
            cn.begin(post, w_cur);
            cn.template load<PREDICATED>(eb, es, cnt, ~sh, true, i, tid, p.K, p.q);
          }
          if (__popc(sh) == 2) {
            // one pair of shared edges (the common case): addresses and stored messages are prepared
            // up front so that a level is just load -> min/sign merge -> store on two posteriors
            const int cA = __ffs(sh) - 1, cB = 31 - __clz(sh);
            const unsigned tA = (unsigned)tid + (unsigned)es[cA], tB = (unsigned)tid + (unsigned)es[cB];
            int aA = (int)__viaddmin_u32(tA, 0xfffffe98u, tA) + (int)eb[cA];
            int aB = (int)__viaddmin_u32(tB, 0xfffffe98u, tB) + (int)eb[cB];
            int nA = 0, nB = 0;
            if (tid >= 360) { aA = 0; aB = 0; }
            else { nA = cn.stored_neg_rt(cA); nB = cn.stored_neg_rt(cB); }
            for (int l = 1; l <= nl; ++l) {
              if (mylev == l) {
                const int vA = cn.shared_in(cA, aA, nA);
                const int vB = cn.shared_in(cB, aB, nB);
                int m0, m1, idn;
                cn.minima(m0, m1, idn);
                shared_neg |= cn.shared_out(cA, aA, vA, m0, m1, idn) << cA;
                shared_neg |= cn.shared_out(cB, aB, vB, m0, m1, idn) << cB;
              }
              __syncthreads();
            }
          } else {
            for (int l = 1; l <= nl; ++l) {
              if (mylev == l) {
                cn.template shared_load_generic<0>(eb, es, sh, tid);
                int m0, m1, idn;
                cn.minima(m0, m1, idn);
                shared_neg |= cn.template shared_store_generic<0>(sh, m0, m1, idn);
              }
              __syncthreads();
            }
          }
          if (tid < 360) {
            cn.template store<PREDICATED>(cnt, ~sh, i, tid, shared_neg, w_cur);
#pragma unroll
            for (int k = 0; k < NS; ++k) __stcg(state + k * R + i * 360 + tid, w_cur[k]);
          }
          __syncthreads();
        }
      }
      ++iters;
    }

    // ---- epilogue: hard decision (+ BCH strip, BB descramble), status ----
    const bool descr = p.flags & T2B200_LDPC_BCH_DESCRAMBLE;
    if (p.bits) {
      if (p.flags & T2B200_LDPC_PACK_BITS) {
        uint8_t* out = p.bits + (size_t)cw * (p.k_out / 8);
        for (int b = tid; b < p.k_out / 8; b += kThreads) {
          unsigned v = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            unsigned bit = post[8 * b + k] < 0;
            if (descr) bit ^= __ldg(p.prbs + 8 * b + k);
            v = (v << 1) | bit;
          }
          out[b] = (uint8_t)v;
        }
      } else {
        uint32_t* out = reinterpret_cast<uint32_t*>(p.bits + (size_t)cw * p.k_out);
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(post);
        const uint32_t* pr = reinterpret_cast<const uint32_t*>(p.prbs);
        for (int k = tid; k < p.k_out / 4; k += kThreads) {
          uint32_t v = (pw[k] >> 7) & 0x01010101u;            // ldpc_decoder.cpp:270-277
          if (descr) v ^= __ldg(pr + k);                      // bch_decoder.cpp:139-142
          out[k] = v;
        }
      }
    }
    if (p.post_out) {
      int2* dst = reinterpret_cast<int2*>(p.post_out + (size_t)cw * p.N);
      const int2* src = reinterpret_cast<const int2*>(post);
      for (int k = tid; k < p.N / 8; k += kThreads) dst[k] = src[k];
    }
    if (tid == 0) {
      if (p.trials_left) p.trials_left[cw] = trials;
      if (p.iters) p.iters[cw] = iters;
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
  // largest shared-memory carve-out: what the two resident decoders leave over can hold the other stream's kernels
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

// smallest instantiated CNL >= cnl_max
const int kCnlBuckets[] = {4, 5, 7, 8, 9, 11, 12, 13, 16, 17, 20};

// CNL bucket instantiations (MINB = 2: the register budget of kLdpcRegs is cut for two co-resident CTAs per SM)
#define DISPATCH(CALL)                                                                    \
  switch (cnl) {                                                                          \
    case 4:  return CALL(4, uint32_t, 2);   case 5:  return CALL(5, uint32_t, 2);         \
    case 7:  return CALL(7, uint32_t, 2);   case 8:  return CALL(8, uint32_t, 2);         \
    case 9:  return CALL(9, uint32_t, 2);   case 11: return CALL(11, uint32_t, 2);        \
    case 12: return CALL(12, uint32_t, 2);  case 13: return CALL(13, uint32_t, 2);        \
    case 16: return CALL(16, uint64_t, 2);  case 17: return CALL(17, uint64_t, 2);        \
    case 20: return CALL(20, uint64_t, 2);                                                \
    default: return cudaErrorInvalidValue;                                                \
  }

cudaError_t launch_dispatch(int cnl, int minb, const LdpcParams& p, int grid, size_t smem, cudaStream_t st, bool cooperative)
{
#define CALL_L(C, T, B) launch<C, B>(p, grid, smem, st, cooperative)
  DISPATCH(CALL_L)
}
cudaError_t occupancy_dispatch(int cnl, int minb, size_t smem, int* bps)
{
#define CALL_O(C, T, B) occupancy<C, B>(smem, bps)
  DISPATCH(CALL_O)
}

}  // namespace

struct LdpcDeviceCode {
  LdpcSchedule s;
  int cnl = 0;              // instantiated bucket
  int minb = 2;             // co-resident CTAs per SM the kernel is built for
  size_t state_bytes = 4, smem = 0;
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
  {
    const int slots = d->cnl + 2, nw = (slots + 7) / 8, tail = slots - 8 * (nw - 1);
    d->state_bytes = 4 * (size_t)(nw + (tail <= 5 ? 0 : 1));       // CnLayout<CNL>::NS words per check node
  }
  // shared memory: posteriors | packed sign plane (+ 2 padding words, rounded); check-node words are in global scratch
  d->smem = (size_t)((s.N + 15) & ~15) + (size_t)(((s.N + 31) / 32 + 3) & ~1) * 4;
  LdpcParams& p = d->proto;
  memset(&p, 0, sizeof(p));
  p.N = s.N; p.K = s.K; p.q = s.q;
  for (int i = 0; i < s.q; ++i) {
    for (int c = 0; c < s.cnl_max; ++c) {
      const uint32_t e = s.edge[(size_t)i * s.cnl_max + c];
      const int shift = e ? 360 - (int)(e >> 16) : 0;              // edge word: 360*g + shift | (360 - shift) << 16
      p.es[i * d->cnl + c] = (uint16_t)shift;
      p.eb[i * d->cnl + c] = (uint16_t)((e & 0xffffu) - shift);
    }
    p.shared[i] = s.shared[i]; p.cidx[i] = s.conflict_index[i]; p.cnt[i] = s.cnt[i]; p.nlev[i] = s.nlev[i];
  }
  std::vector<uint8_t> level = s.level; if (level.empty()) level.resize(360, 1);
  T2_CUDA(ctx, cudaMalloc(&d->d_level, level.size()));
  T2_CUDA(ctx, cudaMemcpy(d->d_level, level.data(), level.size(), cudaMemcpyHostToDevice));
  p.level = d->d_level;
  d->minb = 2;
  T2_CUDA(ctx, occupancy_dispatch(d->cnl, d->minb, d->smem, &d->blocks_per_sm));
  if (d->blocks_per_sm < 1) { ctx->err = "LDPC kernel does not fit on an SM"; return T2B200_ERR_CUDA; }
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
  const int gl = (flags & T2B200_LDPC_GROUP32) ? 32 : 1;
  const size_t n_groups = ((size_t)n_cw + gl - 1) / gl;
  const size_t slots = std::max<size_t>(1, std::min<size_t>((size_t)d->blocks_per_sm * sm_count / gl, n_groups));
  if (gl == 1) return 1;                                          // the claim counter only
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
  const int capacity = d->blocks_per_sm * ctx->sm_count;
  int grid;
  if (p.group_lanes > 1) {
    const int n_groups = (n_cw + 31) / 32;
    const int slots = std::min(capacity / 32, n_groups);
    if (slots < 1) { ctx->err = "GPU cannot co-schedule one 32-lane group"; return T2B200_ERR_CUDA; }
    grid = slots * 32;
    p.gsync = ctx->d_group_sync + sync_off * kSyncStride;
    p.gqueue = p.gsync + (size_t)n_groups * kSyncStride;
  } else {
    grid = std::min(capacity, n_cw);
    p.gqueue = ctx->d_group_sync + sync_off * kSyncStride;
  }
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

extern "C" int t2b200_bch_descramble(t2b200_ctx* ctx, int code, const uint8_t* bits_in, int n_words, uint8_t* bits_out)
{
  if (!ctx) return T2B200_ERR_ARG;
  const int k_ldpc = t2b200_ldpc_k(code), k_bch = t2_ldpc_k_bch(code);
  if (!bits_in || !bits_out || n_words < 0 || !k_bch) { ctx->err = "t2b200_bch_descramble: bad argument"; return T2B200_ERR_ARG; }
  if (n_words == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_prbs(ctx))) return rc;
  const void* d_in; void* d_out;
  if ((rc = t2_to_device(ctx, 0, bits_in, (size_t)n_words * k_ldpc, &d_in))) return rc;
  if ((rc = t2_out_device(ctx, 1, bits_out, (size_t)n_words * k_bch, &d_out))) return rc;
  size_t total = (size_t)n_words * (k_bch / 4);
  int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 8);
  bch_descramble_kernel<<<grid, 256, 0, ctx->stream>>>((const uint8_t*)d_in, (uint8_t*)d_out, ctx->d_prbs, n_words, k_ldpc, k_bch);
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return t2_finish_out(ctx, bits_out, d_out, (size_t)n_words * k_bch);
}
