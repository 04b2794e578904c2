// Check-node arithmetic of the LDPC decoder for a PAIR of codewords: every 32-bit register holds the same quantity of
// two codewords as s16x2 halves (low half: codeword A, high half: codeword B), so that one DPX instruction of sm_100
// (VIADDMNMX.S16x2, VIMNMX.S16x2, VIADD.16x2) or one PRMT / LOP3 serves both.  The arithmetic is the reference's
// (LDPC/layered_decoder.hh:87-107, LDPC/algorithms.hh:250-291: offset min-sum, beta = 1, int8 saturating, stored message
// clamped to [-32, 31]); int8 values live sign-extended in their 16-bit half and are clamped back to [-128, 127] wherever
// the reference saturates.
//
// The header compiles for the host too (T2_LDPC_HOST_EMULATION: the DPX / PRMT instructions are restated in plain C) so
// that tests/cpp/ldpc_pair_emu.cpp can run exactly this code against the CPU oracle without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define T2_HD __host__ __device__ __forceinline__
#else
#define T2_HD inline
#endif

namespace t2pair {

// ---- the handful of machine operations everything below is written in -------------------------------------------------
#if defined(__CUDA_ARCH__)
T2_HD uint32_t vaddmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }   // min(a + b, c) per half
T2_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }   // max(a + b, c) per half
T2_HD uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
T2_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
// prmt.b32 with a selector known at compile time; selector nibbles with bit 3 set replicate the sign of the selected byte
template <uint32_t SEL> T2_HD uint32_t prmt(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(SEL));
  return r;
}
// run-time selector (low 16 bits); the callers' nibbles never have bit 3 set where the result is used
T2_HD uint32_t prmt_rt(uint32_t a, uint32_t b, uint32_t sel)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
T2_HD int mod360(int t) { return (int)__viaddmin_u32((unsigned)t, 0xfffffe98u, (unsigned)t); }    // t in [0, 720): t mod 360
T2_HD uint32_t rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }
T2_HD uint32_t vneg2(uint32_t a) { return __vneg2(a); }                                               // per-half negation
// the pair's posteriors by 32-bit shared-memory address (kept in a register from the load to the store of an edge)
typedef uint32_t post_ref;
T2_HD post_ref post_base(uint16_t* post) { return (uint32_t)__cvta_generic_to_shared(post); }
T2_HD post_ref post_at(post_ref base, int a) { return base + 2u * (uint32_t)a; }
T2_HD uint32_t post_ld(post_ref r) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(r) : "memory"); return v; }
T2_HD void post_st(post_ref r, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(r), "r"(v) : "memory"); }
#else
T2_HD int16_t lo16(uint32_t x) { return (int16_t)(x & 0xffffu); }
T2_HD int16_t hi16(uint32_t x) { return (int16_t)(x >> 16); }
T2_HD uint32_t pk16(int a, int b) { return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); }
T2_HD int imin(int a, int b) { return a < b ? a : b; }
T2_HD int imax(int a, int b) { return a > b ? a : b; }
T2_HD uint32_t vaddmin(uint32_t a, uint32_t b, uint32_t c)
{
  return pk16(imin((int16_t)(lo16(a) + lo16(b)), lo16(c)), imin((int16_t)(hi16(a) + hi16(b)), hi16(c)));
}
T2_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c)
{
  return pk16(imax((int16_t)(lo16(a) + lo16(b)), lo16(c)), imax((int16_t)(hi16(a) + hi16(b)), hi16(c)));
}
T2_HD uint32_t vmin2(uint32_t a, uint32_t b) { return pk16(imin(lo16(a), lo16(b)), imin(hi16(a), hi16(b))); }
T2_HD uint32_t vmax2(uint32_t a, uint32_t b) { return pk16(imax(lo16(a), lo16(b)), imax(hi16(a), hi16(b))); }
T2_HD uint32_t prmt_any(uint32_t a, uint32_t b, uint32_t sel)
{
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = (sel >> (4 * i)) & 0xfu;
    uint32_t byte = (uint32_t)(src >> (8 * (s & 7))) & 0xffu;
    if (s & 8) byte = (byte & 0x80u) ? 0xffu : 0u;
    r |= byte << (8 * i);
  }
  return r;
}
template <uint32_t SEL> T2_HD uint32_t prmt(uint32_t a, uint32_t b) { return prmt_any(a, b, SEL); }
T2_HD uint32_t prmt_rt(uint32_t a, uint32_t b, uint32_t sel) { return prmt_any(a, b, sel); }
T2_HD int mod360(int t) { return t >= 360 ? t - 360 : t; }
T2_HD uint32_t rotl(uint32_t x, int r) { r &= 31; return r ? (x << r) | (x >> (32 - r)) : x; }
T2_HD uint32_t vneg2(uint32_t a) { return pk16(-(int)lo16(a), -(int)hi16(a)); }
typedef uint16_t* post_ref;
T2_HD post_ref post_base(uint16_t* post) { return post; }
T2_HD post_ref post_at(post_ref base, int a) { return base + a; }
T2_HD uint32_t post_ld(post_ref r) { return *r; }
T2_HD void post_st(post_ref r, uint32_t v) { *r = (uint16_t)v; }
#endif

constexpr uint32_t kP127 = 0x007f007fu, kM128 = 0xff80ff80u, kOne2 = 0x00010001u, kIdleKey = 0x7fff7fffu;

// int8 saturating add of two sign-extended pairs (vqadd / vqsub with a negated operand)
T2_HD uint32_t sat8_add(uint32_t a, uint32_t b) { return vmax2(vaddmin(a, b, kP127), kM128); }
T2_HD uint32_t abs2(uint32_t v) { return vaddmax(~v, kOne2, v); }                                   // max(-v, v)
// two posteriors (one per codeword) as they sit in shared memory (low byte A, high byte B) -> sign-extended s16x2
T2_HD uint32_t unpack_post(uint32_t raw16) { return prmt<0x9180u>(raw16, 0u); }
T2_HD uint32_t pack_post(uint32_t v) { return prmt<0x4420u>(v, 0u); }   // low bytes of the two halves, upper half zero
T2_HD uint32_t pack4(int b0, int b1, int b2, int b3)
{
  return ((uint32_t)b0 & 0xffu) | (((uint32_t)b1 & 0xffu) << 8) | (((uint32_t)b2 & 0xffu) << 16) | ((uint32_t)b3 << 24);
}

// Check-node word layout (per codeword).  The min-sum messages of a check node are fully determined by (m0, m1, arg-min
// slot, output signs): message of slot c = sign_c * (c == arg-min ? m1 : m0).  One 2-bit code per slot (bit 0: sign
// negative, bit 1: slot is the arg-min), one code per NIBBLE, so that a single PRMT looks the messages of four slots up in
// a 4-entry byte table {+m0, -m0, +m1, -m1}.  The words of codeword B are kept ROTATED by 16 bits (halves swapped): one
// rotate of (sx ^ input) then drops the output-sign bits of both codewords (bits 15 and 31) onto their nibbles at once.
template <int CNL> struct CnLayout {
  static constexpr int SLOTS = CNL + 2;
  static constexpr int NW = (SLOTS + 7) / 8;                 // code words, 8 nibbles each
  static constexpr int TAIL = SLOTS - 8 * (NW - 1);          // nibbles in use in the last code word
  static constexpr bool MPACK = TAIL <= 5;                   // m0 | m1 (6 + 6 bits) share the last code word
  static constexpr int NS = NW + (MPACK ? 0 : 1);            // 32-bit words per check node and codeword
  // bit position of slot's nibble in its code word; ROT = 0 for codeword A, 16 for codeword B
  template <int ROT> static constexpr int nib(int slot) { return (4 * (slot & 7) + ROT) & 31; }
};

enum SlotMode { ALL_SLOTS, PREDICATED, BRANCHED };

constexpr int kWalkWords = 4;         // words parked per check node by the chain walk (walk_carry's arguments; word 0 returns its carry-in)

// The running state of one check node of the pair -- the two smallest keys and the xor of the inputs seen so far -- as a
// value that can be handed from thread to thread (the chain walk of ldpc.cu parks it in shared memory).
struct CnCore { uint32_t key0, key1, sx; };
T2_HD void core_take(CnCore& c, uint32_t v, int slot)
{
  const uint32_t key = abs2(v) * 32u + (uint32_t)slot * kOne2;
  c.key1 = vmin2(c.key1, vmax2(c.key0, key));
  c.key0 = vmin2(c.key0, key);
  c.sx ^= v;
}
T2_HD void core_minima(const CnCore& c, uint32_t& m0, uint32_t& m1, uint32_t& idn)
{
  const uint32_t c126 = 0x007e007eu, mone = 0xffffffffu;
  m0 = vmin2(vaddmax((c.key0 >> 5) & 0x07ff07ffu, mone, 0u), c126);
  m1 = vmin2(vaddmax((c.key1 >> 5) & 0x07ff07ffu, mone, 0u), c126);
  idn = c.key0 & 0x001f001fu;
}
// new posterior pair of an edge whose input was v (vqadd of v and the check node's message); neg: bit 0 / 16 set when the
// message is negative in codeword A / B
T2_HD uint32_t core_out(const CnCore& c, int slot, uint32_t v, uint32_t m0, uint32_t m1, uint32_t idn, uint32_t& neg)
{
  const uint32_t t = c.sx ^ v;
  const int negA = (t >> 15) & 1u, negB = t >> 31;
  int ma = (int)((slot == (int)(idn & 0xffffu) ? m1 : m0) & 0xffffu);
  int mb = (int)((slot == (int)(idn >> 16) ? m1 : m0) >> 16);
  if (negA) ma = -ma;
  if (negB) mb = -mb;
  neg = (uint32_t)negA | ((uint32_t)negB << 16);
  return sat8_add(v, ((uint32_t)ma & 0xffffu) | ((uint32_t)mb << 16));
}

// The serial core of the chain walk (ldpc.cu): a check node shares one bit with its predecessor in the layer's serial order
// (slot I: its updated posterior pair arrives in `carry`) and one with its successor (slot O).  All the successor needs is
// the new posterior of the O bit: vqadd(vO, message) with |message| = the smallest magnitude among the OTHER inputs (the
// private edges, whose minimum a0mag = (key0 >> 5) is known beforehand, and the I input) after the offset, and sign = product
// of the other inputs' signs (sxp: xor of the private inputs).  ~14 dependent instructions per check node; everything else
// of the check node is done afterwards, in parallel, by its own thread.
T2_HD uint32_t walk_carry(uint32_t carry, uint32_t a0mag, uint32_t sxp, uint32_t vO, uint32_t nI)
{
  const uint32_t vI = sat8_add(carry, nI);
  const uint32_t mag = vmin2(a0mag, abs2(vI));
  const uint32_t m = vmin2(vaddmax(mag, 0xffffffffu, 0u), 0x007e007eu);
  const uint32_t smask = (((sxp ^ vI) >> 15) & kOne2) * 0xffffu;
  return sat8_add(vO, (m & ~smask) | (vneg2(m) & smask));
}

// One check node of BOTH codewords of the pair, split so that the edges private to the check node and the edges it shares
// with another check node of the same layer can be read / written at different times.
template <int CNL>
struct CheckNodePair {
  using LY = CnLayout<CNL>;
  static constexpr int SLOTS = LY::SLOTS, NW = LY::NW, NG = (SLOTS + 3) / 4;
  uint32_t inp[SLOTS];      // vqsub(posterior, stored message), s16x2
  post_ref adr[SLOTS];      // where the edge's pair of posteriors lives
  uint32_t key0, key1;      // two smallest keys |v| * 32 + slot per half
  uint32_t sx;              // xor of the inputs: sign bits at 15 / 31
  uint32_t tinA, tinB;      // bytes {-clamp(+m0), -clamp(-m0), -clamp(+m1), -clamp(-m1)} of the PREVIOUS iteration
  uint32_t cwA[NW], cwB[NW];        // previous iteration's codes
  uint32_t ninA[NG], ninB[NG];      // minus stored message of slots 4g .. 4g+3, one byte each
  uint32_t ncwA[NW], ncwB[NW];      // codes being built
  post_ref post;

  template <int ROT> static T2_HD uint32_t table_in(uint32_t mw)
  {
    constexpr int P = (20 + ROT) & 31;                          // the minima sit behind the nibbles of the last code word
    const int m0c = LY::MPACK ? (int)((mw >> P) & 63u) : (int)(mw & 63u);
    const int m1c = LY::MPACK ? (int)((mw >> (P + 6)) & 63u) : (int)((mw >> 6) & 63u);
    return pack4(-(m0c < 31 ? m0c : 31), m0c, -(m1c < 31 ? m1c : 31), m1c);
  }
  T2_HD void begin(uint16_t* post_, const uint32_t (&wA)[LY::NS], const uint32_t (&wB)[LY::NS])
  {
    post = post_base(post_);
    tinA = table_in<0>(wA[LY::NS - 1]);
    tinB = table_in<16>(wB[LY::NS - 1]);
#pragma unroll
    for (int k = 0; k < NW; ++k) { cwA[k] = wA[k]; cwB[k] = wB[k]; ncwA[k] = 0; ncwB[k] = 0; }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      ninA[g] = prmt_rt(tinA, 0u, (g & 1) ? (cwA[g >> 1] >> 16) : cwA[g >> 1]);
      ninB[g] = prmt_rt(tinB, 0u, (g & 1) ? cwB[g >> 1] : (cwB[g >> 1] >> 16));
    }
    key0 = kIdleKey; key1 = kIdleKey; sx = 0;
  }
  // byte c of a (codeword A) and byte c of b (codeword B), each sign-extended into its half
  template <int C> static T2_HD uint32_t pick(uint32_t a, uint32_t b)
  {
    return prmt<(uint32_t)(C | ((8 | C) << 4) | ((4 + C) << 8) | ((12 + C) << 12))>(a, b);
  }
  template <int SLOT> T2_HD uint32_t stored_neg() const { return pick<SLOT & 3>(ninA[SLOT >> 2], ninB[SLOT >> 2]); }
  T2_HD uint32_t stored_neg_rt(int slot) const
  {
    const int sh4 = 4 * (slot & 7), sh4b = (sh4 + 16) & 31;
    uint32_t ca = (cwA[0] >> sh4) & 3u, cb = (cwB[0] >> sh4b) & 3u;
#pragma unroll
    for (int k = 1; k < NW; ++k) {
      const uint32_t xa = (cwA[k] >> sh4) & 3u, xb = (cwB[k] >> sh4b) & 3u;
      ca = (slot >> 3) == k ? xa : ca;
      cb = (slot >> 3) == k ? xb : cb;
    }
    const int a = (int)(int8_t)(tinA >> (8 * ca)), b = (int)(int8_t)(tinB >> (8 * cb));
    return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16);
  }
  // `active` false: the slot does not take part (it is computed and discarded, branch-free, so that the loads of a layer
  // still issue back to back)
  T2_HD void take(uint32_t v, int slot_const, bool active = true)
  {
    uint32_t key = abs2(v) * 32u + (uint32_t)slot_const * kOne2;             // |v| <= 128: no carry between the halves
    key = active ? key : kIdleKey;
    key1 = vmin2(key1, vmax2(key0, key));
    key0 = vmin2(key0, key);
    sx ^= active ? v : 0u;
  }
  template <int SLOT> T2_HD void edge_in(int a, bool active)
  {
    const post_ref r = post_at(post, a);
    const uint32_t pv = unpack_post(post_ld(r));
    const uint32_t v = sat8_add(pv, stored_neg<SLOT>());                      // vqsub(posterior, stored message)
    inp[SLOT] = v; adr[SLOT] = r;
    take(v, SLOT, active);
  }
  // m0 / m1 (after vqabs and the unsigned vqsub of beta = 1: both monotone, so applied to the two minima only) and the
  // arg-min slot of everything seen so far, per half
  T2_HD void minima(uint32_t& m0, uint32_t& m1, uint32_t& idn) const
  {
    const uint32_t c126 = 0x007e007eu, mone = 0xffffffffu;
    m0 = vmin2(vaddmax((key0 >> 5) & 0x07ff07ffu, mone, 0u), c126);
    m1 = vmin2(vaddmax((key1 >> 5) & 0x07ff07ffu, mone, 0u), c126);
    idn = key0 & 0x001f001fu;
  }
  template <int SLOT> T2_HD void mark_sign(bool active)
  {
    // sign of the product of the OTHER links: bit 15 (A) lands on its nibble, bit 31 (B) on the rotated one
    constexpr int pos = LY::template nib<0>(SLOT), posb = LY::template nib<16>(SLOT);
    const uint32_t u = active ? rotl(sx ^ inp[SLOT], (pos + 17) & 31) : 0u;
    ncwA[SLOT >> 3] |= u & (1u << pos);
    ncwB[SLOT >> 3] |= u & (1u << posb);
  }
  template <SlotMode MODE, int C>
  T2_HD void load_from(const uint16_t* eb, const uint16_t* es, int cnt, uint32_t mask, int j)
  {
    if constexpr (C < CNL) {
      const bool on = MODE == ALL_SLOTS ? true : (C < cnt && ((mask >> C) & 1u));
      if (!(MODE == BRANCHED && !on)) {
        edge_in<C>(mod360(j + (int)es[C]) + (int)eb[C], on);                   // eb + (j + es) mod 360
      }
      load_from<MODE, C + 1>(eb, es, cnt, mask, j);
    }
  }
  template <SlotMode MODE>
  T2_HD void load(const uint16_t* eb, const uint16_t* es, int cnt, uint32_t mask, int i, int j, int K, int q)
  {
    load_from<MODE, 0>(eb, es, cnt, mask, j);
    edge_in<CNL>(K + 360 * i + j, true);
    const bool hasB = (i | j) != 0;
    edge_in<CNL + 1>(i ? K + 360 * (i - 1) + j : K + 360 * (q - 1) + (hasB ? j - 1 : 0), hasB);
  }
  template <SlotMode MODE, int C> T2_HD void sign_from(int cnt, uint32_t mask)
  {
    if constexpr (C < CNL) {
      const bool on = MODE == ALL_SLOTS ? true : (C < cnt && ((mask >> C) & 1u));
      if (!(MODE == BRANCHED && !on)) mark_sign<C>(on);
      sign_from<MODE, C + 1>(cnt, mask);
    }
  }
  template <int SLOT> T2_HD void edge_out(const uint32_t (&noutA)[NG], const uint32_t (&noutB)[NG], bool active)
  {
    const uint32_t o = pick<SLOT & 3>(noutA[SLOT >> 2], noutB[SLOT >> 2]);     // other(mags[i], mins[0], mins[1]) with the sign
    if (active) post_st(adr[SLOT], pack_post(sat8_add(inp[SLOT], o)));          // vqadd
  }
  template <SlotMode MODE, int C>
  T2_HD void out_from(const uint32_t (&noutA)[NG], const uint32_t (&noutB)[NG], int cnt, uint32_t mask)
  {
    if constexpr (C < CNL) {
      const bool on = MODE == ALL_SLOTS ? true : (C < cnt && ((mask >> C) & 1u));
      if (!(MODE == BRANCHED && !on)) edge_out<C>(noutA, noutB, on);
      out_from<MODE, C + 1>(noutA, noutB, cnt, mask);
    }
  }
  // sign codes of the shared slots resolved earlier: bit c of negA / negB
  template <int C> T2_HD void merge_shared(uint32_t negA, uint32_t negB)
  {
    if constexpr (C < CNL) {
      if ((negA >> C) & 1u) ncwA[C >> 3] |= 1u << (4 * (C & 7));
      if ((negB >> C) & 1u) ncwB[C >> 3] |= 1u << LY::template nib<16>(C);
      merge_shared<C + 1>(negA, negB);
    }
  }
  template <int ROT>
  static T2_HD void finish_codes(uint32_t (&ncw)[NW], int idn, int m0, int m1, uint32_t (&nout)[NG], uint32_t (&w)[LY::NS])
  {
    const uint32_t bit = 2u << ((4 * (idn & 7) + ROT) & 31);
#pragma unroll
    for (int k = 0; k < NW; ++k) ncw[k] |= (idn >> 3) == k ? bit : 0u;
    const uint32_t tout = pack4(m0, -m0, m1, -m1);
#pragma unroll
    for (int g = 0; g < NG; ++g) nout[g] = prmt_rt(tout, 0u, ((g & 1) != (ROT != 0)) ? (ncw[g >> 1] >> 16) : ncw[g >> 1]);
    const uint32_t mm = (uint32_t)(m0 < 32 ? m0 : 32) | ((uint32_t)(m1 < 32 ? m1 : 32) << 6);
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = ncw[k];
    if (LY::MPACK) w[NW - 1] |= mm << ((20 + ROT) & 31); else w[LY::NS - 1] = mm;
  }
  // Write the private edges back and finish the check-node words.  negA / negB: bit c set when shared slot c (already
  // written by shared_out) carried a negative output sign in codeword A / B.
  template <SlotMode MODE>
  T2_HD void store(int cnt, uint32_t mask, int i, int j, uint32_t negA, uint32_t negB, uint32_t (&wA)[LY::NS], uint32_t (&wB)[LY::NS])
  {
    uint32_t m0, m1, idn;
    minima(m0, m1, idn);
    sign_from<MODE, 0>(cnt, mask);
    mark_sign<CNL>(true);
    mark_sign<CNL + 1>((i | j) != 0);
    if (MODE != ALL_SLOTS) merge_shared<0>(negA, negB);
    uint32_t noutA[NG], noutB[NG];
    finish_codes<0>(ncwA, (int)(idn & 0xffffu), (int)(m0 & 0xffffu), (int)(m1 & 0xffffu), noutA, wA);
    finish_codes<16>(ncwB, (int)(idn >> 16), (int)(m0 >> 16), (int)(m1 >> 16), noutB, wB);
    out_from<MODE, 0>(noutA, noutB, cnt, mask);
    edge_out<CNL>(noutA, noutB, true);
    edge_out<CNL + 1>(noutA, noutB, (i | j) != 0);
  }
  // ---- shared-edge path: the slot number is a run-time (warp-uniform) value ----
  T2_HD uint32_t shared_in(int slot, int a, uint32_t nbl)
  {
    const uint32_t v = sat8_add(unpack_post(post_ld(post_at(post, a))), nbl);
    take(v, slot);
    return v;
  }
  // returns bit 0 (A) / bit 16 (B) set when the output sign is negative
  T2_HD uint32_t shared_out(int slot, int a, uint32_t v, uint32_t m0, uint32_t m1, uint32_t idn)
  {
    return shared_out_at(slot, post_at(post, a), v, m0, m1, idn);
  }
  T2_HD uint32_t shared_out_at(int slot, post_ref where, uint32_t v, uint32_t m0, uint32_t m1, uint32_t idn)
  {
    uint32_t neg;
    const CnCore c = {key0, key1, sx};
    post_st(where, pack_post(core_out(c, slot, v, m0, m1, idn, neg)));
    return neg;
  }
  // the running state as a value / taken back (chain walk)
  T2_HD CnCore core() const { return CnCore{key0, key1, sx}; }
  T2_HD void set_core(const CnCore& c) { key0 = c.key0; key1 = c.key1; sx = c.sx; }
  T2_HD post_ref post_ref_at(int a) const { return post_at(post, a); }
  template <int C> T2_HD void shared_load_generic(const uint16_t* eb, const uint16_t* es, uint32_t mask, int j)
  {
    if constexpr (C < CNL) {
      if ((mask >> C) & 1u) {
        edge_in<C>(mod360(j + (int)es[C]) + (int)eb[C], true);
      }
      shared_load_generic<C + 1>(eb, es, mask, j);
    }
  }
  template <int C> T2_HD void shared_store_generic(uint32_t mask, uint32_t m0, uint32_t m1, uint32_t idn, uint32_t& negA, uint32_t& negB)
  {
    if constexpr (C < CNL) {
      if ((mask >> C) & 1u) {
        const uint32_t r = shared_out_at(C, adr[C], inp[C], m0, m1, idn);
        negA |= (r & 1u) << C; negB |= (r >> 16) << C;
      }
      shared_store_generic<C + 1>(mask, m0, m1, idn, negA, negB);
    }
  }
};

}  // namespace t2pair
