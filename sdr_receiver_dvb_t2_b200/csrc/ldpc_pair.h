// Check-node arithmetic of the LDPC decoder for a PAIR of codewords: every 32-bit register holds the same quantity of
// two codewords as s16x2 halves (low half: codeword A, high half: codeword B), so that one DPX instruction of sm_100
// (VIADDMNMX.S16x2, VIMNMX.S16x2) or one PRMT / LOP3 serves both.  The arithmetic is the reference's
// (LDPC/layered_decoder.hh:87-107, LDPC/algorithms.hh:250-291: offset min-sum, beta = 1, int8 saturating, stored message
// clamped to [-32, 31]); int8 values live sign-extended in their 16-bit half and are clamped back to [-128, 127] wherever
// the reference saturates.
//
// Everything is kept in the form that costs the fewest ALU instructions per edge:
//   * magnitudes are tracked NEGATED (n = -|v| = min(v, -v): one DPX instruction after the complement), the two largest n
//     of a check node are its two smallest magnitudes;
//   * "other(mag, min0, min1)" is the reference's own compare BY VALUE (algorithms.hh:266): max(n + a0, -1) is 0 for an
//     input that attains the minimum and -1 (all ones) otherwise -- one DPX instruction gives the select mask of both
//     codewords, one LOP3 picks between the two candidate magnitudes, no arg-min index is kept;
//   * the sign of the outgoing message is the sign bit of (xor of all inputs) ^ input, turned into a half-wide mask by one
//     sign-replicating PRMT; a LOP3 picks +mag or -mag;
//   * the stored messages are kept as they are needed next time: negated and clamped, one byte per edge and codeword, two
//     edges (x two codewords) per 32-bit word of an L2-resident scratch -- reading one back is a single PRMT.
//
// The header compiles for the host too (the DPX / PRMT instructions are restated in plain C) so that
// tests/cpp/ldpc_pair_emu.cpp can run exactly this code, phase by phase, against the CPU oracle without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define T2_HD __host__ __device__ __forceinline__
#else
#define T2_HD inline
#endif

namespace t2pair {

#if defined(__CUDACC__)
__constant__ uint32_t kFmaMinus1 = 0xffffffffu;          // see not_fma()
__constant__ uint32_t kFmaOne = 1u;                       // see add_fma()
#endif

// ---- the handful of machine operations everything below is written in -------------------------------------------------
#if defined(__CUDA_ARCH__)
T2_HD uint32_t vaddmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }   // min(a + b, c) per half
T2_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }   // max(a + b, c) per half
T2_HD uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
T2_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
T2_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
// prmt.b32 with a selector known at compile time; selector nibbles with bit 3 set replicate the sign of the selected byte
template <uint32_t SEL> T2_HD uint32_t prmt(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(SEL));
  return r;
}
// Bitwise complement as a multiply-add, x * (-1) + (-1): issues on the FMA pipe -- the ALU pipe is the decoder's
// bottleneck.  The -1 comes from constant memory so that ptxas cannot turn the IMAD back into an ALU instruction.
__device__ __forceinline__ uint32_t not_fma(uint32_t x) { return x * kFmaMinus1 + kFmaMinus1; }
// ... and an addition as x * 1 + y, for the edge addresses
__device__ __forceinline__ uint32_t add_fma(uint32_t x, uint32_t y) { return x * kFmaOne + y; }
T2_HD int mod360(int t) { return (int)__viaddmin_u32((unsigned)t, 0xfffffe98u, (unsigned)t); }    // t in [0, 720): t mod 360
T2_HD uint32_t mod720(uint32_t t) { return __viaddmin_u32(t, 0xfffffd30u, t); }                       // t in [0, 1440): t mod 720
// the pair's posteriors by 32-bit shared-memory address (kept in a register from the load to the store of an edge)
typedef uint32_t post_ref;
T2_HD post_ref post_base(uint16_t* post) { return (uint32_t)__cvta_generic_to_shared(post); }
T2_HD post_ref post_at(post_ref base, int a) { return base + 2u * (uint32_t)a; }
// byte offset 720 g + x behind the base: two multiply-adds
T2_HD post_ref post_at_group(post_ref base, uint32_t g, uint32_t x) { return g * 720u + add_fma(x, base); }
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
T2_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
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
T2_HD uint32_t not_fma(uint32_t x) { return ~x; }
T2_HD uint32_t add_fma(uint32_t x, uint32_t y) { return x + y; }
T2_HD int mod360(int t) { return t >= 360 ? t - 360 : t; }
T2_HD uint32_t mod720(uint32_t t) { return t >= 720 ? t - 720 : t; }
typedef uint16_t* post_ref;
T2_HD post_ref post_base(uint16_t* post) { return post; }
T2_HD post_ref post_at(post_ref base, int a) { return base + a; }
T2_HD post_ref post_at_group(post_ref base, uint32_t g, uint32_t x) { return base + (g * 720u + x) / 2; }
T2_HD uint32_t post_ld(post_ref r) { return *r; }
T2_HD void post_st(post_ref r, uint32_t v) { *r = (uint16_t)v; }
#endif

constexpr uint32_t kP127 = 0x007f007fu, kM128 = 0xff80ff80u, kOne2 = 0x00010001u, kMinus1 = 0xffffffffu;
constexpr uint32_t kIdleN = 0x80008000u;          // "no input": -32768 in the negated-magnitude domain
constexpr uint32_t kC126 = 0x007e007eu, kM126 = 0xff82ff82u;

// int8 saturating add of two sign-extended pairs (vqadd / vqsub with a negated operand)
T2_HD uint32_t sat8_add(uint32_t a, uint32_t b) { return vmax2(vaddmin(a, b, kP127), kM128); }
// two posteriors (one per codeword) as they sit in shared memory (low byte A, high byte B) -> sign-extended s16x2
T2_HD uint32_t unpack_post(uint32_t raw16) { return prmt<0x9180u>(raw16, 0u); }
T2_HD uint32_t pack_post(uint32_t v) { return prmt<0x4420u>(v, 0u); }   // low bytes of the two halves, upper half zero
// half-wide masks (0xffff / 0) from the sign bits of the two halves
T2_HD uint32_t sign_mask2(uint32_t t) { return prmt<0xbb99u>(t, 0u); }
T2_HD uint32_t pick(uint32_t a, uint32_t b, uint32_t m) { return (a & ~m) | (b & m); }              // one LOP3: m ? b : a

// Stored messages of one check node of the pair: NSW words; word k holds slots 2k (bytes 0 / 1: codeword A / B) and
// 2k + 1 (bytes 2 / 3), each byte = -clamp(message, -32, 31).
template <int CNL> struct CnLayout {
  static constexpr int SLOTS = CNL + 2;                      // data edges + the two parity edges
  static constexpr int NSW = (SLOTS + 1) / 2;
};

// what the outgoing messages of a check node are made from: the smallest magnitude a0 (for the compare by value) and the
// two candidate messages after the offset and the clamp -- P1 for an input that attains the minimum, P0 = P1 ^ D for the
// others -- with the sign the message has when the input itself is positive (the product of ALL input signs), and their
// negatives N1 / N0 = N1 ^ ND for a negative input
struct CnOut { uint32_t a0, P1, D, N1, ND; };

// the raw candidates from the two largest negated magnitudes n0 >= n1 (<= 0; kIdleN when there is no second input)
struct CnMags { uint32_t A0, A1, NA0, NA1; };
T2_HD CnMags cn_mags(uint32_t n0, uint32_t n1)
{
  CnMags m;
  m.A0 = vmin2(vmax2(not_fma(n0), 0u), kC126);                              // vqsub(vqabs(v), 1), at most 126
  m.A1 = vmin2(vmax2(not_fma(n1), 0u), kC126);
  m.NA0 = vmin2(vaddmax(n0, kOne2, kM126), 0u);                             // -A0
  m.NA1 = vmin2(vaddmax(n1, kOne2, kM126), 0u);
  return m;
}
T2_HD CnOut cn_minima(uint32_t n0, uint32_t n1, uint32_t sx)
{
  const CnMags m = cn_mags(n0, n1);
  const uint32_t sg = sign_mask2(sx);
  CnOut o;
  o.a0 = vaddmax(not_fma(n0), kOne2, kIdleN);                               // -n0
  const uint32_t P0 = pick(m.A0, m.NA0, sg), N0 = pick(m.NA0, m.A0, sg);
  o.P1 = pick(m.A1, m.NA1, sg); o.N1 = pick(m.NA1, m.A1, sg);
  o.D = P0 ^ o.P1; o.ND = N0 ^ o.N1;
  return o;
}
// One edge out: input v (n = -|v|) -> new posterior pair (vqadd of v and the check node's message) and `nm`, what is
// stored for the next iteration: -clamp(message, -32, 31)
T2_HD uint32_t cn_edge_out(const CnOut& o, uint32_t v, uint32_t n, uint32_t& nm)
{
  const uint32_t mask = vaddmax(n, o.a0, kMinus1);           // 0: this input attains the minimum, -1: it does not
  const uint32_t P = o.P1 ^ (o.D & mask), N = o.N1 ^ (o.ND & mask);
  const uint32_t msg = pick(P, N, sign_mask2(v));            // sign = product of the OTHER inputs' signs
  nm = vmax2(vmin2(P ^ N ^ msg, 0x00200020u), 0xffe1ffe1u);
  return sat8_add(v, msg);
}
T2_HD uint32_t nabs2(uint32_t v) { return vaddmin(not_fma(v), kOne2, v); }                             // min(-v, v)

// The serial core of the chain walk (ldpc.cu): a check node shares one bit with its predecessor in the layer's serial order
// (slot 0: its updated posterior pair arrives in `carry`) and one with its successor (slot 1).  All the successor needs is
// the new posterior of the slot-1 bit: vqadd(vO, message) with |message| = the smallest magnitude among the OTHER inputs
// (the private edges, whose candidate A0p / -A0p is known beforehand, and the slot-0 input) and sign = product of the
// other inputs' signs (sxp: xor of the private inputs).  Nine dependent instructions per check node; everything else of
// the check node is done afterwards, in parallel, by its own thread.
T2_HD uint32_t walk_carry(uint32_t carry, uint32_t A0p, uint32_t NA0p, uint32_t sxp, uint32_t vO, uint32_t nI)
{
  const uint32_t vI = sat8_add(carry, nI);
  const uint32_t c = not_fma(vI);
  const uint32_t F = vmin2(A0p, vaddmax(vaddmax(c, kOne2, vI), kMinus1, 0u));        // min(A0p, max(|vI| - 1, 0))
  const uint32_t NF = vmax2(NA0p, vaddmin(vaddmin(c, kOne2, vI), kOne2, 0u));        // its negative
  return sat8_add(vO, pick(F, NF, sign_mask2(sxp ^ vI)));
}
constexpr int kWalkWords = 8;         // words parked per check node by the chain walk: A0p, -A0p, sxp, vO, nI, carry-in

// One check node of BOTH codewords of the pair.  Slots 0 .. CNL-1 are the data edges (those shared with another check node
// of the layer come first, ldpc_schedule.cpp), slot CNL the parity bit of the check node, slot CNL+1 the previous one.
template <int CNL>
struct CheckNodePair {
  using LY = CnLayout<CNL>;
  static constexpr int SLOTS = LY::SLOTS, NSW = LY::NSW;
  uint32_t v[SLOTS];        // input vqsub(posterior, stored message), s16x2; after the edge has been written: what is stored
  uint32_t n[SLOTS];        // -|v|
  post_ref adr[SLOTS];      // where the edge's pair of posteriors lives
  uint32_t w[NSW];          // stored messages of the previous iteration
  uint32_t n0, n1, sx;      // two largest n, xor of the inputs (sign bits at 15 / 31)
  post_ref post;

  T2_HD void begin(uint16_t* post_, const uint32_t (&w_)[NSW])
  {
    post = post_base(post_);
#pragma unroll
    for (int k = 0; k < NSW; ++k) w[k] = w_[k];
    n0 = kIdleN; n1 = kIdleN; sx = 0;
  }
  template <int SLOT> T2_HD uint32_t stored_neg() const
  {
    return (SLOT & 1) ? prmt<0xb3a2u>(w[SLOT >> 1], 0u) : prmt<0x9180u>(w[SLOT >> 1], 0u);
  }
  T2_HD void take(uint32_t vv, uint32_t nn, bool active = true)
  {
    nn = active ? nn : kIdleN;
    n1 = vmax2(n1, vmin2(n0, nn));
    n0 = vmax2(n0, nn);
    sx ^= active ? vv : 0u;
  }
  // two inputs at once: five min / max instructions instead of six
  T2_HD void take2(uint32_t va, uint32_t na, uint32_t vb, uint32_t nb)
  {
    const uint32_t hi = vmax2(na, nb), lo = vmin2(na, nb);
    n1 = vmax3(n1, vmin2(n0, hi), lo);
    n0 = vmax2(n0, hi);
    sx ^= va ^ vb;
  }
  // `active` false: the slot does not take part (it is computed and discarded, branch-free, so that the loads of a layer
  // still issue back to back)
  template <int SLOT> T2_HD void edge_read(post_ref r)
  {
    const uint32_t vv = sat8_add(unpack_post(post_ld(r)), stored_neg<SLOT>());         // vqsub(posterior, stored message)
    v[SLOT] = vv; n[SLOT] = nabs2(vv); adr[SLOT] = r;
  }
  template <int SLOT> T2_HD void edge_in(post_ref r, bool active)
  {
    edge_read<SLOT>(r);
    take(v[SLOT], n[SLOT], active);
  }
  // data slot C of check node j (j2 = 2 j): bit-group eg, cyclic shift es2 / 2 -> byte offset 720 eg + (2 j + es2) mod 720
  template <int C> T2_HD post_ref data_ref(const uint16_t* eg, const uint16_t* es2, uint32_t j2) const
  {
    return post_at_group(post, (uint32_t)eg[C], mod720(add_fma(j2, (uint32_t)es2[C])));
  }
  // input of a slot that arrives in a register (chain walk)
  template <int SLOT> T2_HD void edge_in_value(uint32_t vv)
  {
    const uint32_t nn = nabs2(vv);
    v[SLOT] = vv; n[SLOT] = nn;
    take(vv, nn);
  }
  template <int SLOT> T2_HD post_ref slot_ref(const uint16_t* eg, const uint16_t* es2, int j)
  {
    adr[SLOT] = data_ref<SLOT>(eg, es2, 2u * (uint32_t)j);
    return adr[SLOT];
  }
  T2_HD CnOut minima() const { return cn_minima(n0, n1, sx); }
  T2_HD CnMags mags() const { return cn_mags(n0, n1); }
  // write the edge back (when `store`) and leave what is stored for the next iteration in v[SLOT]
  template <int SLOT> T2_HD uint32_t edge_out(const CnOut& o, bool active, bool store)
  {
    uint32_t nm;
    const uint32_t np = cn_edge_out(o, v[SLOT], n[SLOT], nm);
    if (active && store) post_st(adr[SLOT], pack_post(np));
    v[SLOT] = active ? nm : 0u;
    return np;
  }
  // data slots lo <= C < hi (ALL: every slot, known at compile time): all the reads first, then the running minima two
  // inputs at a time
  template <bool ALL, int C> T2_HD void read_from(const uint16_t* eg, const uint16_t* es2, uint32_t j2)
  {
    if constexpr (C < CNL) {
      edge_read<C>(data_ref<C>(eg, es2, j2));
      read_from<ALL, C + 1>(eg, es2, j2);
    }
  }
  template <bool ALL, int C> T2_HD void reduce_from(int lo, int hi, bool lastB)
  {
    if constexpr (C + 1 < SLOTS) {
      if (ALL && C + 1 < CNL) take2(v[C], n[C], v[C + 1], n[C + 1]);
      else {
        take(v[C], n[C], C >= CNL ? true : (ALL || (C >= lo && C < hi)));
        take(v[C + 1], n[C + 1], C + 1 == CNL + 1 ? lastB : C + 1 == CNL ? true : (ALL || (C + 1 >= lo && C + 1 < hi)));
      }
      reduce_from<ALL, C + 2>(lo, hi, lastB);
    } else if constexpr (C < SLOTS) {
      take(v[C], n[C], C == CNL + 1 ? lastB : C == CNL ? true : (ALL || (C >= lo && C < hi)));
    }
  }
  // the private edges: data slots lo .. cnt-1 and the two parity slots (ALL: lo = 0 and cnt = CNL)
  template <bool ALL>
  T2_HD void load(const uint16_t* eg, const uint16_t* es2, int lo, int cnt, int i, int j, int K, int q)
  {
    read_from<ALL, 0>(eg, es2, 2u * (uint32_t)j);
    edge_read<CNL>(post_at(post, K + 360 * i + j));
    const bool hasB = (i | j) != 0;
    edge_read<CNL + 1>(post_at(post, i ? K + 360 * (i - 1) + j : K + 360 * (q - 1) + (hasB ? j - 1 : 0)));
    reduce_from<ALL, 0>(lo, cnt, hasB);
  }
  template <bool ALL, int C> T2_HD void out_from(const CnOut& o, int lo, int hi)
  {
    if constexpr (C < CNL) {
      const bool on = ALL || (C >= lo && C < hi);
      if (ALL || C >= lo) edge_out<C>(o, on, true);                    // (slots below lo were written earlier: v[C] holds their stored value)
      out_from<ALL, C + 1>(o, lo, hi);
    }
  }
  // Write the private edges back and pack what is stored.
  template <bool ALL>
  T2_HD void store(int lo, int cnt, int i, int j, uint32_t (&w_)[NSW])
  {
    const CnOut o = minima();
    out_from<ALL, 0>(o, lo, cnt);
    edge_out<CNL>(o, true, true);
    edge_out<CNL + 1>(o, (i | j) != 0, true);
    pack(w_);
  }
  T2_HD void pack(uint32_t (&w_)[NSW]) const
  {
#pragma unroll
    for (int k = 0; k < NSW; ++k) w_[k] = prmt<0x6420u>(v[2 * k], 2 * k + 1 < SLOTS ? v[2 * k + 1] : 0u);
  }
  // ---- shared slots 0 .. ns-1 of a layer resolved level by level ----
  template <int C> T2_HD void shared_load(const uint16_t* eg, const uint16_t* es2, int ns, int j)
  {
    if constexpr (C < CNL && C < 10) {
      if (C < ns) {
        edge_in<C>(data_ref<C>(eg, es2, 2u * (uint32_t)j), true);
        shared_load<C + 1>(eg, es2, ns, j);
      }
    }
  }
  template <int C> T2_HD void shared_store(const CnOut& o, int ns)
  {
    if constexpr (C < CNL && C < 10) {
      if (C < ns) {
        edge_out<C>(o, true, true);
        shared_store<C + 1>(o, ns);
      }
    }
  }
};

}  // namespace t2pair
