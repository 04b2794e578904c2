// TEST INFRASTRUCTURE: runs the check-node arithmetic of the CUDA LDPC decoder (csrc/ldpc_pair.h, the DPX / PRMT
// instructions restated in C) on the HOST, thread by thread and barrier phase by barrier phase exactly as
// ldpc_decode_kernel orders them, so that the arithmetic and the level schedule can be checked against the CPU oracle
// without a GPU (tests/test_ldpc_pair_emu.py).  A pair of codewords is decoded in lock step (both iterate until both pass
// the parity test or the trials run out: LDPC/layered_decoder.hh:168-180 with two lanes).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sdr_receiver_dvb_t2_b200/csrc/ldpc_pair.h"
#include "../../sdr_receiver_dvb_t2_b200/csrc/ldpc_schedule.h"

using namespace t2pair;

namespace {

// bad() of one lane (LDPC/layered_decoder.hh:65-82), plain restatement
bool lane_bad(const LdpcSchedule& s, const std::vector<uint16_t>& post, int lane, const std::vector<uint16_t>& eb,
              const std::vector<uint16_t>& es, int cnl)
{
  auto val = [&](int a) { return (int)(int8_t)(post[a] >> (8 * lane)); };
  for (int i = 0; i < s.q; ++i)
    for (int j = 0; j < 360; ++j) {
      int neg = 0;
      bool zero = false;
      auto use = [&](int a) { const int v = val(a); if (v == 0) zero = true; if (v < 0) neg ^= 1; };
      use(s.K + 360 * i + j);
      if (i) use(s.K + 360 * (i - 1) + j);
      else if (j) use(s.K + 360 * (s.q - 1) + j - 1);
      for (int c = 0; c < s.cnt[i]; ++c) {
        int t = j + es[i * cnl + c];
        if (t >= 360) t -= 360;
        use(t + eb[i * cnl + c]);
      }
      if (zero || neg) return true;
    }
  return false;
}

template <int CNL>
int decode_pair(const LdpcSchedule& s, const int8_t* llrA, const int8_t* llrB, int max_trials, int8_t* postA, int8_t* postB,
                int* iters_out)
{
  using LY = CnLayout<CNL>;
  constexpr int NS = LY::NS;
  std::vector<uint16_t> eb((size_t)s.q * CNL, 0), es((size_t)s.q * CNL, 0);
  for (int i = 0; i < s.q; ++i)
    for (int c = 0; c < s.cnl_max; ++c) {
      const uint32_t e = s.edge[(size_t)i * s.cnl_max + c];
      const int shift = e ? 360 - (int)(e >> 16) : 0;
      es[i * CNL + c] = (uint16_t)shift;
      eb[i * CNL + c] = (uint16_t)((e & 0xffffu) - shift);
    }
  std::vector<uint16_t> post(s.N);
  for (int n = 0; n < s.N; ++n) post[n] = (uint16_t)((uint8_t)llrA[n] | ((uint16_t)(uint8_t)llrB[n] << 8));
  std::vector<uint32_t> state((size_t)2 * NS * s.R, 0);
  int trials = max_trials, iters = 0;
  for (;;) {
    const bool bad = lane_bad(s, post, 0, eb, es, CNL) || lane_bad(s, post, 1, eb, es, CNL);
    if (!(bad && --trials >= 0)) break;
    for (int i = 0; i < s.q; ++i) {
      const int cnt = s.cnt[i], nl = s.nlev[i];
      const uint16_t* peb = eb.data() + i * CNL;
      const uint16_t* pes = es.data() + i * CNL;
      auto ld = [&](int tid, uint32_t (&wA)[NS], uint32_t (&wB)[NS]) {
        for (int k = 0; k < NS; ++k) {
          wA[k] = iters ? state[(size_t)k * s.R + i * 360 + tid] : 0u;
          wB[k] = iters ? state[(size_t)(NS + k) * s.R + i * 360 + tid] : 0u;
        }
      };
      auto st = [&](int tid, const uint32_t (&wA)[NS], const uint32_t (&wB)[NS]) {
        for (int k = 0; k < NS; ++k) {
          state[(size_t)k * s.R + i * 360 + tid] = wA[k];
          state[(size_t)(NS + k) * s.R + i * 360 + tid] = wB[k];
        }
      };
      if (nl == 1) {
        // one barrier phase: every thread reads all its inputs before any thread writes (emulated: two passes)
        std::vector<CheckNodePair<CNL>> cn(360);
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t wA[NS], wB[NS];
          ld(tid, wA, wB);
          cn[tid].begin(post.data(), wA, wB);
          if (cnt == CNL) cn[tid].template load<ALL_SLOTS>(peb, pes, cnt, ~0u, i, tid, s.K, s.q);
          else cn[tid].template load<PREDICATED>(peb, pes, cnt, ~0u, i, tid, s.K, s.q);
        }
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t wA[NS], wB[NS];
          if (cnt == CNL) cn[tid].template store<ALL_SLOTS>(cnt, ~0u, i, tid, 0u, 0u, wA, wB);
          else cn[tid].template store<PREDICATED>(cnt, ~0u, i, tid, 0u, 0u, wA, wB);
          st(tid, wA, wB);
        }
      } else {
        const uint32_t sh = s.shared[i];
        const uint8_t* level = s.level.data() + (size_t)s.conflict_index[i] * 360;
        std::vector<CheckNodePair<CNL>> cn(360);
        std::vector<uint32_t> negA(360, 0), negB(360, 0);
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t wA[NS], wB[NS];
          ld(tid, wA, wB);
          cn[tid].begin(post.data(), wA, wB);
          cn[tid].template load<PREDICATED>(peb, pes, cnt, ~sh, i, tid, s.K, s.q);
        }
        if (__builtin_popcount(sh) == 2) {
          // the chain walk of ldpc_decode_kernel, phase by phase (barriers between the loops)
          const int cA = __builtin_ffs(sh) - 1, cB = 31 - __builtin_clz(sh);
          int D = (int)pes[cA] - (int)pes[cB];
          D += D < 0 ? 360 : 0;
          const bool fwd = D <= 180;
          const int step = fwd ? D : 360 - D;
          const int slotO = fwd ? cA : cB, slotI = fwd ? cB : cA;
          const int esO = pes[slotO], ebO = peb[slotO], esI = pes[slotI], ebI = peb[slotI];
          std::vector<uint32_t> walk(360 * kWalkWords, 0), nI(360), nO(360), vI(360), vO(360), carry(360, 0);
          std::vector<post_ref> rI(360), rO(360);
          for (int tid = 0; tid < 360; ++tid) {
            nI[tid] = cn[tid].stored_neg_rt(slotI); nO[tid] = cn[tid].stored_neg_rt(slotO);
            rI[tid] = cn[tid].post_ref_at(mod360(tid + esI) + ebI); rO[tid] = cn[tid].post_ref_at(mod360(tid + esO) + ebO);
          }
          for (int tid = 0; tid < step; ++tid) {                                 // heads
            vI[tid] = sat8_add(unpack_post(post_ld(rI[tid])), nI[tid]);
            vO[tid] = sat8_add(unpack_post(post_ld(rO[tid])), nO[tid]);
            cn[tid].take(vI[tid], slotI); cn[tid].take(vO[tid], slotO);
            uint32_t m0, m1, idn, gI, gO;
            cn[tid].minima(m0, m1, idn);
            const CnCore c = cn[tid].core();
            post_st(rI[tid], pack_post(core_out(c, slotI, vI[tid], m0, m1, idn, gI)));
            carry[tid] = core_out(c, slotO, vO[tid], m0, m1, idn, gO);
            post_st(rO[tid], pack_post(carry[tid]));
            negA[tid] = ((gI & 1u) << slotI) | ((gO & 1u) << slotO);
            negB[tid] = ((gI >> 16) << slotI) | ((gO >> 16) << slotO);
          }
          for (int tid = step; tid < 360; ++tid) {                               // barrier; the others park what the walker needs
            vO[tid] = sat8_add(unpack_post(post_ld(rO[tid])), nO[tid]);
            uint32_t* st = walk.data() + tid * kWalkWords;
            const CnCore c = cn[tid].core();
            st[0] = (c.key0 >> 5) & 0x07ff07ffu; st[1] = c.sx; st[2] = vO[tid]; st[3] = nI[tid];
          }
          for (int tid = 0; tid < step; ++tid)                                   // barrier; the walk
            for (int j = tid + step; j < 360; j += step) {
              uint32_t* st = walk.data() + j * kWalkWords;
              const uint32_t w0 = st[0], w1 = st[1], w2 = st[2], w3 = st[3];
              st[0] = carry[tid];
              carry[tid] = walk_carry(carry[tid], w0, w1, w2, w3);
            }
          for (int tid = step; tid < 360; ++tid) {                               // barrier; everybody finishes its own check node
            vI[tid] = sat8_add(walk[tid * kWalkWords], nI[tid]);
            cn[tid].take(vI[tid], slotI); cn[tid].take(vO[tid], slotO);
            uint32_t m0, m1, idn, gI, gO;
            cn[tid].minima(m0, m1, idn);
            const CnCore c = cn[tid].core();
            post_st(rI[tid], pack_post(core_out(c, slotI, vI[tid], m0, m1, idn, gI)));
            const uint32_t pO = core_out(c, slotO, vO[tid], m0, m1, idn, gO);
            if (tid + step >= 360) post_st(rO[tid], pack_post(pO));
            negA[tid] = ((gI & 1u) << slotI) | ((gO & 1u) << slotO);
            negB[tid] = ((gI >> 16) << slotI) | ((gO >> 16) << slotO);
          }
        } else {
          for (int l = 1; l <= nl; ++l) {
            std::vector<uint32_t> m0(360), m1(360), idn(360);
            for (int tid = 0; tid < 360; ++tid)
              if (level[tid] == l) {
                cn[tid].template shared_load_generic<0>(peb, pes, sh, tid);
                cn[tid].minima(m0[tid], m1[tid], idn[tid]);
              }
            for (int tid = 0; tid < 360; ++tid)
              if (level[tid] == l) cn[tid].template shared_store_generic<0>(sh, m0[tid], m1[tid], idn[tid], negA[tid], negB[tid]);
          }
        }
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t wA[NS], wB[NS];
          cn[tid].template store<PREDICATED>(cnt, ~sh, i, tid, negA[tid], negB[tid], wA, wB);
          st(tid, wA, wB);
        }
      }
    }
    ++iters;
  }
  for (int n = 0; n < s.N; ++n) { postA[n] = (int8_t)(post[n] & 0xff); postB[n] = (int8_t)(post[n] >> 8); }
  *iters_out = iters;
  return trials;
}

}  // namespace

// llr: int8[2][N]; post_out: int8[2][N]; returns the reference's `trials` counter (< 0: not converged)
extern "C" int emu_ldpc_decode_pair(int code, const int8_t* llr, int max_trials, int8_t* post_out, int* iters)
{
  LdpcSchedule s;
  if (!t2_build_ldpc_schedule(code, s)) return -1000;
  const int8_t *a = llr, *b = llr + s.N;
  int8_t *pa = post_out, *pb = post_out + s.N;
  static const int buckets[] = {4, 5, 7, 8, 9, 11, 12, 13, 16, 17, 20};
  int cnl = 0;
  for (int x : buckets) if (x >= s.cnl_max) { cnl = x; break; }
  switch (cnl) {
#define CASE(C) case C: return decode_pair<C>(s, a, b, max_trials, pa, pb, iters);
    CASE(4) CASE(5) CASE(7) CASE(8) CASE(9) CASE(11) CASE(12) CASE(13) CASE(16) CASE(17) CASE(20)
    default: return -1001;
  }
}
