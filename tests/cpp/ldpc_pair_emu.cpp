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
        int t = j + es[i * cnl + c] / 2;
        if (t >= 360) t -= 360;
        use(t + 360 * eb[i * cnl + c]);
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
  constexpr int NSW = LY::NSW;
  std::vector<uint16_t> eb((size_t)s.q * CNL, 0), es((size_t)s.q * CNL, 0);
  for (int i = 0; i < s.q; ++i)
    for (int c = 0; c < s.cnl_max; ++c) {
      const uint32_t e = s.edge[(size_t)i * s.cnl_max + c];
      const int shift = e ? 360 - (int)(e >> 16) : 0;
      es[i * CNL + c] = (uint16_t)(2 * shift);                                  // the kernel's tables: 2 x shift, bit-group
      eb[i * CNL + c] = (uint16_t)(((e & 0xffffu) - shift) / 360);
    }
  std::vector<uint16_t> post(s.N);
  for (int n = 0; n < s.N; ++n) post[n] = (uint16_t)((uint8_t)llrA[n] | ((uint16_t)(uint8_t)llrB[n] << 8));
  std::vector<uint32_t> state((size_t)s.q * NSW * 360, 0);
  int trials = max_trials, iters = 0;
  for (;;) {
    const bool bad = lane_bad(s, post, 0, eb, es, CNL) || lane_bad(s, post, 1, eb, es, CNL);
    if (!(bad && --trials >= 0)) break;
    for (int i = 0; i < s.q; ++i) {
      const int cnt = s.cnt[i], nl = s.nlev[i];
      const uint16_t* peb = eb.data() + i * CNL;
      const uint16_t* pes = es.data() + i * CNL;
      uint32_t* sp = state.data() + (size_t)i * NSW * 360;
      auto ld = [&](int tid, uint32_t (&w)[NSW]) { for (int k = 0; k < NSW; ++k) w[k] = iters ? sp[k * 360 + tid] : 0u; };
      auto st = [&](int tid, const uint32_t (&w)[NSW]) { for (int k = 0; k < NSW; ++k) sp[k * 360 + tid] = w[k]; };
      std::vector<CheckNodePair<CNL>> cn(360);
      if (nl == 1) {
        // one barrier phase: a thread's bits are touched by no other thread of the layer
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t w[NSW];
          ld(tid, w);
          cn[tid].begin(post.data(), w);
          if (cnt == CNL) {
            cn[tid].template load<true>(peb, pes, 0, cnt, i, tid, s.K, s.q);
            cn[tid].template store<true>(0, cnt, i, tid, w);
          } else {
            cn[tid].template load<false>(peb, pes, 0, cnt, i, tid, s.K, s.q);
            cn[tid].template store<false>(0, cnt, i, tid, w);
          }
          st(tid, w);
        }
      } else {
        const int ns = s.ns[i];
        const uint8_t* level = s.level.data() + (size_t)s.conflict_index[i] * 360;
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t w[NSW];
          ld(tid, w);
          cn[tid].begin(post.data(), w);
          cn[tid].template load<false>(peb, pes, ns, cnt, i, tid, s.K, s.q);
        }
        if (ns == 2) {
          // the chain walk of ldpc_decode_kernel, phase by phase (barriers between the loops)
          const int step = mod360(((int)pes[1] >> 1) + 360 - ((int)pes[0] >> 1));
          std::vector<uint32_t> walk(360 * kWalkWords, 0), nI(360), vO(360), carry(360, 0);
          std::vector<post_ref> rI(360), rO(360);
          for (int tid = 0; tid < 360; ++tid) {
            nI[tid] = cn[tid].template stored_neg<0>();
            rI[tid] = cn[tid].template slot_ref<0>(peb, pes, tid); rO[tid] = cn[tid].template slot_ref<1>(peb, pes, tid);
          }
          for (int tid = 0; tid < step; ++tid) {                                 // heads
            cn[tid].template edge_in_value<0>(sat8_add(unpack_post(post_ld(rI[tid])), nI[tid]));
            cn[tid].template edge_in_value<1>(sat8_add(unpack_post(post_ld(rO[tid])), cn[tid].template stored_neg<1>()));
            const CnOut o = cn[tid].minima();
            cn[tid].template edge_out<0>(o, true, true);
            carry[tid] = cn[tid].template edge_out<1>(o, true, true);
          }
          for (int tid = step; tid < 360; ++tid) {                               // barrier; the others park what the walker needs
            vO[tid] = sat8_add(unpack_post(post_ld(rO[tid])), cn[tid].template stored_neg<1>());
            const CnMags m = cn[tid].mags();
            uint32_t* w = walk.data() + tid * kWalkWords;
            w[0] = m.A0; w[1] = m.NA0; w[2] = cn[tid].sx; w[3] = vO[tid]; w[4] = nI[tid];
          }
          for (int tid = 0; tid < step; ++tid)                                   // barrier; the walk
            for (int j = tid + step; j < 360; j += step) {
              uint32_t* w = walk.data() + j * kWalkWords;
              w[5] = carry[tid];
              carry[tid] = walk_carry(carry[tid], w[0], w[1], w[2], w[3], w[4]);
            }
          for (int tid = step; tid < 360; ++tid) {                               // barrier; everybody finishes its own check node
            cn[tid].template edge_in_value<0>(sat8_add(walk[tid * kWalkWords + 5], nI[tid]));
            cn[tid].template edge_in_value<1>(vO[tid]);
            const CnOut o = cn[tid].minima();
            cn[tid].template edge_out<0>(o, true, true);
            cn[tid].template edge_out<1>(o, true, tid + step >= 360);
          }
        } else {
          for (int l = 1; l <= nl; ++l) {
            // threads of one level: all reads, then all writes (they do not share bits with each other)
            std::vector<CnOut> o(360);
            for (int tid = 0; tid < 360; ++tid)
              if (level[tid] == l) {
                cn[tid].template shared_load<0>(peb, pes, ns, tid);
                o[tid] = cn[tid].minima();
              }
            for (int tid = 0; tid < 360; ++tid)
              if (level[tid] == l) cn[tid].template shared_store<0>(o[tid], ns);
          }
        }
        for (int tid = 0; tid < 360; ++tid) {
          uint32_t w[NSW];
          cn[tid].template store<false>(ns, cnt, i, tid, w);
          st(tid, w);
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
