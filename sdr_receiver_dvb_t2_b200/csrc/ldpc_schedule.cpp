#include "ldpc_schedule.h"
#include <algorithm>
#include <array>

#include "ldpc_tables_data.inc"

static const int kKbch[15] = {32208, 38688, 43040, 48408, 51648, 53840,   // bch_decoder.cpp:79-105
                              7032, 9552, 10632, 11712, 12432, 13152,     // bch_decoder.cpp:107-131
                              0, 0, 0};

const T2LdpcCodeData* t2_ldpc_code_data(int code)
{
  if (code < 0 || code >= 15) return nullptr;
  return &kLdpcCodeData[code];
}
int t2_ldpc_k_bch(int code) { return (code < 0 || code >= 15) ? 0 : kKbch[code]; }

bool t2_build_ldpc_schedule(int code, LdpcSchedule& s)
{
  const T2LdpcCodeData* t = t2_ldpc_code_data(code);
  if (!t) return false;
  s.code = code; s.N = t->N; s.K = t->K; s.R = t->N - t->K; s.q = s.R / 360;
  const int q = s.q;
  struct E { int g, jx; };
  std::vector<std::vector<E>> layer(q);
  const uint16_t* a = t->addr;
  for (int g = 0; g < t->nrows; ++g) {
    for (int n = 0; n < t->rowdeg[g]; ++n) layer[a[n] % q].push_back({g, a[n] / q});
    a += t->rowdeg[g];
  }
  s.cnl_max = 0; s.links_total = 2 * s.R - 1;
  s.cnt.resize(q);
  for (int i = 0; i < q; ++i) {
    s.cnt[i] = (uint8_t)layer[i].size();
    s.cnl_max = std::max<int>(s.cnl_max, (int)layer[i].size());
    s.links_total += 360 * (int)layer[i].size();
  }
  s.edge.assign((size_t)q * s.cnl_max, 0);
  s.shared.assign(q, 0); s.ns.assign(q, 0); s.conflict_index.assign(q, -1); s.nlev.assign(q, 1); s.level.clear(); s.total_substeps = 0;
  for (int i = 0; i < q; ++i) {
    auto& L = layer[i];
    // Data edges that read a bit-group another edge of the layer reads too go FIRST (their slot numbers are then known
    // at compile time in the kernel).  With a single such pair the slots are ordered (I, O): bit (j + shift_O) of check node
    // j is bit (j' + shift_I) of check node j' = j + step, step = (shift_O - shift_I) mod 360 <= 180 -- the O bit is
    // handed to a LATER check node of the serial order (except at the wrap), the I bit comes from an earlier one.
    {
      std::vector<int> is_shared(L.size(), 0);
      for (size_t x = 0; x < L.size(); ++x)
        for (size_t y = x + 1; y < L.size(); ++y)
          if (L[x].g == L[y].g) { is_shared[x] = 1; is_shared[y] = 1; }
      std::vector<E> front, back;
      for (size_t x = 0; x < L.size(); ++x) (is_shared[x] ? front : back).push_back(L[x]);
      s.ns[i] = (uint8_t)front.size();
      if (front.size() == 2) {
        auto shift = [](const E& e) { return (360 - e.jx) % 360; };
        const int D = ((shift(front[0]) - shift(front[1])) % 360 + 360) % 360;      // front[0] as O, front[1] as I
        if (D <= 180) std::swap(front[0], front[1]);                                 // -> (I, O)
      }
      L = front;
      L.insert(L.end(), back.begin(), back.end());
    }
    for (size_t c = 0; c < L.size(); ++c)
    {
      const int sh = (360 - L[c].jx) % 360;
      s.edge[(size_t)i * s.cnl_max + c] = (uint32_t)(360 * L[c].g + sh) | ((uint32_t)(360 - sh) << 16);
    }
    // shared bits: two entries of the same bit-group in this layer
    std::array<std::vector<int>, 360> before;   // before[j] = check nodes that must run before j
    bool conflict = false;
    for (size_t x = 0; x < L.size(); ++x)
      for (size_t y = x + 1; y < L.size(); ++y)
        if (L[x].g == L[y].g) {
          conflict = true;
          s.shared[i] |= (1u << x) | (1u << y);
          for (int m = 0; m < 360; ++m) {
            int ja = (L[x].jx + m) % 360, jb = (L[y].jx + m) % 360;
            before[std::max(ja, jb)].push_back(std::min(ja, jb));
          }
        }
    if (conflict) {
      s.conflict_index[i] = (int16_t)(s.level.size() / 360);
      size_t off = s.level.size();
      s.level.resize(off + 360);
      int mx = 1;
      for (int j = 0; j < 360; ++j) {
        int lv = 1;
        for (int p : before[j]) lv = std::max<int>(lv, s.level[off + p] + 1);
        s.level[off + j] = (uint8_t)lv; mx = std::max(mx, lv);
      }
      s.nlev[i] = (uint8_t)mx;
    }
    s.total_substeps += s.nlev[i];
  }
  return true;
}
