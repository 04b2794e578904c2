/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's LDPC / "BCH" stage.
 * Nothing under sdr_receiver_dvb_t2_b200/ may link, import or call this file.  It exists so the
 * -m gpu parity tests can run on the GPU box, where /root/reference is absent; it is pinned
 * against the compiled, unmodified reference (oracle/_ref/libref_ldpc.so) by
 * tests/test_oracle_ldpc.py in the build container and against tests/golden/ldpc_*.npz
 * (vectors produced by that reference, generator: tools/make_golden_ldpc.py) everywhere.
 *
 * Follows (paths relative to /root/reference/src/DVB_T2):
 *   LDPC/ldpc.hh:39-123              bit -> check-node enumeration of the quasi-cyclic tables
 *   LDPC/layered_decoder.hh:115-167  init(): pos[] build and the (q*j+i) -> (360*i+j) re-order
 *   LDPC/layered_decoder.hh:65-82    bad()
 *   LDPC/layered_decoder.hh:83-110   update()
 *   LDPC/layered_decoder.hh:168-180  operator(): reset, parity re-order, trial loop
 *   LDPC/algorithms.hh:221-292       OffsetMinSumAlgorithm<SIMD<int8_t,W>,NormalUpdate,2>
 *   LDPC/avx2.hh                     saturating int8 add/sub, qabs, sign (as _mm256_*_epi8 define them)
 *   ldpc_decoder.cpp:248-277         lane transposition (cancels against operator()'s), hard decision
 *   bch_decoder.cpp:50-61,139-142    BB-scrambler PRBS, strip BCH parity + XOR
 *
 * The 32 SIMD lanes of the reference are independent except for the termination test, so the
 * port keeps one scalar state per lane and shares only the trial loop.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const char* name; int N, K, nrows; const uint8_t* rowdeg; const uint16_t* addr; } T2LdpcCodeData;
#include "../../sdr_receiver_dvb_t2_b200/csrc/ldpc_tables_data.inc"

#define M360 360

typedef struct {
  int N, K, R, q, CNL, LT;
  uint16_t* pos;   /* [R*CNL], CN (i,j) at row 360*i+j          (layered_decoder.hh:157-161) */
  uint8_t* cnc;    /* [R]  data edges of check q*j+i, indexed by ORIGINAL check number       */
} port_code;

static port_code g_codes[15];
static int g_init[15];

static const port_code* get_code(int id)
{
  if (g_init[id]) return &g_codes[id];
  const T2LdpcCodeData* t = &kLdpcCodeData[id];
  port_code* c = &g_codes[id];
  c->N = t->N; c->K = t->K; c->R = t->N - t->K; c->q = c->R / M360;
  /* LINKS_MAX_CN - 2: the largest number of data edges on any check node */
  uint8_t* cnt = (uint8_t*)calloc(c->R, 1);
  {
    const uint16_t* a = t->addr;
    for (int r = 0; r < t->nrows; ++r) {
      for (int m = 0; m < M360; ++m)
        for (int n = 0; n < t->rowdeg[r]; ++n) cnt[(a[n] + c->q * m) % c->R]++;
      a += t->rowdeg[r];
    }
  }
  int cnl = 0, lt = 0;
  for (int i = 0; i < c->R; ++i) { if (cnt[i] > cnl) cnl = cnt[i]; lt += cnt[i]; }
  c->CNL = cnl; c->LT = lt + 2 * c->R - 1;
  uint16_t* pos = (uint16_t*)calloc((size_t)c->R * cnl, sizeof(uint16_t));
  memset(cnt, 0, c->R);
  {
    const uint16_t* a = t->addr; int bit = 0;
    for (int r = 0; r < t->nrows; ++r) {               /* ldpc.hh first_bit/next_bit */
      for (int m = 0; m < M360; ++m, ++bit)
        for (int n = 0; n < t->rowdeg[r]; ++n) {
          int chk = (a[n] + c->q * m) % c->R;
          pos[cnl * chk + cnt[chk]++] = (uint16_t)bit; /* layered_decoder.hh:140-149 */
        }
      a += t->rowdeg[r];
    }
  }
  c->pos = (uint16_t*)malloc((size_t)c->R * cnl * sizeof(uint16_t));
  for (int i = 0; i < c->q; ++i)
    for (int j = 0; j < M360; ++j)
      for (int k = 0; k < cnl; ++k)
        c->pos[cnl * (M360 * i + j) + k] = pos[cnl * (c->q * j + i) + k];
  free(pos);
  c->cnc = cnt;
  g_init[id] = 1;
  return c;
}

/* avx2.hh: _mm256_adds_epi8 / _mm256_subs_epi8 */
static inline int8_t sat8(int v) { return (int8_t)(v > 127 ? 127 : (v < -128 ? -128 : v)); }
/* avx2.hh vqabs: abs with -128 -> 127 */
static inline int qabs8(int8_t v) { int a = v < 0 ? -v : v; return a > 127 ? 127 : a; }
/* _mm256_sign_epi8(a,b) */
static inline int8_t sign8(int8_t a, int8_t b) { return b < 0 ? (int8_t)(-a) : (b == 0 ? 0 : a); }

/* algorithms.hh:250-271 finalp, scalar */
static void finalp(int8_t* links, int cnt)
{
  int mags[32];
  for (int i = 0; i < cnt; ++i) { int m = qabs8(links[i]) - 1; mags[i] = m < 0 ? 0 : m; }   /* beta = nearbyint(0.5*2) = 1 */
  int min0 = mags[0] < mags[1] ? mags[0] : mags[1];
  int min1 = mags[0] < mags[1] ? mags[1] : mags[0];
  for (int i = 2; i < cnt; ++i) {
    int mx = min0 > mags[i] ? min0 : mags[i];
    if (mx < min1) min1 = mx;
    if (mags[i] < min0) min0 = mags[i];
  }
  uint8_t signs = (uint8_t)links[0];
  for (int i = 1; i < cnt; ++i) signs ^= (uint8_t)links[i];
  for (int i = 0; i < cnt; ++i) {
    int8_t other = (int8_t)(mags[i] == min0 ? min1 : min0);
    int8_t sg = (int8_t)((signs ^ (uint8_t)links[i]) | 127);
    links[i] = sign8(other, sg);
  }
}

typedef struct { int8_t* post; int8_t* bnl; } lane_state;

/* layered_decoder.hh:65-82 for one lane: 1 = some check is not strictly satisfied */
static int lane_bad(const port_code* c, const int8_t* P)
{
  const int8_t* data = P; const int8_t* parity = P + c->K;
  for (int i = 0; i < c->q; ++i) {
    int cnt = c->cnc[i];
    for (int j = 0; j < M360; ++j) {
      int8_t cnv = sign8(1, parity[M360 * i + j]);
      if (i) cnv = sign8(cnv, parity[M360 * (i - 1) + j]);
      else if (j) cnv = sign8(cnv, parity[j + (c->q - 1) * M360 - 1]);
      for (int k = 0; k < cnt; ++k) cnv = sign8(cnv, data[c->pos[c->CNL * (M360 * i + j) + k]]);
      if (!(cnv > 0)) return 1;
    }
  }
  return 0;
}

/* layered_decoder.hh:83-110 for one lane */
static void lane_update(const port_code* c, int8_t* P, int8_t* bnl)
{
  int8_t* data = P; int8_t* parity = P + c->K; int8_t* bl = bnl;
  for (int i = 0; i < c->q; ++i) {
    int cnt = c->cnc[i];
    for (int j = 0; j < M360; ++j) {
      int deg = cnt + 2 - !(i | j);
      int8_t inp[32], out[32];
      const uint16_t* ps = &c->pos[c->CNL * (M360 * i + j)];
      for (int k = 0; k < cnt; ++k) inp[k] = out[k] = sat8(data[ps[k]] - bl[k]);
      inp[cnt] = out[cnt] = sat8(parity[M360 * i + j] - bl[cnt]);
      if (i) inp[cnt + 1] = out[cnt + 1] = sat8(parity[M360 * (i - 1) + j] - bl[cnt + 1]);
      else if (j) inp[cnt + 1] = out[cnt + 1] = sat8(parity[j + (c->q - 1) * M360 - 1] - bl[cnt + 1]);
      finalp(out, deg);
      for (int k = 0; k < cnt; ++k) data[ps[k]] = sat8(inp[k] + out[k]);
      parity[M360 * i + j] = sat8(inp[cnt] + out[cnt]);
      if (i) parity[M360 * (i - 1) + j] = sat8(inp[cnt + 1] + out[cnt + 1]);
      else if (j) parity[j + (c->q - 1) * M360 - 1] = sat8(inp[cnt + 1] + out[cnt + 1]);
      for (int d = 0; d < deg; ++d) {                 /* algorithms.hh:288-291: store clamp(out,-32,31) */
        int v = out[d]; *bl++ = (int8_t)(v < -32 ? -32 : (v > 31 ? 31 : v));
      }
    }
  }
}

int port_ldpc_code_n(int code) { return kLdpcCodeData[code].N; }
int port_ldpc_code_k(int code) { return kLdpcCodeData[code].K; }
int port_ldpc_code_cnl(int code) { return get_code(code)->CNL; }
int port_ldpc_code_links(int code) { return get_code(code)->LT; }

/*
 * One reference batch: `lanes` codewords (the reference always uses 32) decoded in lock-step:
 * iterate while ANY lane fails the check and trials remain (layered_decoder.hh:174).
 *   llr       int8[lanes][N] codeword order;  bits_out uint8[lanes][K] byte per bit (or NULL)
 *   post_out  int8[lanes][N] posteriors after the run (or NULL)
 * returns trials left (<0: the reference drops the whole batch, ldpc_decoder.cpp:264-268).
 */
int port_ldpc_decode_group(int code, const int8_t* llr, int lanes, uint8_t* bits_out, int8_t* post_out, int trials)
{
  const port_code* c = get_code(code);
  int8_t* P = (int8_t*)malloc((size_t)lanes * c->N);
  int8_t* B = (int8_t*)calloc((size_t)lanes * c->LT, 1);     /* reset(): all messages zero */
  memcpy(P, llr, (size_t)lanes * c->N);
  for (;;) {
    int bad = 0;
    for (int l = 0; l < lanes && !bad; ++l) bad = lane_bad(c, P + (size_t)l * c->N);
    if (!(bad && --trials >= 0)) break;
    for (int l = 0; l < lanes; ++l) lane_update(c, P + (size_t)l * c->N, B + (size_t)l * c->LT);
  }
  if (bits_out)
    for (int l = 0; l < lanes; ++l)
      for (int i = 0; i < c->K; ++i) bits_out[(size_t)l * c->K + i] = P[(size_t)l * c->N + i] < 0;
  if (post_out) memcpy(post_out, P, (size_t)lanes * c->N);
  free(P); free(B);
  return trials;
}

/* Per-lane diagnostics used by tests: is lane's word a strictly satisfied codeword? */
int port_ldpc_bad(int code, const int8_t* post) { return lane_bad(get_code(code), post); }

/* bch_decoder.cpp:50-61 */
void port_bb_prbs(uint8_t* prbs, int n)
{
  int sr = 0x4A80;
  for (int i = 0; i < n; i++) {
    uint8_t b = (uint8_t)(((sr) ^ (sr >> 1)) & 1);
    prbs[i] = b;
    sr >>= 1;
    if (b) sr |= 0x4000;
  }
}

/* bch_decoder.cpp:139-142: keep the first k_bch bits of each K_ldpc-bit word, XOR the PRBS */
void port_bch_strip_descramble(const uint8_t* in, int n_words, int k_ldpc, int k_bch, uint8_t* out)
{
  static uint8_t prbs[54000]; static int init = 0;
  if (!init) { port_bb_prbs(prbs, 54000); init = 1; }
  for (int w = 0; w < n_words; ++w)
    for (int i = 0; i < k_bch; ++i) out[(size_t)w * k_bch + i] = in[(size_t)w * k_ldpc + i] ^ prbs[i];
}

/*
 * Systematic IRA encoder (test-only; the reference has no transmitter).  EN 302 755 6.1.2:
 * accumulate information bits into parity addresses, then p[i] ^= p[i-1].  Parity is emitted in
 * the order the reference decoder consumes it (transmitted order).
 */
void port_ldpc_encode(int code, const uint8_t* info, uint8_t* cw)
{
  const T2LdpcCodeData* t = &kLdpcCodeData[code];
  int N = t->N, K = t->K, R = N - K, q = R / M360;
  uint8_t* p = (uint8_t*)calloc(R, 1);
  const uint16_t* a = t->addr; int bit = 0;
  for (int r = 0; r < t->nrows; ++r) {
    for (int m = 0; m < M360; ++m, ++bit)
      if (info[bit])
        for (int n = 0; n < t->rowdeg[r]; ++n) p[(a[n] + q * m) % R] ^= 1;
    a += t->rowdeg[r];
  }
  for (int i = 1; i < R; ++i) p[i] ^= p[i - 1];
  memcpy(cw, info, K);
  /* parity interleaver: u[K + 360*t + s] = p[q*s + t] */
  for (int s = 0; s < M360; ++s)
    for (int tt = 0; tt < q; ++tt) cw[K + M360 * tt + s] = p[q * s + tt];
  free(p);
}
