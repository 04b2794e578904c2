/*
 * TEST INFRASTRUCTURE ONLY -- plain-C BCH encoder / decoder for the DVB-T2 outer code (EN 302 755 clause 6.1.1, tables 6a / 6b):
 * the checker for the GPU's opt-in t2b200_bch_decode.  The reference itself stops at "TODO BCH decode"
 * (bch_decoder.cpp:136: it strips the parity bits and never looks at them), so there is no reference behaviour to pin here:
 * the pins are the standard's generator polynomials (g_1 below is the primitive polynomial of table 6a / 6b; g_2 .. g_12 are
 * the minimal polynomials of alpha^3 .. alpha^23, which reproduces the table rows), the code's algebra (every codeword has
 * zero syndromes; up to t errors are corrected; tests/test_bch.py), and agreement of the two implementations.
 *
 * Geometry: 64 800-bit FECFRAMEs use GF(2^16), N_bch = K_ldpc, t = 12 (rates 1/2, 3/5, 3/4, 4/5) or t = 10 (2/3, 5/6);
 * 16 200-bit FECFRAMEs use GF(2^14), t = 12.  Bit 0 of a word is the coefficient of x^(N_bch - 1).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int m, n, t, ready;            /* field GF(2^m), n = 2^m - 1 */
  uint16_t* exp;                 /* alpha^i, i < 2n */
  uint16_t* log;
  uint8_t* gen;                  /* generator polynomial, gen[i] = coefficient of x^i, degree m*t */
  int deg;
} bch_field;

static bch_field g_f[2][13];     /* [short?][t] */

static const bch_field* field_for(int short_frame, int t)
{
  bch_field* f = &g_f[short_frame ? 1 : 0][t];
  if (f->ready) return f;
  f->m = short_frame ? 14 : 16;
  f->n = (1 << f->m) - 1;
  f->t = t;
  const int prim = short_frame ? ((1 << 14) | (1 << 5) | (1 << 3) | (1 << 1) | 1)       /* 1 + x + x^3 + x^5 + x^14 */
                               : ((1 << 16) | (1 << 5) | (1 << 3) | (1 << 2) | 1);      /* 1 + x^2 + x^3 + x^5 + x^16 */
  f->exp = (uint16_t*)malloc(sizeof(uint16_t) * 2 * (f->n + 1));
  f->log = (uint16_t*)calloc((size_t)f->n + 1, sizeof(uint16_t));
  int x = 1;
  for (int i = 0; i < f->n; ++i) {
    f->exp[i] = (uint16_t)x; f->log[x] = (uint16_t)i;
    x <<= 1;
    if (x >> f->m) x ^= prim;
  }
  for (int i = f->n; i < 2 * (f->n + 1); ++i) f->exp[i] = f->exp[i - f->n];
  /* g(x) = product over i = 1..t of the minimal polynomial of alpha^(2i-1) */
  f->gen = (uint8_t*)calloc((size_t)f->m * t + 1, 1);
  f->gen[0] = 1; f->deg = 0;
  for (int i = 1; i <= t; ++i) {
    /* minimal polynomial: product over the conjugates alpha^(e 2^k) of (x + alpha^c), coefficients end up in GF(2) */
    uint16_t mp[17]; int md = 0;
    memset(mp, 0, sizeof(mp)); mp[0] = 1;
    int e = (2 * i - 1) % f->n, c = e;
    do {
      const uint16_t a = f->exp[c];
      for (int k = md + 1; k >= 1; --k) {
        uint16_t v = mp[k - 1];
        if (k <= md && mp[k]) v ^= f->exp[(f->log[mp[k]] + f->log[a]) % f->n];
        mp[k] = v;
      }
      mp[0] = mp[0] ? f->exp[(f->log[mp[0]] + f->log[a]) % f->n] : 0;
      ++md;
      c = (c * 2) % f->n;
    } while (c != e);
    /* gen *= mp (both over GF(2) now) */
    uint8_t* ng = (uint8_t*)calloc((size_t)f->deg + md + 1, 1);
    for (int a = 0; a <= f->deg; ++a)
      if (f->gen[a]) for (int b = 0; b <= md; ++b) ng[a + b] ^= (uint8_t)(mp[b] & 1);
    memcpy(f->gen, ng, (size_t)f->deg + md + 1);
    f->deg += md;
    free(ng);
  }
  f->ready = 1;
  return f;
}

/* t and field of an LDPC code id (include/t2b200.h numbering: 0..5 normal 1/2 .. 5/6, 6..11 short) */
int port_bch_t(int code) { return code < 0 || code > 11 ? 0 : (code == 2 || code == 5) ? 10 : 12; }
int port_bch_parity_bits(int code) { return code < 0 || code > 11 ? 0 : (code >= 6 ? 14 : 16) * port_bch_t(code); }

/* word: uint8[n_bch], one byte per bit; the last deg bits are overwritten with the parity of the first n_bch - deg */
void port_bch_encode(int code, uint8_t* word, int n_bch)
{
  const bch_field* f = field_for(code >= 6, port_bch_t(code));
  const int k = n_bch - f->deg;
  uint8_t* reg = (uint8_t*)calloc((size_t)f->deg, 1);           /* reg[i] = coefficient of x^i of the running remainder */
  for (int i = 0; i < k; ++i) {
    const uint8_t fb = (uint8_t)((word[i] & 1) ^ reg[f->deg - 1]);
    for (int j = f->deg - 1; j > 0; --j) reg[j] = (uint8_t)(reg[j - 1] ^ (fb & f->gen[j]));
    reg[0] = (uint8_t)(fb & f->gen[0]);
  }
  for (int i = 0; i < f->deg; ++i) word[k + i] = reg[f->deg - 1 - i];
  free(reg);
}

static uint16_t gmul(const bch_field* f, uint16_t a, uint16_t b)
{
  return (a && b) ? f->exp[f->log[a] + f->log[b]] : 0;
}

/* Decode in place.  Returns the number of corrected bit errors (0 .. t) or -1 when the word is uncorrectable (left as is).
 * syn_out (optional): the 2t syndromes S_1 .. S_2t. */
int port_bch_decode(int code, uint8_t* word, int n_bch, uint16_t* syn_out)
{
  const int t = port_bch_t(code);
  const bch_field* f = field_for(code >= 6, t);
  uint16_t S[25];
  int any = 0;
  for (int j = 1; j <= 2 * t; ++j) {
    uint16_t s = 0;
    for (int i = 0; i < n_bch; ++i)
      if (word[i] & 1) s ^= f->exp[(int)(((long long)j * (n_bch - 1 - i)) % f->n)];
    S[j] = s; any |= s;
    if (syn_out) syn_out[j - 1] = s;
  }
  if (!any) return 0;
  /* Berlekamp-Massey: sigma(x) = 1 + sigma_1 x + ... */
  uint16_t C[26] = {1}, B[26] = {1}, T[26];
  int L = 0, mm = 1; uint16_t b = 1;
  for (int n = 0; n < 2 * t; ++n) {
    uint16_t d = S[n + 1];
    for (int i = 1; i <= L; ++i) d ^= gmul(f, C[i], S[n + 1 - i]);
    if (!d) { ++mm; continue; }
    const uint16_t coef = f->exp[(f->log[d] + f->n - f->log[b]) % f->n];
    memcpy(T, C, sizeof(C));
    for (int i = 0; i + mm <= 2 * t; ++i) C[i + mm] ^= gmul(f, coef, B[i]);
    if (2 * L <= n) { L = n + 1 - L; memcpy(B, T, sizeof(C)); b = d; mm = 1; } else ++mm;
  }
  if (L > t) return -1;
  /* Chien search over the positions of the shortened code: error at power p <=> sigma(alpha^-p) = 0 */
  int found = 0, pos[16];
  for (int p = 0; p < n_bch && found <= L; ++p) {
    uint16_t v = 1;
    for (int i = 1; i <= L; ++i)
      if (C[i]) v ^= f->exp[(f->log[C[i]] + (int)(((long long)i * (f->n - p)) % f->n)) % f->n];
    if (!v) { if (found < 16) pos[found] = p; ++found; }
  }
  if (found != L) return -1;
  for (int e = 0; e < L; ++e) word[n_bch - 1 - pos[e]] ^= 1;
  return L;
}
