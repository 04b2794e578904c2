// TEST INFRASTRUCTURE ONLY -- thin C wrapper around the UNMODIFIED reference LDPC decoder.
// Compiled from the sources where they lie under /root/reference (never copied) by
// oracle/Makefile into oracle/_ref/libref_ldpc.so.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg may load it.
//
// What it drives (file:line relative to /root/reference/src/DVB_T2):
//   LDPC/layered_decoder.hh:50-190   LDPCDecoder<SIMD<int8_t,32>, ...>::init / operator()
//   LDPC/algorithms.hh:221-292       OffsetMinSumAlgorithm<SIMD<int8_t,W>, NormalUpdate, 2>
//   LDPC/ldpc.hh:39-123              LDPC<TABLE>
//   ldpc_decoder.cpp:248-277         the lane transposition + hard decision done by the glue
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include "LDPC/dvb_t2_tables.hh"
#include "LDPC/algorithms.hh"
#include "LDPC/layered_decoder.hh"

#define DEF(T) constexpr int T::DEG[]; constexpr int T::LEN[]; constexpr int T::POS[];
DEF(DVB_T2_TABLE_NORMAL_C1_2) DEF(DVB_T2_TABLE_NORMAL_C3_5) DEF(DVB_T2_TABLE_NORMAL_C2_3)
DEF(DVB_T2_TABLE_NORMAL_C3_4) DEF(DVB_T2_TABLE_NORMAL_C4_5) DEF(DVB_T2_TABLE_NORMAL_C5_6)
DEF(DVB_T2_TABLE_SHORT_C1_2) DEF(DVB_T2_TABLE_SHORT_C3_5) DEF(DVB_T2_TABLE_SHORT_C2_3)
DEF(DVB_T2_TABLE_SHORT_C3_4) DEF(DVB_T2_TABLE_SHORT_C4_5) DEF(DVB_T2_TABLE_SHORT_C5_6)
DEF(DVB_T2_TABLE_SHORT_C1_4) DEF(DVB_T2_TABLE_B8) DEF(DVB_T2_TABLE_B9)

namespace {
const int W = 32;                                    // SIZEOF_SIMD under __AVX2__ (ldpc_decoder.h:28-32)
typedef SIMD<int8_t, W> simd_type;
typedef NormalUpdate<simd_type> update_type;
typedef OffsetMinSumAlgorithm<simd_type, update_type, 2> algorithm_type;   // ldpc_decoder.h:34-62
typedef LDPCDecoder<simd_type, algorithm_type> decoder_type;

struct Code { int N, K; LDPCInterface* (*make)(); decoder_type* dec; };
template <class T> LDPCInterface* mk() { return new LDPC<T>(); }
// code ids follow include/t2b200.h: fec*6 + rate for the twelve PLP codes, then 12..14 for the L1 ones
Code codes[15] = {
  {64800, 32400, mk<DVB_T2_TABLE_NORMAL_C1_2>, nullptr}, {64800, 38880, mk<DVB_T2_TABLE_NORMAL_C3_5>, nullptr},
  {64800, 43200, mk<DVB_T2_TABLE_NORMAL_C2_3>, nullptr}, {64800, 48600, mk<DVB_T2_TABLE_NORMAL_C3_4>, nullptr},
  {64800, 51840, mk<DVB_T2_TABLE_NORMAL_C4_5>, nullptr}, {64800, 54000, mk<DVB_T2_TABLE_NORMAL_C5_6>, nullptr},
  {16200, 7200,  mk<DVB_T2_TABLE_SHORT_C1_2>, nullptr},  {16200, 9720,  mk<DVB_T2_TABLE_SHORT_C3_5>, nullptr},
  {16200, 10800, mk<DVB_T2_TABLE_SHORT_C2_3>, nullptr},  {16200, 11880, mk<DVB_T2_TABLE_SHORT_C3_4>, nullptr},
  {16200, 12600, mk<DVB_T2_TABLE_SHORT_C4_5>, nullptr},  {16200, 13320, mk<DVB_T2_TABLE_SHORT_C5_6>, nullptr},
  {16200, 3240,  mk<DVB_T2_TABLE_SHORT_C1_4>, nullptr},  {16200, 5400,  mk<DVB_T2_TABLE_B8>, nullptr},
  {16200, 6480,  mk<DVB_T2_TABLE_B9>, nullptr},
};
thread_local simd_type* tl_simd = nullptr;
}

extern "C" {

int ref_ldpc_code_n(int code) { return codes[code].N; }
int ref_ldpc_code_k(int code) { return codes[code].K; }

// Decode one batch of exactly 32 codewords the way ldpc_decoder::execute does.
//   llr       int8[32][N]   codeword order (positive => bit 0)
//   bits_out  uint8[32][K]  one byte per bit (may be null)
//   post_out  int8[32][N]   posteriors after the run, codeword order (may be null)
// returns the reference's `count` (trials left, <0 => "could not recover": the reference
// then emits nothing for the whole batch; bits_out is still filled here for inspection).
// NOT thread-safe per code (the decoder object holds the message memory): one caller per code id,
// use ref_ldpc_clone_decode32 from worker threads.
static int decode32(decoder_type* dec, int code, const int8_t* llr, uint8_t* bits_out, int8_t* post_out, int trials)
{
  const int N = codes[code].N, K = codes[code].K, q = (N - K) / 360;
  if (!tl_simd) tl_simd = new (std::align_val_t(sizeof(simd_type))) simd_type[64800];
  simd_type* simd = tl_simd;
  for (int k = 0; k < W; ++k) {                      // ldpc_decoder.cpp:248-260
    const int8_t* in = llr + (size_t)k * N;
    for (int i = 0; i < K; ++i) reinterpret_cast<int8_t*>(simd + i)[k] = in[i];
    for (int t = 0; t < q; ++t)
      for (int s = 0; s < 360; ++s)
        reinterpret_cast<int8_t*>(simd + K + q * s + t)[k] = in[K + 360 * t + s];
  }
  int count = (*dec)(simd, simd + K, trials, W);     // ldpc_decoder.cpp:262-263
  if (bits_out)
    for (int j = 0; j < W; ++j)                      // ldpc_decoder.cpp:270-277
      for (int i = 0; i < K; ++i)
        bits_out[(size_t)j * K + i] = reinterpret_cast<int8_t*>(simd + i)[j] < 0 ? 1 : 0;
  if (post_out)
    for (int k = 0; k < W; ++k) {
      int8_t* o = post_out + (size_t)k * N;
      for (int i = 0; i < K; ++i) o[i] = reinterpret_cast<int8_t*>(simd + i)[k];
      for (int t = 0; t < q; ++t)
        for (int s = 0; s < 360; ++s)
          o[K + 360 * t + s] = reinterpret_cast<int8_t*>(simd + K + q * s + t)[k];
    }
  return count;
}

void* ref_ldpc_new(int code)
{
  decoder_type* d = new decoder_type();
  LDPCInterface* it = codes[code].make();
  d->init(it);
  delete it;
  return d;
}

int ref_ldpc_decode32(void* dec, int code, const int8_t* llr, uint8_t* bits_out, int8_t* post_out, int trials)
{
  return decode32(static_cast<decoder_type*>(dec), code, llr, bits_out, post_out, trials);
}

}
