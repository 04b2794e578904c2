// K5 + K6: layered offset-min-sum LDPC decoder for every DVB-T2 code, one CTA per codeword,
// whole decoder state resident in shared memory, with the BCH-parity strip + BB descramble fused
// into the epilogue.
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
//   * posteriors int8[N] (<= 64.8 KB) + one packed word per check node (two clamped minima, arg-min
//     slot, output signs: the min-sum messages of a check node are fully determined by those) stay
//     in shared memory for the whole decode -> HBM traffic is the compulsory N bytes in, K out.
//   * the quasi-cyclic structure makes every edge of a layer a contiguous (rotated) run of 360
//     posteriors: thread j of the CTA owns check node (i, j), so all shared-memory traffic is
//     conflict-free byte-contiguous across a warp; no position table is read, addresses come from
//     q * CNL (base, shift) pairs.
//   * the reference's serial j order matters only where two check nodes of one layer share a bit;
//     the host precomputes the dependency depth of every check node in such layers and the CTA runs
//     them level by level (ldpc_schedule.cpp), everything else is one parallel step per layer.
//   * lock-step groups of 32 (reference batch semantics) are 32 co-resident CTAs that exchange their
//     parity verdict through one global word per iteration.
#include "ctx.h"
#include "ldpc_schedule.h"
#include <cstring>
#include <algorithm>

namespace {

constexpr int kThreads = 384;          // 360 check nodes of a layer + 24 idle lanes (12 warps)
constexpr int kSyncStride = 64;        // group-sync words per group (max_trials + 1 <= 64)

struct LdpcParams {
  const int8_t* llr; uint8_t* bits; int32_t* trials_left; int32_t* iters; int8_t* post_out;
  int n_cw, group_lanes, max_trials; unsigned flags;
  int N, K, q, k_out;
  const uint32_t* edge; const uint8_t* cnt; const int16_t* cidx; const uint8_t* nlev; const uint8_t* level;
  const uint8_t* prbs; unsigned* gsync;
};

template <typename ST> struct StateCodec;
template <> struct StateCodec<uint32_t> {       // <= 15 slots: signs[0,15) idx[15,20) m0[20,26) m1[26,32)
  static __device__ __forceinline__ void unpack(uint32_t w, uint32_t& sg, int& idx, int& m0, int& m1) {
    sg = w & 0x7fffu; idx = (w >> 15) & 31; m0 = (w >> 20) & 63; m1 = w >> 26;
  }
  static __device__ __forceinline__ uint32_t pack(uint32_t sg, int idx, int m0, int m1) {
    return sg | ((uint32_t)idx << 15) | ((uint32_t)m0 << 20) | ((uint32_t)m1 << 26);
  }
};
template <> struct StateCodec<uint64_t> {       // <= 32 slots: signs in the low word
  static __device__ __forceinline__ void unpack(uint64_t w, uint32_t& sg, int& idx, int& m0, int& m1) {
    sg = (uint32_t)w; uint32_t h = (uint32_t)(w >> 32); idx = h & 31; m0 = (h >> 8) & 63; m1 = (h >> 16) & 63;
  }
  static __device__ __forceinline__ uint64_t pack(uint32_t sg, int idx, int m0, int m1) {
    return (uint64_t)sg | ((uint64_t)((uint32_t)idx | ((uint32_t)m0 << 8) | ((uint32_t)m1 << 16)) << 32);
  }
};

__device__ __forceinline__ int rot360(int j, int sh) { int t = j + sh; return t >= 360 ? t - 360 : t; }

// One check node, both passes (LDPC/layered_decoder.hh:87-107 + algorithms.hh:250-291).
template <int CNL, typename ST>
__device__ __forceinline__ void check_node(int8_t* __restrict__ post, ST* __restrict__ state,
                                           const uint32_t* __restrict__ edge_i, int cnt, int i, int j,
                                           int K, int q)
{
  constexpr int SLOTS = CNL + 2;
  ST w = state[i * 360 + j];
  uint32_t sg; int idx, m0c, m1c;
  StateCodec<ST>::unpack(w, sg, idx, m0c, m1c);
  int inp[SLOTS], adr[SLOTS];
  int key0 = 1 << 20, key1 = 1 << 20, sx = 0;
  const bool hasB = (i | j) != 0;

  auto edge_in = [&](int slot, int a) {
    int pv = post[a];
    int m = (slot == idx) ? m1c : m0c;                       // stored message magnitude (<= 32)
    int bl = ((sg >> slot) & 1u) ? -m : min(m, 31);          // clamp(out, -32, 31)
    int v = max(min(pv - bl, 127), -128);                    // vqsub
    inp[slot] = v; adr[slot] = a;
    int mag = max(min(abs(v), 127) - 1, 0);                  // vqabs, then unsigned vqsub beta
    int key = (mag << 5) | slot;
    key1 = min(key1, max(key0, key));
    key0 = min(key0, key);
    sx ^= v;
  };
#pragma unroll
  for (int c = 0; c < CNL; ++c)
    if (c < cnt) {
      uint32_t e = __ldg(edge_i + c);
      edge_in(c, (int)(e & 0xffffu) + rot360(j, (int)(e >> 16)));
    }
  edge_in(CNL, K + 360 * i + j);
  if (hasB) edge_in(CNL + 1, i ? K + 360 * (i - 1) + j : K + 360 * (q - 1) + j - 1);

  const int m0 = key0 >> 5, idn = key0 & 31, m1 = key1 >> 5;
  uint32_t nsg = 0;
  auto edge_out = [&](int slot) {
    int v = inp[slot];
    int mg = (slot == idn) ? m1 : m0;                        // other(mags[i], mins[0], mins[1])
    bool neg = ((sx ^ v) < 0);                               // sign of the product of the other links
    int o = neg ? -mg : mg;
    post[adr[slot]] = (int8_t)max(min(v + o, 127), -128);    // vqadd
    nsg |= (uint32_t)neg << slot;
  };
#pragma unroll
  for (int c = 0; c < CNL; ++c)
    if (c < cnt) edge_out(c);
  edge_out(CNL);
  if (hasB) edge_out(CNL + 1);
  state[i * 360 + j] = StateCodec<ST>::pack(nsg, idn, min(m0, 32), min(m1, 32));
}

// bad() for the check nodes of row j in every layer (LDPC/layered_decoder.hh:65-82)
template <int CNL>
__device__ __forceinline__ int syndrome_rows(const int8_t* __restrict__ post, const LdpcParams& p, int j)
{
  int bad = 0;
  for (int i = 0; i < p.q; ++i) {
    const int cnt = __ldg(p.cnt + i);
    const uint32_t* edge_i = p.edge + i * CNL;
    int x = 0, nz = 1;
#pragma unroll
    for (int c = 0; c < CNL; ++c)
      if (c < cnt) {
        uint32_t e = __ldg(edge_i + c);
        int pv = post[(int)(e & 0xffffu) + rot360(j, (int)(e >> 16))];
        x ^= pv; nz &= (pv != 0);
      }
    int pa = post[p.K + 360 * i + j];
    x ^= pa; nz &= (pa != 0);
    if (i | j) {
      int pb = post[i ? p.K + 360 * (i - 1) + j : p.K + 360 * (p.q - 1) + j - 1];
      x ^= pb; nz &= (pb != 0);
    }
    bad |= (x < 0) | !nz;
  }
  return bad;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int CNL, typename ST>
__global__ void __launch_bounds__(kThreads, 1) ldpc_decode_kernel(const LdpcParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  int8_t* post = reinterpret_cast<int8_t*>(smem);
  ST* state = reinterpret_cast<ST*>(smem + ((p.N + 15) & ~15));
  __shared__ int s_flag;

  const int tid = threadIdx.x;
  const int GL = p.group_lanes;
  const int lane = blockIdx.x % GL, slot = blockIdx.x / GL, nslots = gridDim.x / GL;
  const int n_groups = (p.n_cw + GL - 1) / GL;
  const int R = p.N - p.K;

  for (int g = slot; g < n_groups; g += nslots) {
    const int lanes_here = min(GL, p.n_cw - g * GL);
    if (lane >= lanes_here) continue;
    const int cw = g * GL + lane;

    // ---- load the codeword's channel LLRs, clear the check-node state (reset()) ----
    {
      const int2* src = reinterpret_cast<const int2*>(p.llr + (size_t)cw * p.N);
      int2* dst = reinterpret_cast<int2*>(post);
      for (int k = tid; k < p.N / 8; k += kThreads) dst[k] = __ldg(src + k);
      for (int k = tid; k < R; k += kThreads) state[k] = 0;
    }
    __syncthreads();

    int trials = p.max_trials, iters = 0, lane_bad = 0;
    for (;;) {
      int bad = (tid < 360) ? syndrome_rows<CNL>(post, p, tid) : 0;
      lane_bad = __syncthreads_or(bad);
      int group_bad = lane_bad;
      if (GL > 1) {
        if (tid == 0) {
          unsigned* w = p.gsync + (size_t)g * kSyncStride + iters;
          atomicAdd(w, 1u | (lane_bad ? 0x10000u : 0u));
          unsigned v;
          while (((v = ld_acquire(w)) & 0xffffu) != (unsigned)lanes_here) __nanosleep(64);
          s_flag = (v >> 16) != 0;
        }
        __syncthreads();
        group_bad = s_flag;
      }
      if (!(group_bad && --trials >= 0)) break;
      // ---- one update() ----
      for (int i = 0; i < p.q; ++i) {
        const int cnt = __ldg(p.cnt + i);
        const int nl = __ldg(p.nlev + i);
        const uint32_t* edge_i = p.edge + i * CNL;
        if (nl == 1) {
          if (tid < 360) check_node<CNL, ST>(post, state, edge_i, cnt, i, tid, p.K, p.q);
          __syncthreads();
        } else {
          const int mylev = (tid < 360) ? __ldg(p.level + (int)__ldg(p.cidx + i) * 360 + tid) : 0;
          for (int l = 1; l <= nl; ++l) {
            if (mylev == l) check_node<CNL, ST>(post, state, edge_i, cnt, i, tid, p.K, p.q);
            __syncthreads();
          }
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

template <int CNL, typename ST>
cudaError_t launch(const LdpcParams& p, int grid, size_t smem, cudaStream_t st)
{
  auto k = ldpc_decode_kernel<CNL, ST>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (p.group_lanes > 1) {     // lock-step lanes spin on each other: they must be co-resident
    void* args[] = {(void*)&p};
    return cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(kThreads), args, smem, st);
  }
  k<<<grid, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int CNL, typename ST>
cudaError_t occupancy(size_t smem, int* blocks_per_sm)
{
  auto k = ldpc_decode_kernel<CNL, ST>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, kThreads, smem);
}

// smallest instantiated CNL >= cnl_max
const int kCnlBuckets[] = {4, 5, 7, 8, 9, 11, 12, 13, 16, 17, 20};

#define DISPATCH(CALL)                                                             \
  switch (cnl) {                                                                   \
    case 4:  return CALL(4, uint32_t);   case 5:  return CALL(5, uint32_t);        \
    case 7:  return CALL(7, uint32_t);   case 8:  return CALL(8, uint32_t);        \
    case 9:  return CALL(9, uint32_t);   case 11: return CALL(11, uint32_t);       \
    case 12: return CALL(12, uint32_t);  case 13: return CALL(13, uint32_t);       \
    case 16: return CALL(16, uint64_t);  case 17: return CALL(17, uint64_t);       \
    case 20: return CALL(20, uint64_t);                                            \
    default: return cudaErrorInvalidValue;                                         \
  }

cudaError_t launch_dispatch(int cnl, const LdpcParams& p, int grid, size_t smem, cudaStream_t st)
{
#define CALL_L(C, T) launch<C, T>(p, grid, smem, st)
  DISPATCH(CALL_L)
}
cudaError_t occupancy_dispatch(int cnl, size_t smem, int* bps)
{
#define CALL_O(C, T) occupancy<C, T>(smem, bps)
  DISPATCH(CALL_O)
}

}  // namespace

struct LdpcDeviceCode {
  LdpcSchedule s;
  int cnl = 0;              // instantiated bucket
  size_t state_bytes = 4, smem = 0;
  int blocks_per_sm = 0;
  uint32_t* d_edge = nullptr; uint8_t* d_cnt = nullptr; int16_t* d_cidx = nullptr;
  uint8_t* d_nlev = nullptr; uint8_t* d_level = nullptr;
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
  if (!d->cnl) { delete d; ctx->err = "LDPC code degree not instantiated"; return T2B200_ERR_ARG; }
  d->state_bytes = d->cnl + 2 <= 15 ? 4 : 8;
  d->smem = (size_t)((s.N + 15) & ~15) + (size_t)s.R * d->state_bytes;
  // edge table re-strided to the instantiated CNL
  std::vector<uint32_t> edge((size_t)s.q * d->cnl, 0);
  for (int i = 0; i < s.q; ++i)
    for (int c = 0; c < s.cnl_max; ++c) edge[(size_t)i * d->cnl + c] = s.edge[(size_t)i * s.cnl_max + c];
  std::vector<uint8_t> level = s.level; if (level.empty()) level.resize(360, 1);
  T2_CUDA(ctx, cudaMalloc(&d->d_edge, edge.size() * 4));
  T2_CUDA(ctx, cudaMalloc(&d->d_cnt, s.q));
  T2_CUDA(ctx, cudaMalloc(&d->d_cidx, s.q * 2));
  T2_CUDA(ctx, cudaMalloc(&d->d_nlev, s.q));
  T2_CUDA(ctx, cudaMalloc(&d->d_level, level.size()));
  T2_CUDA(ctx, cudaMemcpyAsync(d->d_edge, edge.data(), edge.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  T2_CUDA(ctx, cudaMemcpyAsync(d->d_cnt, s.cnt.data(), s.q, cudaMemcpyHostToDevice, ctx->stream));
  T2_CUDA(ctx, cudaMemcpyAsync(d->d_cidx, s.conflict_index.data(), s.q * 2, cudaMemcpyHostToDevice, ctx->stream));
  T2_CUDA(ctx, cudaMemcpyAsync(d->d_nlev, s.nlev.data(), s.q, cudaMemcpyHostToDevice, ctx->stream));
  T2_CUDA(ctx, cudaMemcpyAsync(d->d_level, level.data(), level.size(), cudaMemcpyHostToDevice, ctx->stream));
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
  T2_CUDA(ctx, occupancy_dispatch(d->cnl, d->smem, &d->blocks_per_sm));
  if (d->blocks_per_sm < 1) { ctx->err = "LDPC kernel does not fit on an SM"; return T2B200_ERR_CUDA; }
  ctx->ldpc[code] = d;
  *out = d;
  return T2B200_OK;
}

void t2_ldpc_free(t2b200_ctx* ctx)
{
  for (auto& kv : ctx->ldpc) {
    LdpcDeviceCode* d = kv.second;
    cudaFree(d->d_edge); cudaFree(d->d_cnt); cudaFree(d->d_cidx); cudaFree(d->d_nlev); cudaFree(d->d_level);
    delete d;
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

  LdpcParams p{};
  const void* d_llr; void *d_bits = nullptr, *d_tr = nullptr, *d_it = nullptr, *d_post = nullptr;
  if ((rc = t2_to_device(ctx, 0, llr, (size_t)n_cw * s.N, &d_llr))) return rc;
  if (bits_out && (rc = t2_out_device(ctx, 1, bits_out, out_row * n_cw, &d_bits))) return rc;
  if (trials_left && (rc = t2_out_device(ctx, 2, trials_left, 4 * (size_t)n_cw, &d_tr))) return rc;
  if (iterations && (rc = t2_out_device(ctx, 3, iterations, 4 * (size_t)n_cw, &d_it))) return rc;
  if ((flags & T2B200_LDPC_WANT_POST) && (rc = t2_out_device(ctx, 4, post_out, (size_t)n_cw * s.N, &d_post))) return rc;

  p.llr = (const int8_t*)d_llr; p.bits = (uint8_t*)d_bits; p.trials_left = (int32_t*)d_tr; p.iters = (int32_t*)d_it;
  p.post_out = (int8_t*)d_post;
  p.n_cw = n_cw; p.max_trials = max_trials; p.flags = flags;
  p.group_lanes = (flags & T2B200_LDPC_GROUP32) ? 32 : 1;
  p.N = s.N; p.K = s.K; p.q = s.q; p.k_out = k_out;
  p.edge = d->d_edge; p.cnt = d->d_cnt; p.cidx = d->d_cidx; p.nlev = d->d_nlev; p.level = d->d_level;
  p.prbs = ctx->d_prbs;

  const int capacity = d->blocks_per_sm * ctx->sm_count;
  int grid;
  if (p.group_lanes > 1) {
    const int n_groups = (n_cw + 31) / 32;
    int slots = std::min(capacity / 32, n_groups);
    if (slots < 1) { ctx->err = "GPU cannot co-schedule one 32-lane group"; return T2B200_ERR_CUDA; }
    grid = slots * 32;
    size_t need = (size_t)n_groups * kSyncStride * sizeof(unsigned);
    if (ctx->group_sync_cap < need) {
      if (ctx->d_group_sync) cudaFree(ctx->d_group_sync);
      T2_CUDA(ctx, cudaMalloc(&ctx->d_group_sync, need));
      ctx->group_sync_cap = need;
    }
    T2_CUDA(ctx, cudaMemsetAsync(ctx->d_group_sync, 0, need, ctx->stream));
    p.gsync = ctx->d_group_sync;
  } else {
    grid = std::min(capacity, n_cw);
  }
  T2_CUDA(ctx, launch_dispatch(d->cnl, p, grid, d->smem, ctx->stream));
  ctx->launches++;

  if (bits_out && (rc = t2_finish_out(ctx, bits_out, d_bits, out_row * n_cw))) return rc;
  if (trials_left && (rc = t2_finish_out(ctx, trials_left, d_tr, 4 * (size_t)n_cw))) return rc;
  if (iterations && (rc = t2_finish_out(ctx, iterations, d_it, 4 * (size_t)n_cw))) return rc;
  if (d_post && (rc = t2_finish_out(ctx, post_out, d_post, (size_t)n_cw * s.N))) return rc;
  return T2B200_OK;
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
