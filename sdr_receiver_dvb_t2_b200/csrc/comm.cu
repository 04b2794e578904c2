// Multi-GPU FEC: FEC blocks are independent, so the LDPC / BCH stage of a pooled batch is sharded by codeword
// across the GPUs of one box.  ONE exchange step each way -- the int8 LLRs of every rank's shard out from the demodulating
// rank, the BBFRAME bits back -- as NCCL point-to-point transfers over NVLink / NVSwitch, issued from here (not from
// Python): a side stream carries the transfers in chunks of <= 1152 codewords, double-buffered against the decode on the
// context's stream, so a rank decodes chunk t-1 while chunk t arrives and the bits of chunk t-2 leave.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": in a torch process that is the NCCL torch itself uses), so
// libt2b200.so has no link-time dependency on it and single-GPU users never touch it.
#include "stages.h"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

struct CommState {
  void* dl = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  bool wide = false;                  // more than two ranks: wide NCCL kernels, decoders on one group slot less (t2b200_comm_init)
  cudaStream_t s_comm = nullptr;
  cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_in[2] = {}, ev_dec[2] = {};
  int8_t* in_buf = nullptr; uint8_t* out_buf = nullptr; size_t in_cap = 0, out_cap = 0;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;    // optional (NCCL >= 2.14)
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

namespace {

constexpr int kNcclMaxCtasDefault = 2;      // see t2b200_comm_init

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e && *e ? atoi(e) : dflt; }

bool load_nccl(CommState* s, std::string& err)
{
  if (s->dl) return true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    s->dl = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (s->dl) break;
  }
  if (!s->dl) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define T2_SYM(field, sym)                                                       \
  s->field = reinterpret_cast<decltype(s->field)>(dlsym(s->dl, sym));           \
  if (!s->field) { err = std::string("NCCL symbol missing: ") + sym; return false; }
  T2_SYM(GetUniqueId, "ncclGetUniqueId") T2_SYM(CommInitRank, "ncclCommInitRank") T2_SYM(CommDestroy, "ncclCommDestroy")
  T2_SYM(Send, "ncclSend") T2_SYM(Recv, "ncclRecv") T2_SYM(GroupStart, "ncclGroupStart") T2_SYM(GroupEnd, "ncclGroupEnd")
  T2_SYM(GetErrorString, "ncclGetErrorString")
#undef T2_SYM
  s->CommInitRankConfig = reinterpret_cast<decltype(s->CommInitRankConfig)>(dlsym(s->dl, "ncclCommInitRankConfig"));
  return true;
}

// contiguous shard of rank r: whole 32-codeword lock-step groups (the reference's batch, ldpc_decoder.h:28-32), in proportion
// to the group slots each rank decodes on (weight[r])
void span_of(int n_cw, int r, const std::vector<int>& weight, int* lo, int* hi)
{
  const long long units = (n_cw + 31) / 32;
  long long wsum = 0, wlo = 0;
  for (size_t i = 0; i < weight.size(); ++i) { wsum += weight[i]; if ((int)i < r) wlo += weight[i]; }
  const long long lo_u = units * wlo / wsum, hi_u = units * (wlo + weight[r]) / wsum;
  *lo = (int)std::min<long long>(lo_u * 32, n_cw); *hi = (int)std::min<long long>(hi_u * 32, n_cw);
}

// more than two ranks (T2B200_SHARD_WIDE=0 / 1 forces it: development aid)
bool wide_mode(int nranks) { const int e = env_int("T2B200_SHARD_WIDE", -1); return e >= 0 ? e != 0 : nranks > 2; }

CommState g_loader;                   // t2b200_comm_unique_id needs NCCL before any context has a communicator

}  // namespace

#define T2_NCCL(ctx, s, call)                                                                   \
  do {                                                                                          \
    ncclResult_t r__ = (call);                                                                  \
    if (r__ != ncclSuccess) {                                                                   \
      (ctx)->err = std::string(#call) + ": " + ((s)->GetErrorString ? (s)->GetErrorString(r__) : "NCCL error"); \
      return T2B200_ERR_CUDA;                                                                   \
    }                                                                                           \
  } while (0)

void t2_comm_free(t2b200_ctx* ctx)
{
  CommState* s = ctx->comm;
  if (!s) return;
  if (s->comm && s->CommDestroy) s->CommDestroy(s->comm);
  if (s->s_comm) cudaStreamDestroy(s->s_comm);
  for (cudaEvent_t e : {s->ev_start, s->ev_done, s->ev_in[0], s->ev_in[1], s->ev_dec[0], s->ev_dec[1]}) if (e) cudaEventDestroy(e);
  cudaFree(s->in_buf); cudaFree(s->out_buf);
  delete s;
  ctx->comm = nullptr;
}

extern "C" int t2b200_comm_unique_id(void* id_out, size_t id_bytes)
{
  std::string err;
  if (!id_out || id_bytes < sizeof(ncclUniqueId) || !load_nccl(&g_loader, err)) return T2B200_ERR_ARG;
  ncclUniqueId id;
  if (g_loader.GetUniqueId(&id) != ncclSuccess) return T2B200_ERR_CUDA;
  memcpy(id_out, &id, sizeof(id));
  return T2B200_OK;
}

extern "C" int t2b200_comm_init(t2b200_ctx* ctx, int rank, int nranks, const void* unique_id, size_t id_bytes)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (rank < 0 || nranks < 1 || rank >= nranks || !unique_id || id_bytes < sizeof(ncclUniqueId)) {
    ctx->err = "t2b200_comm_init: bad argument"; return T2B200_ERR_ARG;
  }
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  t2_comm_free(ctx);
  CommState* s = new CommState();
  ctx->comm = s;
  if (!load_nccl(s, ctx->err)) { t2_comm_free(ctx); return T2B200_ERR_STATE; }
  s->rank = rank; s->nranks = nranks;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  // NCCL's send / recv kernels share the SMs with the decoder.  Next to the decoder's full grid (9 group slots = 144 CTAs) only
  // TWO NCCL CTAs find room (measured: with three or more the cooperative decoder launch and the NCCL kernel wait for each
  // other and the exchange serialises with the decoding; profiles/r02_sharded_experiments.txt) -- ~70 GB/s, enough between
  // two ranks.  With more ranks the root has to feed everybody (1.8 GB per 8 x 4032 codewords while one shard decodes):
  // every rank then decodes on 8 of the 9 slots and NCCL gets 16 CTAs (~350 GB/s out of the root).  The setting must be the
  // same on all ranks of a communicator (NCCL fails otherwise).
  s->wide = wide_mode(nranks);
  if (s->CommInitRankConfig) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    const int max_ctas = env_int("T2B200_NCCL_MAX_CTAS", s->wide ? 16 : kNcclMaxCtasDefault);      // (development aid; 0: NCCL's default)
    if (max_ctas > 0) { cfg.minCTAs = 1; cfg.maxCTAs = max_ctas; }
    T2_NCCL(ctx, s, s->CommInitRankConfig(&s->comm, nranks, id, rank, &cfg));
  } else {
    T2_NCCL(ctx, s, s->CommInitRank(&s->comm, nranks, id, rank));
  }
  T2_CUDA(ctx, cudaStreamCreateWithFlags(&s->s_comm, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&s->ev_start, &s->ev_done, &s->ev_in[0], &s->ev_in[1], &s->ev_dec[0], &s->ev_dec[1]})
    T2_CUDA(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return T2B200_OK;
}

extern "C" int t2b200_comm_destroy(t2b200_ctx* ctx)
{
  if (!ctx) return T2B200_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  t2_comm_free(ctx);
  return T2B200_OK;
}

extern "C" int t2b200_ldpc_decode_sharded(t2b200_ctx* ctx, int code, int root, const int8_t* llr, int n_cw, uint8_t* bits_out,
                                          int max_trials, unsigned flags)
{
  if (!ctx) return T2B200_ERR_ARG;
  CommState* s = ctx->comm;
  if (!s || !s->comm) { ctx->err = "t2b200_ldpc_decode_sharded: call t2b200_comm_init first"; return T2B200_ERR_STATE; }
  const int N = t2b200_ldpc_n(code), K = t2b200_ldpc_k(code);
  if (!N || n_cw < 0 || root < 0 || root >= s->nranks || (flags & T2B200_LDPC_WANT_POST) || max_trials <= 0 || max_trials > 60) {
    ctx->err = "t2b200_ldpc_decode_sharded: bad argument"; return T2B200_ERR_ARG;
  }
  const bool is_root = s->rank == root;
  if (is_root && (!llr || !bits_out || !t2_is_device_ptr(llr) || !t2_is_device_ptr(bits_out))) {
    ctx->err = "t2b200_ldpc_decode_sharded: the root's llr / bits_out must be device memory"; return T2B200_ERR_ARG;
  }
  if (n_cw == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  int k_out = K;
  if (flags & T2B200_LDPC_BCH_DESCRAMBLE) {
    k_out = t2b200_ldpc_k_bch(code);
    if (!k_out) { ctx->err = "code has no BCH geometry"; return T2B200_ERR_ARG; }
  }
  const size_t row = (flags & T2B200_LDPC_PACK_BITS) ? (size_t)k_out / 8 : (size_t)k_out;
  const bool trace = getenv("T2B200_SHARD_TRACE") != nullptr;      // (development aid: device times of every exchange / decode step)
  std::vector<cudaEvent_t> tr;
  auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tr.push_back(e); } };
  // wide mode (see t2b200_comm_init): every decoder leaves 20 SMs to the NCCL kernels
  const int full_slots = std::max(1, ctx->sm_count / 16);
  const int slots = s->wide ? std::max(1, full_slots - 1) : full_slots;
  struct SlotsCap { t2b200_ctx* c; int old; ~SlotsCap() { c->ldpc_slots_cap = old; } } cap_guard{ctx, ctx->ldpc_slots_cap};
  ctx->ldpc_slots_cap = env_int("T2B200_SHARD_SLOTS", s->wide ? slots : 0);
  std::vector<int> weight(s->nranks, 1);
  // codewords per transfer: four full rounds of the group slots (a chunk that leaves a half-empty last round costs the remote
  // ranks 12 % of their decode time)
  const int kChunk = std::max(32, env_int("T2B200_SHARD_CHUNK", 4 * slots * 32) / 32 * 32);      // (development aid)
  // chunk t of a shard: a short first chunk (one round of the group slots) so that the remote decoders start almost at once,
  // then chunks that double up to kChunk (wide mode: 2 x kChunk; measured, profiles/r02_sharded_experiments.txt): every chunk ends with a tail in which group slots wait for the
  // chunk's slowest lock-step group, so the fewer and larger the later chunks the better -- as long as a chunk still arrives
  // while the one before it decodes
  const int kFirst = std::max(32, std::min(kChunk, env_int("T2B200_SHARD_FIRST", slots * 32) / 32 * 32));
  const int kMaxChunk = kChunk * std::max(1, env_int("T2B200_SHARD_GROW", s->wide ? 2 : 1));
  std::vector<int> lo(s->nranks), hi(s->nranks);
  std::vector<int> offs(1, 0);                                    // chunk boundaries inside a shard (the same for every rank)
  {
    int longest = 0;
    for (int r = 0; r < s->nranks; ++r) { int a, b; span_of(n_cw, r, weight, &a, &b); longest = std::max(longest, b - a); }
    for (int sz = kFirst; offs.back() < longest; sz = std::min(2 * sz, kMaxChunk)) offs.push_back(offs.back() + sz);
  }
  auto chunk_off = [&](int t) { return t <= 0 ? 0 : t < (int)offs.size() ? offs[t] : INT_MAX / 2; };
  auto chunk_n = [&](int r, int t) { return t < 0 ? 0 : std::max(0, std::min(chunk_off(t + 1), hi[r] - lo[r]) - chunk_off(t)); };
  int rounds = 0;
  for (int r = 0; r < s->nranks; ++r) {
    span_of(n_cw, r, weight, &lo[r], &hi[r]);
    if (r != root) while (chunk_n(r, rounds)) ++rounds;
  }
  int rc;
  // everything queued on the context's stream so far (the LLRs on the root, earlier users of the buffers) comes first
  T2_CUDA(ctx, cudaEventRecord(s->ev_start, ctx->stream));
  T2_CUDA(ctx, cudaStreamWaitEvent(s->s_comm, s->ev_start, 0));
  mark(ctx->stream);
  if (is_root) {
    // The root's own shard decodes on the context's stream while the side stream moves the other shards.  The decode is
    // queued FIRST: its 144 CTAs then hold their SMs and the few CTAs of NCCL's send / recv kernels (which sit waiting for
    // the peers' results for milliseconds) take the SMs left over, instead of the decoder waiting behind them.
    const int mine = hi[root] - lo[root];
    if (mine > 0 && (rc = t2_ldpc_device(ctx, code, llr + (size_t)lo[root] * N, mine, bits_out + (size_t)lo[root] * row, nullptr, nullptr,
                                         max_trials, flags))) return rc;
    for (int t = 0; t < rounds + 2; ++t) {
      bool any = false;
      for (int r = 0; r < s->nranks && !any; ++r) any = r != root && (chunk_n(r, t) || chunk_n(r, t - 2));
      if (!any) continue;
      T2_NCCL(ctx, s, s->GroupStart());
      for (int r = 0; r < s->nranks; ++r) {
        if (r == root) continue;
        if (const int n = chunk_n(r, t))
          T2_NCCL(ctx, s, s->Send(llr + (size_t)(lo[r] + chunk_off(t)) * N, (size_t)n * N, ncclInt8, r, s->comm, s->s_comm));
        if (const int n = chunk_n(r, t - 2))
          T2_NCCL(ctx, s, s->Recv(bits_out + (size_t)(lo[r] + chunk_off(t - 2)) * row, (size_t)n * row, ncclUint8, r, s->comm, s->s_comm));
      }
      T2_NCCL(ctx, s, s->GroupEnd());
      mark(s->s_comm);
    }
    mark(ctx->stream);
  } else {
    const int r = s->rank;
    if (s->in_cap < 2 * (size_t)kMaxChunk * N) {
      T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(s->in_buf); s->in_buf = nullptr; s->in_cap = 0;
      T2_CUDA(ctx, cudaMalloc(&s->in_buf, 2 * (size_t)kMaxChunk * N));
      s->in_cap = 2 * (size_t)kMaxChunk * N;
    }
    if (s->out_cap < 2 * (size_t)kMaxChunk * row) {
      T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(s->out_buf); s->out_buf = nullptr; s->out_cap = 0;
      T2_CUDA(ctx, cudaMalloc(&s->out_buf, 2 * (size_t)kMaxChunk * row));
      s->out_cap = 2 * (size_t)kMaxChunk * row;
    }
    for (int t = 0; t < rounds + 2; ++t) {
      const int n_in = chunk_n(r, t), n_out = chunk_n(r, t - 2);
      if (!n_in && !n_out) continue;
      int8_t* in = s->in_buf + (size_t)(t & 1) * kMaxChunk * N;
      uint8_t* out = s->out_buf + (size_t)(t & 1) * kMaxChunk * row;
      // round t re-uses the buffers of round t - 2: its decode must be over (it is what the send below carries anyway)
      if (t >= 2) T2_CUDA(ctx, cudaStreamWaitEvent(s->s_comm, s->ev_dec[t & 1], 0));
      T2_NCCL(ctx, s, s->GroupStart());
      if (n_in) T2_NCCL(ctx, s, s->Recv(in, (size_t)n_in * N, ncclInt8, root, s->comm, s->s_comm));
      if (n_out) T2_NCCL(ctx, s, s->Send(out, (size_t)n_out * row, ncclUint8, root, s->comm, s->s_comm));
      T2_NCCL(ctx, s, s->GroupEnd());
      mark(s->s_comm);
      if (n_in) {
        T2_CUDA(ctx, cudaEventRecord(s->ev_in[t & 1], s->s_comm));
        T2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s->ev_in[t & 1], 0));
        if ((rc = t2_ldpc_device(ctx, code, in, n_in, out, nullptr, nullptr, max_trials, flags))) return rc;
        T2_CUDA(ctx, cudaEventRecord(s->ev_dec[t & 1], ctx->stream));
        mark(ctx->stream);
      }
    }
  }
  // later work on the context's stream sees the gathered bits (root) / may re-use the buffers (others)
  T2_CUDA(ctx, cudaEventRecord(s->ev_done, s->s_comm));
  T2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s->ev_done, 0));
  if (trace) {
    cudaStreamSynchronize(ctx->stream);
    std::string line = "shard trace rank " + std::to_string(s->rank) + ":";
    for (size_t i = 1; i < tr.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, tr[0], tr[i]); char b[32]; snprintf(b, sizeof b, " %.2f", ms); line += b; }
    fprintf(stderr, "%s\n", line.c_str());
    for (auto e : tr) cudaEventDestroy(e);
  }
  return T2B200_OK;
}
