// N1: BBFRAME -> MPEG transport stream re-packetiser on the GPU: high-efficiency mode by a parallel scan (this
// file), normal mode and mixed batches by the general path (ts_general.h, ts_general_kernel below).
//
// Reference semantics reproduced (bb_de_header.cpp, paths relative to the reference's src/DVB_T2):
//   :70-82,101-113  CRC-8 of the 80 header bits: residue 0xAB => HEM, 0 => normal mode, else the frame is dropped
//   :136-163        DFL, SYNCD (bits); SYNCD == 65535 => frame dropped
//   :332-428        HEM: 187-byte user packets on air, 0x47 re-inserted in front of each; bytes are emitted while at
//                   least 188 bytes of data field remain, the rest (< 188 bytes, with the sync byte if a packet
//                   boundary falls inside) is held back and opens the NEXT frame's datagram, completed by SYNCD / 8
//                   bytes of that frame (equal / longer / shorter SYNCD: :341-382, 0xF0 fill in the last case)
//   :431-441        one BBFRAME = one datagram
// The reference walks the frame bit by bit; here the only serial part is the packet phase carried from frame to
// frame: a one-thread scan over 12-byte header records turns every frame into a descriptor (where its datagram
// starts, which bit ranges feed it, where the sync bytes fall), and one CTA per frame then builds the datagram
// with one thread per output byte.  A batch with a normal-mode frame (status 3), or entered with normal mode's run-away
// packet index, is switched ON THE DEVICE to the general path: ts_parse_kernel raises TsDevState::general, the three scan
// kernels return at once and ts_general_kernel (one CTA, frame after frame: plan by one thread, segments by the warps,
// CRC tasks by the threads) builds the datagrams exactly as the reference's byte-serial loop would.
#include "ctx.h"
#include "ts_general.h"
#include <algorithm>

namespace {

constexpr int PKT = 188;

struct TsHdr {                                         // status: 0 HEM ok, 1 header CRC, 2 SYNCD == 65535, 3 normal mode, 4 DFL too long
  int status, dfl, syncd;
  // what a frame does when it is entered with a held-back tail (the usual case: its packet phase after the head is 0, so
  // everything behind the head depends on the frame alone) -- computed by the parallel parse kernel, divisions included
  int main_n, main_out;                                // data bytes / output bytes (with sync bytes) of the main run
  int ntail, tail_sync, end_packet, end_buffer;        // held-back bytes, sync position inside them, state afterwards
};
struct TsDesc {
  long long out_off; int out_len;
  int carry_len, carry_src, carry_bit, carry_sync;     // carry_src: frame index, -1 = state buffer of the previous call
  int head_n, head_f0;
  int main_bit, main_n, main_phase;
};
struct TsDevState {                                    // survives between calls (device memory)
  int split, idx_packet, idx_buffer, general;          // general: this call runs on the general path (set by the parse kernel)
  unsigned crc; int pad[3];                            // normal mode: CRC-8 of the packet piece in flight
  long long total;                                     // bytes emitted by the last call
  int tail_src, tail_bit, tail_ndata, tail_sync;       // where the held-back bytes of the last call live
  uint8_t buffer[PKT + 4];
};

__device__ __forceinline__ unsigned field(const uint8_t* b, int n)
{
  unsigned v = 0;
  for (int i = 0; i < n; ++i) v = (v << 1) | (b[i] & 1u);
  return v;
}

// main run + held-back tail of a frame whose data field (behind SYNCD) has `dfl` bits left and whose packet index is
// idx_packet when the main run starts (bb_de_header.cpp:384-428)
struct TsBody { int M, T, phase, ntail, tail_sync, end_packet, end_buffer, split; };
__device__ __forceinline__ TsBody ts_body(int dfl, int idx_packet)
{
  TsBody b = {0, 0, 0, 0, -1, idx_packet, 0, 0};
  if (dfl >= PKT * 8) {
    const int M = (dfl - PKT * 8) / 8 + 1;
    const int phase = idx_packet == PKT ? 0 : idx_packet;
    int T;
    if (phase == 0) T = M + (M + PKT - 2) / (PKT - 1);
    else if (M <= PKT - phase) T = M;
    else T = M + (M - (PKT - phase) + PKT - 2) / (PKT - 1);
    b.M = M; b.T = T; b.phase = phase;
    dfl -= 8 * M;
    idx_packet = (phase + T) % PKT;
    if (idx_packet == 0) idx_packet = PKT;
  }
  if (dfl > 0) {                                         // held back for the next frame (bb_de_header.cpp:386-402)
    b.split = 1;
    const int ntail = dfl / 8;
    b.ntail = ntail;
    if (idx_packet == PKT) { if (ntail > 0) b.tail_sync = 0; }
    else if (idx_packet != 0 && PKT - idx_packet < ntail) b.tail_sync = PKT - idx_packet;
    if (b.tail_sync >= 0) idx_packet = 1 + (ntail - b.tail_sync); else idx_packet += ntail;
    b.end_buffer = ntail + (b.tail_sync >= 0 ? 1 : 0);
  }
  b.end_packet = idx_packet;
  return b;
}

__global__ void ts_parse_kernel(const uint8_t* __restrict__ frames, int n_frames, int k_bch, TsHdr* __restrict__ hdr,
                                TsDevState* __restrict__ st)
{
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  if (f == 0 && st->idx_packet > PKT) st->general = 1;   // left behind by normal mode's resynchronisation (bb_de_header.cpp:208-226)
  const uint8_t* b = frames + (size_t)f * k_bch;
  unsigned reg = 0;                                      // bb_de_header.cpp:70-82
  for (int i = 0; i < 80; ++i) {
    const unsigned bit = (b[i] ^ reg) & 1u;
    reg >>= 1;
    if (bit) reg ^= 0xABu;
  }
  TsHdr h;
  h.dfl = (int)field(b + 32, 16);
  h.syncd = (int)field(b + 56, 16);
  h.status = reg == 0xABu ? 0 : reg == 0u ? 3 : 1;
  if (h.status != 1 && h.syncd == 65535) h.status = 2;
  // A header whose CRC-8 passes by chance (or a crafted one) can announce a data field longer than the frame: the reference
  // would walk off the end of its buffer (it has no such check); here the frame is dropped like a header CRC error, without
  // touching the carried packet state.
  if ((h.status == 0 || h.status == 3) && 80 + h.dfl > k_bch) h.status = 4;
  if (h.status == 3) st->general = 1;                    // (same value from every thread that writes it)
  const TsBody body = ts_body(h.dfl - h.syncd, PKT);     // entered with a tail: the head completes a packet first
  h.main_n = body.M; h.main_out = body.T; h.ntail = body.split ? body.ntail : -1; h.tail_sync = body.tail_sync;
  h.end_packet = body.end_packet; h.end_buffer = body.end_buffer;
  hdr[f] = h;
}

// One warp: the lanes stage 32 header records in shared memory and write 32 descriptors back; lane 0 walks them.
__global__ void __launch_bounds__(32) ts_scan_kernel(const TsHdr* __restrict__ hdr, int n_frames, TsDesc* __restrict__ desc,
                                                     TsDevState* __restrict__ st, int32_t* __restrict__ dlen, int32_t* __restrict__ status)
{
  __shared__ TsHdr sh[32];
  __shared__ TsDesc sd[32];
  if (st->general) return;
  const int lane = threadIdx.x;
  int split = st->split, idx_packet = st->idx_packet, idx_buffer = st->idx_buffer;
  int tail_src = -1, tail_bit = 0, tail_ndata = 0, tail_sync = -1;   // the carried bytes of a previous call sit in st->buffer
  long long off = 0;
  for (int f0 = 0; f0 < n_frames; f0 += 32) {
    const int nf = min(32, n_frames - f0);
    if (lane < nf) sh[lane] = hdr[f0 + lane];
    __syncwarp();
    if (lane == 0) {
      for (int i = 0; i < nf; ++i) {
        const int f = f0 + i;
        const TsHdr h = sh[i];
        TsDesc d;
        d.out_off = off; d.out_len = 0; d.carry_len = 0; d.carry_src = -1; d.carry_bit = 0; d.carry_sync = -1;
        d.head_n = 0; d.head_f0 = 0; d.main_bit = 0; d.main_n = 0; d.main_phase = 0;
        if (h.status == 0) {
          int in_bit = 80;
          TsBody b;
          if (split) {
            d.carry_src = tail_src; d.carry_bit = tail_bit; d.carry_sync = tail_sync;
            d.carry_len = idx_buffer;
            const int missing = PKT - idx_packet, sb = h.syncd >> 3;
            if (missing <= sb) {
              d.head_n = missing;
              in_bit += missing == sb ? missing * 8 : h.syncd;
            } else {
              d.head_n = sb; d.head_f0 = missing - sb;
              in_bit += sb * 8;
            }
            // behind the head the packet index is 188: the parse kernel has done the rest
            b.M = h.main_n; b.T = h.main_out; b.phase = 0; b.split = h.ntail >= 0; b.ntail = max(h.ntail, 0);
            b.tail_sync = h.tail_sync; b.end_packet = h.end_packet; b.end_buffer = h.end_buffer;
          } else {
            in_bit += h.syncd;
            b = ts_body(h.dfl - h.syncd, idx_packet);        // rare: first frame, or the previous one left no tail
          }
          d.main_bit = in_bit; d.main_n = b.M; d.main_phase = b.phase;
          in_bit += 8 * b.M;
          idx_packet = b.end_packet;
          split = b.split;
          if (b.split) {
            tail_src = f; tail_bit = in_bit; tail_ndata = b.ntail; tail_sync = b.tail_sync;
            idx_buffer = b.end_buffer;
          }
          d.out_len = d.carry_len + d.head_n + d.head_f0 + b.T;
        }
        sd[i] = d;
        off += d.out_len;
      }
    }
    __syncwarp();
    if (lane < nf) {
      desc[f0 + lane] = sd[lane];
      if (dlen) dlen[f0 + lane] = sd[lane].out_len;
      if (status) status[f0 + lane] = sh[lane].status;
    }
    __syncwarp();
  }
  if (lane == 0) {
    st->split = split; st->idx_packet = idx_packet; st->idx_buffer = idx_buffer; st->total = off;
    st->tail_src = tail_src; st->tail_bit = tail_bit; st->tail_ndata = tail_ndata; st->tail_sync = tail_sync;
  }
}

__device__ __forceinline__ uint8_t gather8(const uint8_t* __restrict__ bits)
{
  unsigned v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) v = (v << 1) | (bits[i] & 1u);
  return (uint8_t)v;
}

// carried byte j of a tail that lives in `frame` (sync byte inserted at index `sync`)
__device__ __forceinline__ uint8_t tail_byte(const uint8_t* __restrict__ frame, int bit, int sync, int j)
{
  if (j == sync) return 0x47;
  const int idx = j - (sync >= 0 && j > sync ? 1 : 0);
  return gather8(frame + bit + 8 * idx);
}

__global__ void __launch_bounds__(256) ts_assemble_kernel(const uint8_t* __restrict__ frames, int k_bch,
                                                          const TsDesc* __restrict__ desc, const TsDevState* __restrict__ st,
                                                          uint8_t* __restrict__ out, long long cap)
{
  if (st->general) return;
  const uint8_t* carry0 = st->buffer;
  const TsDesc d = desc[blockIdx.x];
  const uint8_t* me = frames + (size_t)blockIdx.x * k_bch;
  const int c1 = d.carry_len, c2 = c1 + d.head_n, c3 = c2 + d.head_f0;
  for (int j = threadIdx.x; j < d.out_len; j += blockDim.x) {
    uint8_t v;
    if (j < c1) v = d.carry_src >= 0 ? tail_byte(frames + (size_t)d.carry_src * k_bch, d.carry_bit, d.carry_sync, j) : carry0[j];
    else if (j < c2) v = gather8(me + 80 + 8 * (j - c1));
    else if (j < c3) v = 0xF0;
    else {
      const int jj = j - c3, t = d.main_phase + jj;
      if (t % PKT == 0) v = 0x47;
      else {
        const int nsync = t / PKT + (d.main_phase == 0 ? 1 : 0);
        v = gather8(me + d.main_bit + 8 * (jj - nsync));
      }
    }
    if (d.out_off + j < cap) out[d.out_off + j] = v;
  }
}

// keep the held-back bytes of the call's last frame for the next call
__global__ void ts_save_tail_kernel(const uint8_t* __restrict__ frames, int k_bch, TsDevState* __restrict__ st)
{
  if (st->general || !st->split || st->tail_src < 0) return;   // nothing new held back (an older tail stays where it is)
  const int n = st->tail_ndata + (st->tail_sync >= 0 ? 1 : 0);
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    st->buffer[j] = tail_byte(frames + (size_t)st->tail_src * k_bch, st->tail_bit, st->tail_sync, j);
}

// ---- the general path (ts_general.h): one CTA walks the batch frame by frame ----
__device__ __forceinline__ unsigned byte_at(const uint8_t* __restrict__ frame, int k_bch, int bit)
{
  unsigned v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) v = (v << 1) | ((bit + i < k_bch ? frame[bit + i] : 0) & 1u);      // past the frame: zero
  return v;
}

__global__ void __launch_bounds__(256) ts_general_kernel(const uint8_t* __restrict__ frames, int n_frames, int k_bch,
                                                         const TsHdr* __restrict__ hdr, TsDevState* __restrict__ st,
                                                         uint8_t* __restrict__ out, long long cap, int32_t* __restrict__ dlen,
                                                         int32_t* __restrict__ status)
{
  if (!st->general) return;
  __shared__ TsgPlan P;
  __shared__ TsgState S;
  __shared__ uint8_t old_buf[PKT + 4], new_buf[PKT + 4];
  __shared__ unsigned crc_in, crc_out;
  __shared__ long long off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { S.split = st->split; S.idx_packet = st->idx_packet; S.idx_buffer = st->idx_buffer; S.crc = st->crc; off = 0; }
  for (int j = tid; j < PKT + 4; j += blockDim.x) new_buf[j] = st->buffer[j];
  __syncthreads();
  for (int f = 0; f < n_frames; ++f) {
    const TsHdr h = hdr[f];
    if (status && tid == 0) status[f] = h.status;
    if (h.status != 0 && h.status != 3) { if (dlen && tid == 0) dlen[f] = 0; continue; }      // dropped: the state stays (uniform branch)
    for (int j = tid; j < PKT + 4; j += blockDim.x) old_buf[j] = new_buf[j];
    if (tid == 0) { crc_in = S.crc; crc_out = S.crc; tsg_plan_frame(S, h.status == 3, h.dfl, h.syncd, P); }
    __syncthreads();
    const uint8_t* me = frames + (size_t)f * k_bch;
    const long long base = off;
    for (int g = warp; g < P.n_seg; g += 8) {                           // one warp per segment
      const TsgSeg sg = P.seg[g];
      for (int j = lane; j < sg.n; j += 32) {
        const uint8_t v = sg.kind == TSG_DATA ? (uint8_t)byte_at(me, k_bch, sg.src + 8 * j) : sg.kind == TSG_SYNC ? 0x47
                        : sg.kind == TSG_FILL ? 0xF0 : old_buf[min(sg.src + j, PKT + 3)];
        if (sg.to_buffer) { if (sg.dst + j < PKT + 4) new_buf[sg.dst + j] = v; }
        else if (base + sg.dst + j < cap) out[base + sg.dst + j] = v;
      }
    }
    __syncthreads();
    if (tid < P.n_task) {                                               // one thread per CRC task (they write disjoint bytes or the same bit)
      const TsgTask t = P.task[tid];
      if (t.check == -2) { if (t.tei >= 0 && base + t.tei < cap) out[base + t.tei] |= 0x80; }
      else {
        unsigned crc = t.chain ? crc_in : 0u;
        for (int j = 0; j < t.n; ++j) crc = tsg_crc8_byte(crc, byte_at(me, k_bch, t.src + 8 * j));
        if (t.check >= 0) {
          if (byte_at(me, k_bch, t.check) != crc && t.tei >= 0 && base + t.tei < cap) out[base + t.tei] |= 0x80;
        }
        if (tid == P.n_task - 1) crc_out = t.check >= 0 ? 0u : crc;      // the last task leaves the carried CRC
      }
    }
    __syncthreads();
    if (tid == 0) {
      // (a flagged resynchronisation as the last task leaves the CRC as it was: crc_out was preset to crc_in; if an
      // earlier task of the frame reset it, that task was a check and the value is 0)
      if (P.n_task > 0 && P.task[P.n_task - 1].check == -2) { crc_out = crc_in; for (int i = 0; i < P.n_task - 1; ++i) if (P.task[i].check >= 0) crc_out = 0u; }
      S.crc = crc_out;
      if (dlen) dlen[f] = P.out_len;
      off = base + P.out_len;
    }
    __syncthreads();
  }
  if (tid == 0) {
    st->split = S.split; st->idx_packet = S.idx_packet; st->idx_buffer = S.idx_buffer; st->crc = S.crc; st->total = off;
    st->tail_src = -1; st->tail_bit = 0; st->tail_ndata = 0; st->tail_sync = -1;
  }
  for (int j = tid; j < PKT + 4; j += blockDim.x) st->buffer[j] = new_buf[j];
}

}  // namespace

struct TsState { std::map<int, TsDevState*> plp; };

void t2_ts_free(t2b200_ctx* ctx)
{
  if (!ctx->ts) return;
  for (auto& kv : ctx->ts->plp) cudaFree(kv.second);
  delete ctx->ts;
  ctx->ts = nullptr;
}

static int ts_state(t2b200_ctx* ctx, int plp, TsDevState** out)
{
  if (!ctx->ts) ctx->ts = new TsState();
  auto it = ctx->ts->plp.find(plp);
  if (it == ctx->ts->plp.end()) {
    TsDevState* d;
    T2_CUDA(ctx, cudaMalloc(&d, sizeof(TsDevState)));
    T2_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(TsDevState), ctx->stream));
    it = ctx->ts->plp.emplace(plp, d).first;
  }
  *out = it->second;
  return T2B200_OK;
}

extern "C" int t2b200_ts_reset(t2b200_ctx* ctx, int plp)
{
  if (!ctx) return T2B200_ERR_ARG;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  TsDevState* st; int rc;
  if ((rc = ts_state(ctx, plp, &st))) return rc;
  T2_CUDA(ctx, cudaMemsetAsync(st, 0, sizeof(TsDevState), ctx->stream));
  return T2B200_OK;
}

extern "C" int t2b200_ts_packetize(t2b200_ctx* ctx, int plp, const uint8_t* bbframes, int n_frames, int k_bch,
                                   uint8_t* ts_out, size_t ts_cap, int32_t* datagram_len, int32_t* status,
                                   long long* total_out)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (!bbframes || n_frames < 0 || k_bch < 80 + PKT * 8 || k_bch > 65535 || !ts_out) { ctx->err = "t2b200_ts_packetize: bad argument"; return T2B200_ERR_ARG; }
  if (total_out) *total_out = 0;
  if (n_frames == 0) return T2B200_OK;
  T2_CUDA(ctx, cudaSetDevice(ctx->device));
  TsDevState* st; int rc;
  if ((rc = ts_state(ctx, plp, &st))) return rc;
  const void* din; void *dout, *dlen = nullptr, *dstat = nullptr, *dhdr, *ddesc;
  if ((rc = t2_to_device(ctx, 0, bbframes, (size_t)n_frames * k_bch, &din))) return rc;
  if ((rc = t2_out_device(ctx, 1, ts_out, ts_cap, &dout))) return rc;
  if (datagram_len && (rc = t2_out_device(ctx, 2, datagram_len, 4 * (size_t)n_frames, &dlen))) return rc;
  if (status && (rc = t2_out_device(ctx, 3, status, 4 * (size_t)n_frames, &dstat))) return rc;
  if ((rc = t2_dev_scratch(ctx, 9, sizeof(TsHdr) * (size_t)n_frames, &dhdr))) return rc;
  if ((rc = t2_dev_scratch(ctx, 10, sizeof(TsDesc) * (size_t)n_frames, &ddesc))) return rc;
  T2_CUDA(ctx, cudaMemsetAsync(&st->general, 0, sizeof(int), ctx->stream));
  ts_parse_kernel<<<(n_frames + 127) / 128, 128, 0, ctx->stream>>>((const uint8_t*)din, n_frames, k_bch, (TsHdr*)dhdr, st);
  T2_CUDA(ctx, cudaGetLastError());
  ts_scan_kernel<<<1, 32, 0, ctx->stream>>>((const TsHdr*)dhdr, n_frames, (TsDesc*)ddesc, st, (int32_t*)dlen, (int32_t*)dstat);
  T2_CUDA(ctx, cudaGetLastError());
  ts_assemble_kernel<<<n_frames, 256, 0, ctx->stream>>>((const uint8_t*)din, k_bch, (const TsDesc*)ddesc, st,
                                                        (uint8_t*)dout, (long long)ts_cap);
  T2_CUDA(ctx, cudaGetLastError());
  ts_save_tail_kernel<<<1, 256, 0, ctx->stream>>>((const uint8_t*)din, k_bch, st);
  T2_CUDA(ctx, cudaGetLastError());
  ts_general_kernel<<<1, 256, 0, ctx->stream>>>((const uint8_t*)din, n_frames, k_bch, (const TsHdr*)dhdr, st, (uint8_t*)dout,
                                                (long long)ts_cap, (int32_t*)dlen, (int32_t*)dstat);
  T2_CUDA(ctx, cudaGetLastError());
  ctx->launches += 5;
  if (total_out) {
    T2_CUDA(ctx, cudaMemcpyAsync(total_out, &st->total, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if ((size_t)*total_out > ts_cap) { ctx->err = "t2b200_ts_packetize: ts_out too small"; return T2B200_ERR_ARG; }
  }
  if ((rc = t2_finish_out(ctx, ts_out, dout, total_out ? (size_t)*total_out : ts_cap))) return rc;
  if (datagram_len && (rc = t2_finish_out(ctx, datagram_len, dlen, 4 * (size_t)n_frames))) return rc;
  if (status && (rc = t2_finish_out(ctx, status, dstat, 4 * (size_t)n_frames))) return rc;
  return T2B200_OK;
}
