// N1, the general path of the BBFRAME -> TS re-packetiser: normal-mode frames, batches that mix the two modes
// (a header CRC bit-flip turns one mode's residue into the other's) and the irregular packet states normal mode can leave
// behind (a packet index that has run past 188 and never comes back, bb_de_header.cpp:208-226).  ts.cu keeps its parallel
// scan for pure high-efficiency-mode batches and switches to this path, on the device, when a batch needs it.
//
// The reference (bb_de_header.cpp:166-428) walks a frame byte by byte.  Here ONE thread turns a frame into a short plan
// without touching the data: copy SEGMENTS (runs of bytes between packet boundaries, sync bytes, 0xF0 fill, the held-back
// bytes of the previous frame) and CRC TASKS (normal mode: the CRC-8 of every packet piece, compared with the byte that
// follows it on air; a mismatch raises the transport_error_indicator of the packet header last written to the datagram).
// Segments and tasks are then executed in parallel (ts_general_kernel: one warp per segment, one thread per task).
// Compiles for the host as well: tests/cpp/ts_emu.cpp runs the plan + a serial executor against the oracle port.
//
// Reference behaviour kept as it is: the CRC bytes of normal mode are read without being counted against DFL, so a frame
// reads a few bytes past its data field (past the frame they read as zero, which is what the oracle pins); a held-back tail
// is opened with buffer[0] even when it is empty; the too-short-SYNCD resynchronisation copies SYNCD / 8 bytes and leaves
// the packet index beyond 188 for good.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TSG_FN __host__ __device__ __forceinline__
#else
#define TSG_FN static inline
#endif

enum { TSG_PKT = 188, TSG_MAX_SEG = 160, TSG_MAX_TASK = 80 };
enum { TSG_DATA = 0, TSG_SYNC = 1, TSG_FILL = 2, TSG_OLDBUF = 3 };

struct TsgState { int split, idx_packet, idx_buffer; unsigned crc; };       // + the 188-byte buffer, kept by the caller
struct TsgSeg { int kind, to_buffer, dst, src, n; };                         // src: bit position in the frame (DATA), byte index (OLDBUF)
struct TsgTask { int chain, src, n, check, tei; };   // crc = chain ? carried : 0; over n bytes at bit src; check: bit position of the
                                                     // byte to compare (-1: none, the crc is carried on; -2: flag unconditionally)
struct TsgPlan { int n_seg, n_task, out_len, overflow; TsgSeg seg[TSG_MAX_SEG]; TsgTask task[TSG_MAX_TASK]; };

struct TsgCursor {                 // what the byte-serial loop of the reference carries while it walks one frame
  int inb, o, b, tei, idx_packet;
  int t_open, t_src, t_n, t_chain; // the packet piece whose CRC is being accumulated
};

TSG_FN void tsg_seg(TsgPlan& P, int kind, int to_buffer, int dst, int src, int n)
{
  if (n <= 0) return;
  if (P.n_seg >= TSG_MAX_SEG) { P.overflow = 1; return; }
  TsgSeg& s = P.seg[P.n_seg++];
  s.kind = kind; s.to_buffer = to_buffer; s.dst = dst; s.src = src; s.n = n;
}
TSG_FN void tsg_close_task(TsgPlan& P, TsgCursor& c, int check, int tei)
{
  if (P.n_task >= TSG_MAX_TASK) { P.overflow = 1; return; }
  TsgTask& t = P.task[P.n_task++];
  t.chain = c.t_chain; t.src = c.t_src; t.n = c.t_n; t.check = check; t.tei = tei;
  c.t_chain = 0; c.t_n = 0; c.t_src = 0;
}

// a run of R data bytes of the main loop (to_buffer = 0, bb_de_header.cpp:264-330 / :404-428) or of the held-back tail
// (to_buffer = 1, :244-263 / :386-402), cut at the packet boundaries
TSG_FN void tsg_run(TsgPlan& P, TsgCursor& c, int R, int to_buffer, int normal_mode)
{
  while (R > 0) {
    const int boundary = c.idx_packet == TSG_PKT;
    if (boundary || (c.idx_packet == 0 && !to_buffer)) {
      if (boundary && normal_mode) { tsg_close_task(P, c, c.inb, c.tei); c.inb += 8; }   // the CRC byte on air, compared and skipped
      tsg_seg(P, TSG_SYNC, to_buffer, to_buffer ? c.b : c.o, 0, 1);
      if (to_buffer) ++c.b; else { ++c.o; c.tei = c.o; }
      c.idx_packet = 1;
    }
    const int n = c.idx_packet < TSG_PKT ? (R < TSG_PKT - c.idx_packet ? R : TSG_PKT - c.idx_packet) : R;
    tsg_seg(P, TSG_DATA, to_buffer, to_buffer ? c.b : c.o, c.inb, n);
    if (normal_mode) { if (c.t_n == 0) c.t_src = c.inb; c.t_n += n; }
    if (to_buffer) c.b += n; else c.o += n;
    c.inb += 8 * n; c.idx_packet += n; R -= n;
  }
}

// One frame.  normal_mode: the header's CRC-8 residue was 0 (else 0xAB, high-efficiency mode).  S is updated.
TSG_FN void tsg_plan_frame(TsgState& S, int normal_mode, int dfl, int syncd, TsgPlan& P)
{
  P.n_seg = 0; P.n_task = 0; P.out_len = 0; P.overflow = 0;
  TsgCursor c;
  c.inb = 80; c.o = 0; c.b = 0; c.tei = -1; c.idx_packet = S.idx_packet;
  c.t_open = 0; c.t_src = 0; c.t_n = 0; c.t_chain = 1;
  const int sb = syncd / 8;
  if (S.split) {
    S.split = 0;
    const int missing = TSG_PKT - c.idx_packet;
    if (normal_mode) {
      const int n_old = S.idx_buffer > 1 ? S.idx_buffer : 1;                 // :171-178: buffer[0] goes out in any case
      tsg_seg(P, TSG_OLDBUF, 0, c.o, 0, n_old);
      c.o += n_old; c.tei = 1;
      if (missing <= sb) {
        const int n = missing == sb ? missing : sb;                          // :183-226
        tsg_seg(P, TSG_DATA, 0, c.o, c.inb, n);
        if (n > 0) { c.t_src = c.inb; c.t_n = n; }
        c.o += n; c.inb += 8 * n; c.idx_packet += n;
        tsg_close_task(P, c, c.inb, c.tei);
        c.inb += 8;
      } else {                                                               // :227-248: no CRC over these bytes, flagged in any case
        tsg_seg(P, TSG_DATA, 0, c.o, c.inb, sb);
        c.o += sb; c.inb += 8 * sb;
        tsg_seg(P, TSG_FILL, 0, c.o, 0, missing - sb);
        c.o += missing - sb;
        c.idx_packet += missing;
        if (P.n_task < TSG_MAX_TASK) { TsgTask& t = P.task[P.n_task++]; t.chain = 2; t.src = 0; t.n = 0; t.check = -2; t.tei = c.tei; }
      }
    } else {
      tsg_seg(P, TSG_OLDBUF, 0, c.o, 0, S.idx_buffer);                        // :341-382
      c.o += S.idx_buffer > 0 ? S.idx_buffer : 0;
      if (missing <= sb) {
        tsg_seg(P, TSG_DATA, 0, c.o, c.inb, missing);
        const int m = missing > 0 ? missing : 0;
        c.o += m; c.inb += 8 * m; c.idx_packet += m;
        if (missing < sb) c.inb += syncd - missing * 8;
      } else {
        tsg_seg(P, TSG_DATA, 0, c.o, c.inb, sb);
        c.o += sb; c.inb += 8 * sb;
        tsg_seg(P, TSG_FILL, 0, c.o, 0, missing - sb);
        c.o += missing - sb;
        c.idx_packet += missing;
      }
    }
  } else {
    c.inb += syncd + (normal_mode ? 8 : 0);
  }
  int rest = dfl - syncd - (normal_mode ? 8 : 0);
  if (rest >= TSG_PKT * 8) {
    const int M = (rest - TSG_PKT * 8) / 8 + 1;
    tsg_run(P, c, M, 0, normal_mode);
    rest -= 8 * M;
  }
  if (rest > 0) {
    S.split = 1;
    tsg_run(P, c, rest / 8, 1, normal_mode);
    S.idx_buffer = c.b;
  }
  // the packet piece that is still open carries its CRC into the next frame (task without a check); chain == 2 marks the
  // "crc left as it was" case of the flagged resynchronisation
  if (normal_mode && (c.t_n > 0 || c.t_chain == 1)) tsg_close_task(P, c, -1, -1);
  S.idx_packet = c.idx_packet;
  P.out_len = c.o;
}

// CRC-8 of bb_de_header.cpp:54-68 (polynomial 0xD5, MSB first) over one byte
TSG_FN unsigned tsg_crc8_byte(unsigned crc, unsigned byte)
{
  unsigned x = (crc ^ byte) & 0xffu;
  for (int i = 0; i < 8; ++i) x = (x & 0x80u) ? ((x << 1) ^ 0xD5u) & 0xffu : (x << 1) & 0xffu;
  return x;
}
