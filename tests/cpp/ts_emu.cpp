// Host execution of the general (normal-mode / mixed-mode) path of the TS re-packetiser: the frame plan of
// sdr_receiver_dvb_t2_b200/csrc/ts_general.h (the same source the GPU runs) plus a serial executor of its segments and
// CRC tasks.  Built and compared with the oracle port by tests/test_ts_general_emu.py.
#include <cstring>
#include <vector>
#include "../../sdr_receiver_dvb_t2_b200/csrc/ts_general.h"

namespace {
struct Emu { TsgState s; uint8_t buffer[192]; };
unsigned byte_at(const uint8_t* bits, int k_bch, int bit)
{
  unsigned v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 1) | ((bit + i < k_bch ? bits[bit + i] : 0) & 1u);      // past the frame: zero
  return v;
}
}

extern "C" {
int emu_ts_state_size() { return (int)sizeof(Emu); }
void emu_ts_reset(Emu* e) { std::memset(e, 0, sizeof(*e)); }

// one frame -> datagram length, or -1 (header CRC) / -2 (SYNCD 65535) / -4 (data field longer than the frame) / -9 (plan overflow)
int emu_ts_frame(Emu* e, const uint8_t* bits, int k_bch, uint8_t* out, int out_cap)
{
  unsigned reg = 0;
  for (int i = 0; i < 80; ++i) { const unsigned b = (bits[i] ^ reg) & 1u; reg >>= 1; if (b) reg ^= 0xABu; }
  if (reg != 0 && reg != 0xABu) return -1;
  int dfl = 0, syncd = 0;
  for (int i = 0; i < 16; ++i) { dfl = (dfl << 1) | (bits[32 + i] & 1); syncd = (syncd << 1) | (bits[56 + i] & 1); }
  if (syncd == 65535) return -2;
  if (80 + dfl > k_bch) return -4;
  static TsgPlan P;
  uint8_t old[192];
  std::memcpy(old, e->buffer, sizeof(old));
  const unsigned crc_in = e->s.crc;
  tsg_plan_frame(e->s, reg == 0, dfl, syncd, P);
  if (P.overflow) return -9;
  for (int i = 0; i < P.n_seg; ++i) { const TsgSeg& g = P.seg[i]; if (g.dst < 0 || g.dst + g.n > (g.to_buffer ? 192 : out_cap) || (g.kind == TSG_OLDBUF && g.src + g.n > 192)) return -10 - i; }
  for (int i = 0; i < P.n_seg; ++i) {
    const TsgSeg& g = P.seg[i];
    uint8_t* dst = (g.to_buffer ? e->buffer : out) + g.dst;
    for (int j = 0; j < g.n; ++j)
      dst[j] = g.kind == TSG_DATA ? (uint8_t)byte_at(bits, k_bch, g.src + 8 * j) : g.kind == TSG_SYNC ? 0x47 : g.kind == TSG_FILL ? 0xF0 : old[g.src + j];
  }
  unsigned crc_out = crc_in;
  for (int i = 0; i < P.n_task; ++i) {
    const TsgTask& t = P.task[i];
    if (t.check == -2) { if (t.tei >= 0) out[t.tei] |= 0x80; continue; }                    // crc untouched
    unsigned crc = t.chain ? crc_in : 0;
    for (int j = 0; j < t.n; ++j) crc = tsg_crc8_byte(crc, byte_at(bits, k_bch, t.src + 8 * j));
    if (t.check >= 0) {
      if (byte_at(bits, k_bch, t.check) != crc && t.tei >= 0) out[t.tei] |= 0x80;
      crc_out = 0;
    } else crc_out = crc;
  }
  e->s.crc = crc_out;
  return P.out_len;
}
}
extern "C" int emu_ts_plan_dump(int split, int idx_packet, int idx_buffer, int nm, int dfl, int syncd, int* segs, int* state_out)
{
  static TsgPlan P; TsgState s; s.split = split; s.idx_packet = idx_packet; s.idx_buffer = idx_buffer; s.crc = 0;
  tsg_plan_frame(s, nm, dfl, syncd, P);
  for (int i = 0; i < P.n_seg; ++i) { segs[5*i]=P.seg[i].kind; segs[5*i+1]=P.seg[i].to_buffer; segs[5*i+2]=P.seg[i].dst; segs[5*i+3]=P.seg[i].src; segs[5*i+4]=P.seg[i].n; }
  state_out[0]=s.split; state_out[1]=s.idx_packet; state_out[2]=s.idx_buffer; state_out[3]=P.out_len;
  return P.n_seg;
}
