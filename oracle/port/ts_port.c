/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's BBFRAME -> transport-stream
 * re-packetiser.  Pinned against the compiled reference (oracle/_ref/libref_chain.so, ref_bb_deheader) by
 * tests/test_oracle_ts.py and against tests/golden/ts_ref.npz.  Never linked by the product.
 *
 * Follows (paths relative to /root/reference/src/DVB_T2):
 *   bb_de_header.cpp:70-82     check_crc8_mode: CRC-8 of the 80 header bits, residue 0 => normal mode,
 *                              0xAB => high-efficiency mode, anything else => frame dropped
 *   bb_de_header.cpp:84-164    header fields (only SIS/MIS, DFL, SYNCD steer the packetiser); SYNCD == 65535 => dropped
 *   bb_de_header.cpp:166-331   normal mode: the sync byte on air is the CRC-8 of the previous packet; checked,
 *                              replaced by 0x47, a mismatch sets the transport_error_indicator of that packet
 *   bb_de_header.cpp:332-428   high-efficiency mode: 187-byte packets on air, 0x47 re-inserted
 *   both: bytes are emitted while at least 188 bytes of data field remain; the rest (< 188 bytes) is held back
 *   and opens the next frame's datagram, completed by SYNCD / 8 bytes of that frame (three resync cases).
 * One call = one BBFRAME = one UDP datagram (bb_de_header.cpp:431-441).
 */
#include <stdint.h>
#include <string.h>

#define PKT 188

typedef struct {
  int idx_packet, idx_buffer, split;
  uint8_t crc;
  uint8_t buffer[PKT];
  uint8_t crc_table[256];
  int table_ready;
} port_ts_state;

static void crc_table_init(port_ts_state* s)
{
  for (int i = 0; i < 256; ++i) {                       /* bb_de_header.cpp:54-68: poly 0xD5, MSB first */
    int crc = 0;
    for (int j = 7; j >= 0; --j) {
      int in = (i >> j) & 1, top = (crc >> 7) & 1;
      crc = (crc << 1);
      if (in ^ top) crc ^= 0xD5;
    }
    s->crc_table[i] = (uint8_t)crc;
  }
  s->table_ready = 1;
}

int port_ts_state_size(void) { return (int)sizeof(port_ts_state); }
void port_ts_reset(port_ts_state* s) { memset(s, 0, sizeof(*s)); crc_table_init(s); }

static unsigned take(const uint8_t** in, int nbits)
{
  unsigned v = 0;
  for (int i = 0; i < nbits; ++i) v = (v << 1) | (*(*in)++ & 1u);
  return v;
}

/* bits: one byte per bit, k_bch of them.  out: datagram bytes.  Returns the datagram length (0 is possible), -1 when
 * the frame is dropped because of its header CRC, -2 when dropped because SYNCD == 65535. */
int port_ts_frame(port_ts_state* s, const uint8_t* bits, int len_in, uint8_t* out)
{
  (void)len_in;
  if (!s->table_ready) crc_table_init(s);
  /* bit-serial CRC of the header, LSB-first register (bb_de_header.cpp:70-82) */
  uint8_t reg = 0;
  for (int i = 0; i < 80; ++i) {
    uint8_t b = bits[i] ^ (reg & 1);
    reg >>= 1;
    if (b) reg ^= 0xAB;
  }
  int hem;
  if (reg == 0) hem = 0; else if (reg == 0xAB) hem = 1; else return -1;
  const uint8_t* in = bits + 32;                         /* MATYPE (16) + UPL (16) */
  int dfl = (int)take(&in, 16);
  (void)take(&in, 8);                                    /* SYNC */
  int syncd = (int)take(&in, 16);
  if (syncd == 65535) return -2;
  in += 8;                                               /* CRC-8 */
  uint8_t* o = out;
  uint8_t* tei = NULL;                                   /* byte holding the transport_error_indicator of the open packet */
  const int syncd_byte = syncd / 8;

  if (!hem) {
    if (s->split) {
      s->split = 0;
      *o++ = s->buffer[0];
      tei = o;
      for (int i = 1; i < s->idx_buffer; ++i) *o++ = s->buffer[i];
      const int missing = PKT - s->idx_packet;
      if (missing <= syncd_byte) {
        const int n = missing == syncd_byte ? missing : syncd_byte;      /* :183-226: the second case copies SYNCD/8 bytes */
        for (int i = 0; i < n; ++i) { uint8_t t = (uint8_t)take(&in, 8); s->crc = s->crc_table[t ^ s->crc]; *o++ = t; ++s->idx_packet; }
        uint8_t t = (uint8_t)take(&in, 8);
        if (t != s->crc) *tei |= 0x80;
        s->crc = 0;
      } else {
        for (int i = 0; i < syncd_byte; ++i) { *o++ = (uint8_t)take(&in, 8); ++s->idx_packet; }
        for (int i = 0; i < missing - syncd_byte; ++i) { *o++ = 0xF0; ++s->idx_packet; }
        *tei |= 0x80;
      }
    } else {
      in += syncd + 8;
    }
    dfl -= syncd + 8;
    while (dfl > 0) {
      if (dfl < PKT * 8) {
        s->split = 1;
        s->idx_buffer = 0;
        for (int i = 0; i < dfl / 8; ++i) {
          if (s->idx_packet == PKT) {
            s->idx_packet = 0;
            uint8_t t = (uint8_t)take(&in, 8);
            if (t != s->crc && tei) *tei |= 0x80;
            s->crc = 0;
            s->buffer[s->idx_buffer++] = 0x47;
            ++s->idx_packet;
          }
          uint8_t t = (uint8_t)take(&in, 8);
          s->crc = s->crc_table[t ^ s->crc];
          s->buffer[s->idx_buffer++] = t;
          ++s->idx_packet;
        }
        dfl = 0;
      } else {
        if (s->idx_packet == PKT || s->idx_packet == 0) {
          if (s->idx_packet == PKT) {
            uint8_t t = (uint8_t)take(&in, 8);
            if (t != s->crc && tei) *tei |= 0x80;
            s->crc = 0;
          }
          s->idx_packet = 0;
          *o++ = 0x47; ++s->idx_packet;
          tei = o;
        }
        uint8_t t = (uint8_t)take(&in, 8);
        s->crc = s->crc_table[t ^ s->crc];
        *o++ = t; ++s->idx_packet;
        dfl -= 8;
      }
    }
  } else {
    if (s->split) {
      s->split = 0;
      for (int i = 0; i < s->idx_buffer; ++i) *o++ = s->buffer[i];
      const int missing = PKT - s->idx_packet;
      if (missing <= syncd_byte) {
        for (int i = 0; i < missing; ++i) { *o++ = (uint8_t)take(&in, 8); ++s->idx_packet; }
        if (missing < syncd_byte) in += syncd - missing * 8;
      } else {
        for (int i = 0; i < syncd_byte; ++i) { *o++ = (uint8_t)take(&in, 8); ++s->idx_packet; }
        for (int i = 0; i < missing - syncd_byte; ++i) { *o++ = 0xF0; ++s->idx_packet; }
      }
    } else {
      in += syncd;
    }
    dfl -= syncd;
    while (dfl > 0) {
      if (dfl < PKT * 8) {
        s->split = 1;
        s->idx_buffer = 0;
        for (int i = 0; i < dfl / 8; ++i) {
          if (s->idx_packet == PKT) { s->idx_packet = 0; s->buffer[s->idx_buffer++] = 0x47; ++s->idx_packet; }
          s->buffer[s->idx_buffer++] = (uint8_t)take(&in, 8);
          ++s->idx_packet;
        }
        dfl = 0;
      } else if (s->idx_packet == PKT || s->idx_packet == 0) {
        s->idx_packet = 0;
        *o++ = 0x47; ++s->idx_packet;
      } else {
        *o++ = (uint8_t)take(&in, 8); ++s->idx_packet;
        dfl -= 8;
      }
    }
  }
  return (int)(o - out);
}
