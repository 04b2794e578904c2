// Host-side construction of the layered-decoding schedule for one DVB-T2 LDPC code.
// Re-derives what LDPCDecoder::init builds (reference LDPC/layered_decoder.hh:115-167, from the
// quasi-cyclic enumeration in LDPC/ldpc.hh:39-123) in closed form: because every table address x
// of bit-group g feeds check (x + q*m) mod R for bit m of the group, check node (layer i = x mod q,
// row j) reads bit 360*g + ((j - x/q) mod 360).  So a layer is described by a short list of
// (group base, cyclic shift) pairs instead of a 360-row position table.
#pragma once
#include <cstdint>
#include <vector>

struct T2LdpcCodeData { const char* name; int N, K, nrows; const uint8_t* rowdeg; const uint16_t* addr; };
const T2LdpcCodeData* t2_ldpc_code_data(int code);   // nullptr if out of range
int t2_ldpc_k_bch(int code);

struct LdpcSchedule {
  int code = -1, N = 0, K = 0, R = 0, q = 0;
  int cnl_max = 0;                 // most data edges on any check node (LINKS_MAX_CN - 2)
  int links_total = 0;
  std::vector<uint8_t> cnt;        // [q] data edges per check node of layer i
  std::vector<uint32_t> edge;      // [q][cnl_max]  (360*g + shift) in the low 16 bits | (360 - shift) << 16:
                                   //   CN (i,j) edge c reads posterior 360*g + (j + shift) mod 360
                                   //   = j + low16 - (j >= high16 ? 360 : 0)
  std::vector<uint32_t> shared;    // [q] bit c set: data edge c of this layer reads a bit-group that another
                                   //   edge of the same layer reads too (two check nodes share each such bit)
  std::vector<uint8_t> ns;         // [q] number of such edges; they are slots 0 .. ns-1 (with ns == 2: slot 0 takes the
                                   //   bit from an earlier check node of the serial order, slot 1 hands it to a later one)
  // Exact emulation of the reference's serial j = 0..359 order inside a layer: two check nodes of
  // one layer that share a bit must run smaller-j first.  level[][] is the longest-chain depth.
  std::vector<int16_t> conflict_index;  // [q] row into level[], or -1 when the layer has no shared bit
  std::vector<uint8_t> nlev;            // [q] number of levels (1 when conflict-free)
  std::vector<uint8_t> level;           // [n_conflict_layers][360], 1-based
  int total_substeps = 0;
};

bool t2_build_ldpc_schedule(int code, LdpcSchedule& s);
