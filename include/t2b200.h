/*
 * t2b200 -- C-ABI of the B200-native DVB-T2 demodulation + FEC hot path.
 *
 * Plain C: opaque context, plain pointers and sizes, int status returns.  Every entry point names
 * the reference interface it replaces (paths relative to Oleg-Malyutin/sdr_receiver_dvb_t2 src/).
 * The C++ facade classes in sdr_receiver_dvb_t2_b200/host/ keep the reference's per-stage
 * signatures and forward to these calls; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Pointers: every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer of
 * the context's GPU; the library detects which (cudaPointerGetAttributes) and stages host buffers
 * through its own pinned/device scratch on the context's stream.  All calls are asynchronous with
 * respect to device memory and synchronous with respect to host memory they were given: when a
 * call that received a host OUTPUT pointer returns, that buffer is filled.  t2b200_sync() drains
 * the stream.  There is no CPU fallback: without a usable GPU every compute call fails with
 * T2B200_ERR_CUDA.
 *
 * Threads: a context is used by one host thread at a time; use one context (and stream) per thread -- they share the GPU
 * (bench.py runs two).  Kernels of different contexts overlap: the LDPC decoder leaves room on every SM for the streaming
 * stages of another context.
 */
#ifndef T2B200_H
#define T2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t2b200_ctx t2b200_ctx;

enum {
  T2B200_OK = 0,
  T2B200_ERR_ARG = 1,       /* bad argument (null pointer, unknown code, size mismatch)        */
  T2B200_ERR_CUDA = 2,      /* CUDA runtime error or no device; see t2b200_last_error()        */
  T2B200_ERR_STATE = 3,     /* tables for this stage were not configured                        */
  T2B200_ERR_NOMEM = 4
};

/* dvbt2_definition.h:61-85 -- same numeric values as the reference enums */
enum { T2B200_C1_2 = 0, T2B200_C3_5, T2B200_C2_3, T2B200_C3_4, T2B200_C4_5, T2B200_C5_6 };
enum { T2B200_MOD_QPSK = 0, T2B200_MOD_16QAM, T2B200_MOD_64QAM, T2B200_MOD_256QAM };
enum { T2B200_FEC_SHORT = 0, T2B200_FEC_NORMAL = 1 };
/* extra LDPC code ids beyond fec/rate (tables present in the reference, never instantiated there:
 * LDPC/dvb_t2_tables.hh:874,1198,1232) */
enum { T2B200_CODE_SHORT_1_4 = 12, T2B200_CODE_SHORT_B8 = 13, T2B200_CODE_SHORT_B9 = 14 };

/* flags of t2b200_ldpc_decode */
enum {
  /* Reference batch semantics (ldpc_decoder.cpp:262-268, LDPC/layered_decoder.hh:174): codewords
   * are decoded in lock-step groups of 32, every lane iterates until ALL 32 pass the parity test or
   * 25 trials are spent.  Without it every codeword stops on its own (native mode).              */
  T2B200_LDPC_GROUP32 = 1,
  /* Fuse bch_decoder::execute (bch_decoder.cpp:139-142): emit only the first K_bch bits of each
   * word, XORed with the BB-scrambler PRBS (bch_decoder.cpp:50-61).                              */
  T2B200_LDPC_BCH_DESCRAMBLE = 2,
  /* Pack output bits MSB-first, 8 per byte (row stride = ceil(bits/8)); default is the reference
   * layout, one byte per bit (ldpc_decoder.cpp:270-277).                                        */
  T2B200_LDPC_PACK_BITS = 4,
  /* Also return the int8 posteriors (debug / parity tests): post_out int8[n][N].                */
  T2B200_LDPC_WANT_POST = 8
};

/* ---- context ------------------------------------------------------------------------------ */
int  t2b200_create(int device, t2b200_ctx** out);
void t2b200_destroy(t2b200_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t as void*); NULL restores the context's own */
int  t2b200_set_stream(t2b200_ctx* ctx, void* cuda_stream);
int  t2b200_sync(t2b200_ctx* ctx);
const char* t2b200_last_error(const t2b200_ctx* ctx);
const char* t2b200_version(void);
/* Options.  Everything defaults to the reference's behaviour.
 * T2B200_OPT_DEMAP_SATURATE (default 0): the reference converts LLRs with a C cast that WRAPS modulo 256
 * (llr_demapper.cpp:722-737; 193.f -> -63).  For 256-QAM its decision-directed precision is never below
 * ~116, so the outer constellation levels always wrap and the reference's own LDPC stage cannot converge on
 * a clean AWGN signal (DESIGN.md "reference quirks").  1 = clamp to [-128,127] instead -- NOT bit-compatible
 * with the reference, provided so the engine is usable; parity tests run with 0.                         */
enum { T2B200_OPT_DEMAP_SATURATE = 1, T2B200_OPT_LDPC_PLAIN_LAUNCH = 2, T2B200_OPT_BCH_CORRECT = 3, T2B200_OPT_STAGE_TIMING = 4 };
/* T2B200_OPT_LDPC_PLAIN_LAUNCH (default 0): lock-step (GROUP32) decodes are launched cooperatively, which makes a decode
 * wait until the whole GPU is free.  1 = ordinary launch of the same grid: with TWO contexts on two streams taking turns
 * (bench.py, chain.py) the next decode starts on the SMs the previous one's last groups have left.  Do not run more than
 * two such decodes concurrently on one GPU: partially resident groups of a third could starve the second (the kernel
 * traps after a few seconds rather than hang).  Results are identical either way.                              */
/* T2B200_OPT_BCH_CORRECT (default 0): the reference never decodes the BCH code ("TODO BCH decode", bch_decoder.cpp:136).
 * 1 = t2b200_frames_decode corrects up to t bit errors per BBFRAME (t2b200_bch_decode) between the LDPC stage and the
 * parity strip / descramble -- beyond the reference; off in every parity test.                                          */
/* T2B200_OPT_STAGE_TIMING (default 0): 1 = t2b200_frames_decode brackets every stage with CUDA events on the context's
 * stream; t2b200_frames_stage_ms reads the device times of the last call (it waits for that call to finish). */
int t2b200_set_option(t2b200_ctx* ctx, int option, int value);
/* number of kernels this library launched on the context since creation (bench.py: gpu_launches) */
long long t2b200_launch_count(const t2b200_ctx* ctx);

/* ---- K5 + K6: LDPC decode (+ BCH-parity strip and BB descramble) -------------------------- */
/* LDPC code geometry (ldpc_decoder.cpp:177-245, bch_decoder.cpp:79-131). code = ldpc code id:
 * fec_type(0 short,1 normal) and code_rate map to an id with t2b200_ldpc_code_id().            */
int t2b200_ldpc_code_id(int fec_type, int code_rate);
int t2b200_ldpc_n(int code);        /* 64800 / 16200 */
int t2b200_ldpc_k(int code);        /* K_ldpc */
int t2b200_ldpc_k_bch(int code);    /* K_bch (0 for the L1-only codes) */

/*
 * Replaces ldpc_decoder::execute (ldpc_decoder.cpp:157-301) -> LDPCDecoder::operator()
 * (LDPC/layered_decoder.hh:168-180) and, with T2B200_LDPC_BCH_DESCRAMBLE, bch_decoder::execute
 * (bch_decoder.cpp:63-164).
 *   llr         int8[n_codewords][N], codeword (transmitted) order, positive => bit 0
 *   bits_out    per codeword K (or K_bch) bits; one byte per bit, or packed (flag)
 *   trials_left int32[n_codewords] or NULL: the reference's `count` for the codeword's group
 *               (>= 0 converged, < 0 "could not recover": the reference drops the whole group)
 *   iterations  int32[n_codewords] or NULL: update() passes executed
 *   post_out    int8[n_codewords][N] or NULL (needs T2B200_LDPC_WANT_POST)
 *   max_trials  reference value 25 (ldpc_decoder.h:62); 0 < max_trials <= 60
 * With GROUP32 a trailing partial group (n_codewords % 32 lanes) is decoded as a smaller lock-step
 * group; the reference never sees one (llr_demapper.cpp:749-765 only emits full batches).
 */
int t2b200_ldpc_decode(t2b200_ctx* ctx, int code, const int8_t* llr, int n_codewords,
                       uint8_t* bits_out, int32_t* trials_left, int32_t* iterations,
                       int8_t* post_out, int max_trials, unsigned flags);

/* Replaces bch_decoder::execute alone (bch_decoder.cpp:63-164) for callers that kept the
 * byte-per-bit LDPC output: out[w][i] = in[w][i] ^ prbs[i], i < K_bch.                          */
int t2b200_bch_descramble(t2b200_ctx* ctx, int code, const uint8_t* bits_in, int n_words,
                          uint8_t* bits_out);

/* N3, opt-in: true BCH decoding, which the reference leaves as a TODO (bch_decoder.cpp:136).  bits_inout:
 * uint8[n_words][K_ldpc], one byte per bit (what t2b200_ldpc_decode emits without BCH_DESCRAMBLE): the K_ldpc = N_bch bits
 * of each word are decoded in place -- shortened BCH over GF(2^16) / GF(2^14), t = t2b200_bch_t(code) (EN 302 755 6.1.1).
 * corrected int32[n_words] or NULL: bit errors corrected (0 .. t), -1 = more than t errors, word left unchanged.         */
int t2b200_bch_t(int code);
int t2b200_bch_decode(t2b200_ctx* ctx, int code, uint8_t* bits_inout, int n_words, int32_t* corrected);

/* ---- K3: time / cell de-interleaver + cyclic-Q-delay removal ------------------------------- */
/* Host-side table builders (no GPU needed):
 * cell de-interleaver permutation of time_deinterleaver::address_cell_deinterleaving
 * (time_deinterleaver.cpp:174-266): perm_out int32[n_fec_blocks * cells_per_fec];
 * bit de-interleaver + demux address table of llr_demapper::address_generator
 * (llr_demapper.cpp:110-130, constants llr_demapper.h:64-78): address_out int32[64800 | 16200].        */
int t2b200_cell_permutation(int n_fec_blocks, int cells_per_fec, int32_t* perm_out);
int t2b200_demap_address_table(int fec_type, int mod, int code_rate, int32_t* address_out);
/* frequency de-interleaver tables of address_freq_deinterleaver (address_freq_deinterleaver.cpp:28-209; EN 302 755 8.5):
 * h_even_out / h_odd_out int32[n_cells] for a symbol kind with n_cells cells (c_p2, c_data or n_fc); fft_size 16384 or
 * 32768.  The drop-in may hand the reference's own tables to t2b200_eq_configure instead.                               */
int t2b200_freq_deinterleaver_table(int fft_size, int n_cells, int32_t* h_even_out, int32_t* h_odd_out);

/* Replaces time_deinterleaver::start for one PLP (time_deinterleaver.cpp:38-145): geometry from
 * (fec_type, mod), permutation for plp_num_blocks_max FEC blocks.  permutation == NULL builds it
 * natively; the drop-in facade may pass the reference's own table.                                     */
int t2b200_ti_configure(t2b200_ctx* ctx, int plp, int fec_type, int mod, int n_fec_blocks_max,
                        const int32_t* permutation);

/* Replaces the cell loop of time_deinterleaver::execute (time_deinterleaver.cpp:316-374) for whole TI
 * blocks: cells_in = complex<float> cells of n_ti_blocks consecutive TI blocks in arrival order
 * (block b holds n_fec_per_block[b] * cells_per_fec cells), cells_out = the same blocks de-interleaved,
 * with the Q component moved back one cell inside each FEC block (unconditionally, like the reference). */
int t2b200_ti_deinterleave(t2b200_ctx* ctx, int plp, const float* cells_in, int n_ti_blocks,
                           const int32_t* n_fec_per_block, float* cells_out);

/* ---- K4: soft demapper + bit de-interleave / demux ------------------------------------------ */
/* Replaces llr_demapper::execute -> qpsk/qam16/qam64/qam256 (llr_demapper.cpp:132-768) for
 * n_ti_blocks TI blocks of one PLP:
 *   ti_cells      complex<float>, de-interleaved TI blocks back to back; DEROTATED IN PLACE when
 *                 rotation != 0, exactly like the reference (llr_demapper.cpp:555-557)
 *   llr_out       int8[sum(n_fec_per_block)][64800 | 16200], codeword order, ready for t2b200_ldpc_decode
 *   snr_out       float[n_ti_blocks] or NULL: the value the reference emits as signal_noise_ratio
 *   precision_out float[n_ti_blocks] or NULL: 8*a*sum_s/sum_e actually used
 *   precision_in  float[n_ti_blocks] or NULL: override (parity tests pin the one order-dependent
 *                 float reduction of this stage with it)                                              */
int t2b200_demap(t2b200_ctx* ctx, float* ti_cells, int n_ti_blocks, const int32_t* n_fec_per_block,
                 int mod, int rotation, int fec_type, int code_rate, int8_t* llr_out,
                 float* snr_out, float* precision_out, const float* precision_in);

/* ---- K2: channel estimation + equalisation + frequency de-interleaving ---------------------- */
/* kind: 0 = P2 symbol, 1 = data symbols, 2 = frame-closing symbol */
enum { T2B200_SYM_P2 = 0, T2B200_SYM_DATA = 1, T2B200_SYM_FC = 2 };

/* Replaces p2_symbol::init / data_symbol::init / fc_symbol::init (p2_symbol.cpp:43-76,
 * data_symbol.cpp:39-106, fc_symbol.cpp:37-80): hands the engine the init-time tables that the
 * reference's pilot_generator and address_freq_deinterleaver objects hold (pilot_generator.h:28-33,
 * address_freq_deinterleaver.h:33-38), which the drop-in classes receive as arguments.
 *   n_symbols     symbols of this kind per T2 frame (n_data for DATA, 1 for P2 / FC)
 *   first_symbol  frame index of the first one (0 for P2, n_p2 for DATA, len_frame-1 for FC)
 *   carrier_map   int32[n_symbols][k_total]  carrier types (dvbt2_definition.h:103-113)
 *   pilot_refer   float[n_symbols][k_total]  +-amplitude for pilots, 0 elsewhere
 *   h_even/h_odd  int32[>= n_out]            de-interleaver tables; the symbol's parity picks one
 *   amp_main      amp_p2 (P2) or amp_sp (DATA, FC); amp_cp for continual pilots (DATA only)          */
int t2b200_eq_configure(t2b200_ctx* ctx, int kind, int n_symbols, int first_symbol, int fft_size,
                        int k_total, int l_nulls, int n_out, const int32_t* carrier_map,
                        const float* pilot_refer, const int32_t* h_even, const int32_t* h_odd,
                        float amp_main, float amp_cp);

/* Replaces the equaliser of p2_symbol::execute (p2_symbol.cpp:89-259), data_symbol::execute
 * (data_symbol.cpp:108-335) and fc_symbol::execute (fc_symbol.cpp:82-271) for a batch of symbols:
 *   idx_symbol int32[n]  frame index of each symbol (parity selects h_odd / h_even, index selects the map)
 *   freq       complex<float>[n][fft_size]  FFT output, halves swapped (fast_fourier_transform::execute)
 *   cells_out  complex<float>[n][n_out]     equalised, frequency-de-interleaved cells
 *   sro, phase float[n] or NULL             the two feedback estimates the reference returns by reference
 * n = 1 is the synchronous per-symbol call of live reception.                                        */
int t2b200_equalize(t2b200_ctx* ctx, int kind, int n_symbols, const int32_t* idx_symbol,
                    const float* freq, float* cells_out, float* sro, float* phase);

/* ---- a2: transmission-mode parameters and pilot tables, built natively ------------------------ */
/* The subset of dvbt2_parameters (dvbt2_definition.h:215-260) the hot path needs, for SISO 16K / 32K.  Enum values are
 * the reference's: fft_mode 4 = 16K, 5 = 32K (dvbt2_definition.h:121-131); carrier_mode 0 normal / 1 extended;
 * pilot_pattern 0..7 = PP1..PP8; guard_interval_mode 0..6 = 1/32 1/16 1/8 1/4 1/128 19/128 19/256; papr_mode 0..3.    */
typedef struct {
  int fft_mode, carrier_mode, pilot_pattern, guard_interval_mode, papr_mode;
  int fft_size, k_total, k_ext, k_offset, l_nulls, guard_interval_size;
  int n_p2, c_p2, c_data, n_fc, c_fc, l_fc;
  int n_data, len_frame;          /* n_data = symbols behind P2 including the frame-closing one; len_frame = n_p2 + n_data */
  int dx, dy;                     /* scattered-pilot spacing of the pilot pattern                                       */
  float amp_p2, amp_sp, amp_cp;   /* pilot boosts (pilot_generator.cpp:376-507)                                         */
} t2b200_mode;
/* Replaces dvbt2_p2_parameters_init + dvbt2_bwt_ext_parameters_init + dvbt2_data_parameters_init
 * (dvbt2_definition.cpp:20-159,161-648).  T2B200_ERR_ARG for combinations EN 302 755 does not define (c_data == 0).     */
int t2b200_mode_init(int fft_mode, int carrier_mode, int pilot_pattern, int guard_interval_mode, int n_data,
                     int papr_mode, t2b200_mode* out);
/* Replaces pilot_generator::p2_generator / data_generator (pilot_generator.cpp:69-132): the carrier-type map
 * (dvbt2_definition.h:103-113) and the BPSK pilot reference (+-amplitude, 0 on data carriers) of one symbol kind.
 *   kind T2B200_SYM_P2 / _FC: carrier_map int32[k_total], pilot_refer float[k_total]
 *   kind T2B200_SYM_DATA:     [len_frame - l_fc - n_p2][k_total] each, one row per data symbol                         */
int t2b200_pilot_tables(const t2b200_mode* mode, int kind, int32_t* carrier_map, float* pilot_refer);
/* t2b200_eq_configure for P2, data and frame-closing symbols of a mode from natively built pilot and frequency
 * de-interleaver tables: what p2_symbol::init / data_symbol::init / fc_symbol::init set up (p2_symbol.cpp:43-76,
 * data_symbol.cpp:39-106, fc_symbol.cpp:37-80) without the reference's pilot_generator / address_freq_deinterleaver.  */
int t2b200_eq_configure_mode(t2b200_ctx* ctx, const t2b200_mode* mode);

/* ---- K1: OFDM FFT --------------------------------------------------------------------------- */
/* Replaces fast_fourier_transform::init + execute (DSP/fast_fourier_transform.h:54-70) for a batch of
 * symbols: out[b] = halves-swapped, unnormalised forward DFT (FFTW_FORWARD sign) of in[b].
 *   n      a power of two from 256 to 32768: 16K / 32K OFDM symbols (the modes the reference runs) and the 1K transform of
 *          the P1 symbol (p1_symbol.cpp:34-35,114)
 *   in     complex<float>[batch][n]  the n samples after the guard interval (dvbt2_demodulator.cpp:332)
 *   out    complex<float>[batch][n]  carrier k of the active band sits at index l_nulls + k
 * FFTW is a binary-only dependency of the reference: agreement is to <= 1e-5 * max|X| (float64 DFT). */
int t2b200_fft(t2b200_ctx* ctx, int n, const float* in, int batch, float* out);

/* ---- N1: BBFRAME -> transport stream ---------------------------------------------------------- */
/* Replaces bb_de_header::execute (bb_de_header.cpp:84-445) for a batch of BBFRAMEs of ONE PLP (the caller applies the
 * reference's need_plp filter): high-efficiency mode (:332-428) and normal mode (:166-331: the sync byte on air is the
 * CRC-8 of the previous packet, checked, replaced by 0x47, a mismatch sets that packet's transport_error_indicator).
 * Batches of high-efficiency frames are built by a parallel scan; a batch with a normal-mode frame (also one a header bit
 * error turned into normal mode) is switched on the device to a frame-by-frame path that follows the reference's
 * byte-serial loop exactly, its quirks included: the CRC bytes are read without being counted against DFL (a frame reads a
 * few bytes behind its data field; past the frame they read as zero), and a too-short SYNCD after a held-back tail leaves
 * the packet index beyond 188 for good (:208-226).
 *   bbframes      uint8[n_frames][k_bch], one byte per bit: what t2b200_ldpc_decode(BCH_DESCRAMBLE) or
 *                 bch_decoder::execute emit (bch_decoder.cpp:139-160)
 *   ts_out        the datagrams the reference would send (bb_de_header.cpp:431-441), back to back; writes are bounded by
 *                 ts_cap (a header with an absurd SYNCD makes the reference write past its datagram buffer)
 *   datagram_len  int32[n_frames] or NULL: bytes of each frame's datagram (0 for a dropped frame)
 *   status        int32[n_frames] or NULL: 0 high-efficiency frame; 3 normal-mode frame; dropped frames: 1 header CRC-8
 *                 error (:108-113); 2 SYNCD == 65535 (:160-163); 4 the header announces a data field longer than the
 *                 frame (80 + DFL > k_bch) -- dropped without touching the carried state; the reference has no such
 *                 check and would read past its buffer
 *   total_out     bytes written to ts_out, or NULL (then the call stays asynchronous for device buffers)
 * The packet phase and the held-back tail (< 188 bytes) persist per PLP between calls; t2b200_ts_reset clears them. */
int t2b200_ts_reset(t2b200_ctx* ctx, int plp);
int t2b200_ts_packetize(t2b200_ctx* ctx, int plp, const uint8_t* bbframes, int n_frames, int k_bch,
                        uint8_t* ts_out, size_t ts_cap, int32_t* datagram_len, int32_t* status, long long* total_out);

/* ---- whole frames in one call (replay mode) ----------------------------------------------------- */
/* The per-stage calls above chained on the device for whole T2 frames of one PLP: what dvbt2_demodulator::
 * symbol_acquisition (dvbt2_demodulator.cpp:332-385) and the signal chain behind it (time_deinterleaver -> llr_demapper ->
 * ldpc_decoder -> bch_decoder) do symbol by symbol, for already synchronised frames.  Configure the symbol tables
 * (t2b200_eq_configure for P2 / DATA / FC) and the PLP (t2b200_ti_configure) first.                                  */
typedef struct {
  int fft_size, len_frame, n_p2, l_fc;   /* symbols of a frame: n_p2 P2 symbols, data symbols, l_fc frame-closing symbol   */
  int c_p2, c_data, n_fc;                /* cells per symbol of each kind (dvbt2_parameters, dvbt2_definition.h:215-260)  */
  int first_cell;                        /* first PLP cell of the frame cell stream: 1840 + l1_post_size + cells of the    */
                                         /* PLPs in front (time_deinterleaver.cpp:44, l1_postsignalling_dynamic.start)    */
  int plp, mod, rotation, fec_type, code_rate, n_blocks, ti_len;   /* l1_postsignalling_plp + this frame's num_blocks       */
} t2b200_frame_cfg;
int t2b200_frames_configure(t2b200_ctx* ctx, const t2b200_frame_cfg* cfg);
/*   iq          complex<float>[n_frames][len_frame][fft_size]: the FFT window of every symbol (dvbt2_demodulator.cpp:332)
 *   bits_out    [n_frames * n_blocks][K_bch | K_ldpc] per ldpc_flags (T2B200_LDPC_GROUP32 | _BCH_DESCRAMBLE | _PACK_BITS)
 *   trials_left int32[n_frames * n_blocks] or NULL;  sro, phase float[n_frames][len_frame] or NULL (the equalisers'
 *   feedback floats);  snr float[n_frames * ti_len] or NULL.  Device buffers keep the call asynchronous.                 */
int t2b200_frames_decode(t2b200_ctx* ctx, const float* iq, int n_frames, uint8_t* bits_out, int32_t* trials_left,
                         float* sro, float* phase, float* snr, int max_trials, unsigned ldpc_flags);
/* The same with the FFT windows as the device front-ends deliver samples (rx_sdrplay.cpp:246: int16 I and Q), interleaved
 *   iq   int16[n_frames][len_frame][fft_size][2]  (I, Q); sample = (I + jQ) * scale
 * converted on the device while the FFT loads them -- the first step of dvbt2_demodulator::execute
 * (dvbt2_demodulator.cpp:182-186, short_to_float = 2^-14 / 2^-12 / 2^-11 by device).  Half the bytes over PCIe and HBM. */
int t2b200_frames_decode_i16(t2b200_ctx* ctx, const int16_t* iq, float scale, int n_frames, uint8_t* bits_out,
                             int32_t* trials_left, float* sro, float* phase, float* snr, int max_trials, unsigned ldpc_flags);
/* Device time of every stage of the last t2b200_frames_decode[_i16] call made with T2B200_OPT_STAGE_TIMING on, in ms:
 * [0] FFT, [1] equalise + frequency de-interleave, [2] time / cell de-interleave (+ derotation), [3] demap,
 * [4] LDPC + BCH strip / descramble (+ opt-in BCH correction), [5] the whole call.  Measurement aid of bench.py. */
int t2b200_frames_stage_ms(t2b200_ctx* ctx, float ms_out[6]);


/* ---- N2: receiver front-end, the step in front of the FFT -------------------------------------------- */
/* One chunk (about one OFDM symbol) of int16 I/Q of each of n_streams independent streams goes through what
 * dvbt2_demodulator::execute does per chunk (dvbt2_demodulator.cpp:178-221): int16 -> float, DC removal
 * (DSP/loop_filters.hh:58-73), the 1-bit IQ-imbalance statistics (:256-265) and correction (:191-192), the NCO
 * derotation (:194-212, sin / cos tables of DSP/fast_math.h), the Farrow resampler (DSP/interpolator_farrow.hh:41-68) and
 * the half-band decimator (DSP/filter_decimator.h:72-131).  The five per-sample recurrences of the reference are evaluated
 * in closed form per chunk (csrc/frontend_kernels.h), so a launch is parallel over samples and streams; the loop filters
 * and the symbol state machine stay the host's (the reference's dvbt2_demodulator, which hands the loop values in here).
 * Parity: the NCO phase is reproduced exactly; samples agree with the reference to 2e-6 of the signal RMS (it is built
 * -Ofast; DC average, resampler phase and statistics are summed in another order), chunk lengths exactly.
 *   t2b200_frontend_configure   n_streams streams, chunks of at most max_chunk_in input samples (<= 262 144); resets them
 *   t2b200_frontend_reset       dvbt2_demodulator::reset (:111-127) for one stream (-1: all): zero state, x1 = -0.5
 *   t2b200_frontend_execute     one chunk per stream:
 *     i_in, q_in   sample n of stream s at i_in[s * stream_stride + n * sample_step] (the reference's two pointers and its
 *                  convert_input, dvbt2_demodulator.cpp:181-183); host or device memory
 *     chunk        [n_streams] (host): what the host loop knows, see t2b200_fe_chunk
 *     out          complex<float>[n_streams][out_stride]: the decimator output (out_decimator, :220)
 *     result       [n_streams] (host): number of samples written, the chunk's contribution to theta1..3 (:256-265)
 *   t2b200_frontend_get_state / _set_state   the carried state of one stream (tests, checkpointing)
 *   t2b200_cp_correlate         guard-interval correlation of dvbt2_demodulator.cpp:321-330 for n_symbols buffered symbols
 *                               (guard + fft_size samples each, symbol_stride apart): frequency_est[n_symbols]          */
typedef struct {
  int   len_in;                   /* chunk, dvbt2_demodulator.cpp:165-167 */
  float short_to_float;           /* 2^-14 (sdrplay), 2^-12 (airspy), 2^-11 (plutosdr), :35-52 */
  float c1, c2;                   /* IQ-imbalance correction of this execute() call, :240-243 */
  float frequency_est_filtered;   /* NCO decrement per input sample, :194 */
  float phase_nco;                /* after the per-chunk update of :170-176 */
  float resample;                 /* (float)arbitrary_resample, :162-163, interpolator_farrow.hh:45 */
} t2b200_fe_chunk;
typedef struct { int len_out, len_interp; float theta1, theta2, theta3; } t2b200_fe_result;
typedef struct {
  float dc_re, dc_im, frequency_nco, x1;
  float delay[3][2];              /* the last three derotated samples, newest first (delay_data_1.._3) */
  float hist[63][2];              /* the 63 resampler outputs in front of the next one, oldest first (the decimator's buffer) */
  int   parity, pad;              /* filter_decimator::execute's static d */
} t2b200_fe_state;
int t2b200_frontend_configure(t2b200_ctx* ctx, int n_streams, int max_chunk_in);
int t2b200_frontend_reset(t2b200_ctx* ctx, int stream);
int t2b200_frontend_execute(t2b200_ctx* ctx, const int16_t* i_in, const int16_t* q_in, long long stream_stride, int sample_step,
                            const t2b200_fe_chunk* chunk, float* out, long long out_stride, t2b200_fe_result* result);
int t2b200_frontend_get_state(t2b200_ctx* ctx, int stream, t2b200_fe_state* state);
int t2b200_frontend_set_state(t2b200_ctx* ctx, int stream, const t2b200_fe_state* state);
int t2b200_cp_correlate(t2b200_ctx* ctx, const float* symbols, int n_symbols, long long symbol_stride, int fft_size, int guard,
                        float* frequency_est);
/* The sliding correlator of p1_symbol::execute (p1_symbol.cpp:75-178, block diagram :56-74): for every sample of a block the
 * product of the two branch sums (`out`) and its squared magnitude (`correlation`), which the reference compares with its
 * begin / end thresholds.  Closed form (window sums as differences of prefix sums) instead of the reference's delay lines and
 * running sums: agreement to 1e-5 of the correlation peak.  Stateless: the caller hands over what the reference keeps in its
 * buffers -- the 2046 samples in front of the block (NULL: zeros, i.e. after reset_buffer) and the position of the block's
 * first sample in the 1024-step frequency-shift table (the reference's static idx_fq_shift).  The threshold state machine, the
 * 1K FFT of the detected symbol (t2b200_fft) and the S1 / S2 decoding stay with the caller.
 *   samples complex<float>[n]; history complex<float>[2046] or NULL; correlation float[n]; out complex<float>[n] or NULL */
int t2b200_p1_correlate(t2b200_ctx* ctx, const float* samples, int n, const float* history, int fq_index,
                        float* correlation, float* out);

/* ---- multi-GPU: the LDPC / BCH stage sharded by codeword over the GPUs of one box --------------------- */
/* FEC blocks are independent, so one rank demodulates (it holds the int8 LLRs of a pooled batch), every rank decodes a
 * contiguous shard of whole 32-codeword groups and the BBFRAME bits return to that rank: ONE exchange each way, NCCL
 * send / recv over NVLink issued by the library on a side stream, in chunks of <= 1152 codewords double-buffered against
 * the decode.  One context (= one GPU) per process; NCCL is loaded at run time (libnccl.so.2).
 *   t2b200_comm_unique_id   rank 0 obtains the rendezvous id (ncclGetUniqueId, 128 bytes) and hands it to the other
 *                           ranks by any means (the tests use torch.distributed, the reference has no such step)
 *   t2b200_comm_init        every rank: joins the communicator
 *   t2b200_ldpc_decode_sharded  collective, every rank calls it with the same code / n_cw / flags / root:
 *     llr, bits_out  DEVICE memory on `root` (int8[n_cw][N] in, [n_cw][K_bch | K_ldpc] per flags out), ignored elsewhere
 *     flags          as t2b200_ldpc_decode (GROUP32 | BCH_DESCRAMBLE | PACK_BITS); asynchronous on the context's stream */
#define T2B200_COMM_ID_BYTES 128
int t2b200_comm_unique_id(void* id_out, size_t id_bytes);
int t2b200_comm_init(t2b200_ctx* ctx, int rank, int nranks, const void* unique_id, size_t id_bytes);
int t2b200_comm_destroy(t2b200_ctx* ctx);
int t2b200_ldpc_decode_sharded(t2b200_ctx* ctx, int code, int root, const int8_t* llr, int n_codewords, uint8_t* bits_out,
                               int max_trials, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* T2B200_H */
