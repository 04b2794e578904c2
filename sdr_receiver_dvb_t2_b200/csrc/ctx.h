// Internal context of libt2b200 (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include "../../include/t2b200.h"

struct LdpcDeviceCode;   // ldpc.cu
struct FftPlan;          // fft.cu
struct SymbolTables;     // equalizer.cu
struct TiDemapState;     // demap.cu
struct TsState;          // ts.cu
struct FramePipe;        // frames.cu
struct CommState;        // comm.cu
struct BchState;         // bch.cu
struct FeState;          // frontend.cu

struct Scratch {
  void* p = nullptr; size_t cap = 0; bool pinned_host = false;
};

struct t2b200_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;      // copy streams of the host-buffer pipelines
  cudaEvent_t ev_in[2] = {}, ev_k[2] = {}, ev_out[2] = {};
  std::string err;
  long long launches = 0;
  int opt_demap_saturate = 0;                 // T2B200_OPT_DEMAP_SATURATE
  int opt_ldpc_plain_launch = 0;              // T2B200_OPT_LDPC_PLAIN_LAUNCH
  int opt_bch_correct = 0;                    // T2B200_OPT_BCH_CORRECT
  int opt_stage_timing = 0;                   // T2B200_OPT_STAGE_TIMING
  int ldpc_slots_cap = 0;                     // > 0: lock-step decodes use at most this many group slots (16 CTAs each): the sharded
                                              // FEC stage leaves SMs to the NCCL kernels that way (comm.cu)
  std::map<int, LdpcDeviceCode*> ldpc;        // by code id
  float* d_lut = nullptr;                     // sin | cos tables of DSP/fast_math.h, 2 x 65536 floats
  uint8_t* d_prbs = nullptr;                  // BB descrambler PRBS, 54000 bytes
  unsigned* d_group_sync = nullptr; size_t group_sync_cap = 0;
  unsigned* d_err_flag = nullptr;             // raised by a kernel that gave up a lock-step wait (ldpc.cu)
  cudaEvent_t ev_h2d = nullptr;               // completion of the last copy out of a pinned host source
  std::map<int, FftPlan*> fft;                // by log2 n
  SymbolTables* sym[3] = {nullptr, nullptr, nullptr};
  TiDemapState* ti = nullptr;
  TsState* ts = nullptr;
  FramePipe* frames = nullptr;
  BchState* bch = nullptr;                    // GF(2^16) / GF(2^14) tables of the opt-in BCH decoder
  FeState* fe = nullptr;                      // receiver front-end streams (t2b200_frontend_configure)
  float* d_p1_fq = nullptr;                   // p1_symbol's frequency-shift table, 1024 x {sin, cos}
  CommState* comm = nullptr;                  // NCCL communicator of the sharded FEC stage (t2b200_comm_init)
  // staging scratch, grown on demand
  Scratch dev[24];
  Scratch pin[8];
};

#define T2_CUDA(ctx, call)                                                            \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
      return T2B200_ERR_CUDA;                                                         \
    }                                                                                 \
  } while (0)

// The LDPC decoder asks for the largest shared-memory carve-out: two resident decoders then leave 81 KB of shared memory
// (next to a quarter of the registers and 1 280 threads) on which the streaming kernels of another stream run.
#define T2_CARVEOUT(kernel) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)

// true if p is device-accessible memory of this process (device or managed)
bool t2_is_device_ptr(const void* p);
// grow-on-demand scratch; slot identifies the user
int t2_dev_scratch(t2b200_ctx* ctx, int slot, size_t bytes, void** out);
int t2_pin_scratch(t2b200_ctx* ctx, int slot, size_t bytes, void** out);
// Bring `bytes` at src (host or device) into device memory; returns a device pointer (src itself if
// already on the device). Async on ctx->stream; pageable host memory is staged through pinned scratch.
int t2_to_device(t2b200_ctx* ctx, int slot, const void* src, size_t bytes, const void** dptr);
// Output side: returns a device pointer to write to (dst itself if device memory, else scratch)
int t2_out_device(t2b200_ctx* ctx, int slot, void* dst, size_t bytes, void** dptr);
// Copy a scratch-produced result back to a host dst and wait for it (no-op if dst was device memory)
int t2_finish_out(t2b200_ctx* ctx, void* dst, const void* dptr, size_t bytes);
// T2B200_ERR_CUDA (and ctx->err) if a kernel raised the device error flag since the last check; synchronous
int t2_check_device_flag(t2b200_ctx* ctx);
// the {cos, sin} tables of DSP/fast_math.h on the device (ctx->d_lut), built on first use (equalizer.cu)
int t2_ensure_lut(t2b200_ctx* ctx);
