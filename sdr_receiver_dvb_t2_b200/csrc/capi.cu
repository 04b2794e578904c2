// Context management and host<->device staging of the t2b200 C-ABI (include/t2b200.h).
#include "ctx.h"

void t2_ldpc_free(t2b200_ctx* ctx);
void t2_fft_free(t2b200_ctx* ctx);
void t2_eq_free(t2b200_ctx* ctx);
void t2_ti_free(t2b200_ctx* ctx);
void t2_ts_free(t2b200_ctx* ctx);
void t2_frames_free(t2b200_ctx* ctx);
void t2_comm_free(t2b200_ctx* ctx);
void t2_bch_free(t2b200_ctx* ctx);
void t2_fe_free(t2b200_ctx* ctx);

bool t2_is_device_ptr(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static bool is_pinned_host(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int t2_dev_scratch(t2b200_ctx* ctx, int slot, size_t bytes, void** out)
{
  Scratch& s = ctx->dev[slot];
  if (s.cap < bytes) {
    // the old block may still be in use by queued work on the stream
    T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (s.p) cudaFree(s.p);
    s.p = nullptr; s.cap = 0;
    size_t cap = bytes + bytes / 4 + 256;
    T2_CUDA(ctx, cudaMalloc(&s.p, cap));
    s.cap = cap;
  }
  *out = s.p;
  return T2B200_OK;
}

int t2_pin_scratch(t2b200_ctx* ctx, int slot, size_t bytes, void** out)
{
  Scratch& s = ctx->pin[slot];
  if (s.cap < bytes) {
    T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (s.p) cudaFreeHost(s.p);
    s.p = nullptr; s.cap = 0;
    size_t cap = bytes + bytes / 4 + 256;
    T2_CUDA(ctx, cudaMallocHost(&s.p, cap));
    s.cap = cap;
  }
  *out = s.p;
  return T2B200_OK;
}

int t2_to_device(t2b200_ctx* ctx, int slot, const void* src, size_t bytes, const void** dptr)
{
  if (t2_is_device_ptr(src)) { *dptr = src; return T2B200_OK; }
  void* d; int rc;
  if ((rc = t2_dev_scratch(ctx, slot, bytes, &d))) return rc;
  // Pageable memory is staged by the driver: cudaMemcpyAsync returns once the source has been consumed.  A PINNED source is
  // read by the copy engine later, so the call waits for that copy (only the copy, not the kernels behind it): when an
  // entry point returns, the caller may reuse every host buffer it passed in -- the contract of this ABI.
  T2_CUDA(ctx, cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (is_pinned_host(src)) {
    T2_CUDA(ctx, cudaEventRecord(ctx->ev_h2d, ctx->stream));
    T2_CUDA(ctx, cudaEventSynchronize(ctx->ev_h2d));
  }
  *dptr = d;
  return T2B200_OK;
}

int t2_out_device(t2b200_ctx* ctx, int slot, void* dst, size_t bytes, void** dptr)
{
  if (t2_is_device_ptr(dst)) { *dptr = dst; return T2B200_OK; }
  return t2_dev_scratch(ctx, slot, bytes, dptr);
}

int t2_finish_out(t2b200_ctx* ctx, void* dst, const void* dptr, size_t bytes)
{
  if (dst == dptr) return T2B200_OK;
  T2_CUDA(ctx, cudaMemcpyAsync(dst, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return T2B200_OK;
}

// A kernel that had to give up a wait (ldpc.cu: lock-step lanes that never became co-resident) raises a flag in device
// memory instead of trapping; every call that waits for the GPU anyway turns it into an error here.
int t2_check_device_flag(t2b200_ctx* ctx)
{
  unsigned f = 0;
  T2_CUDA(ctx, cudaMemcpy(&f, ctx->d_err_flag, sizeof(f), cudaMemcpyDeviceToHost));
  if (!f) return T2B200_OK;
  cudaMemset(ctx->d_err_flag, 0, sizeof(unsigned));
  ctx->err = "LDPC decoder: a lock-step wait timed out (lanes of a group were not co-resident); results of that call are invalid";
  return T2B200_ERR_CUDA;
}

extern "C" {

const char* t2b200_version(void) { return "t2b200 0.2 (sm_100a)"; }

int t2b200_create(int device, t2b200_ctx** out)
{
  if (!out) return T2B200_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return T2B200_ERR_CUDA;      // no CPU fallback: the caller must fail
  }
  t2b200_ctx* ctx = new t2b200_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return T2B200_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return T2B200_ERR_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return T2B200_ERR_CUDA; }
  ctx->stream = ctx->own_stream;
  if (cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc(&ctx->d_err_flag, sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(ctx->d_err_flag, 0, sizeof(unsigned)) != cudaSuccess) { t2b200_destroy(ctx); return T2B200_ERR_CUDA; }
  *out = ctx;
  return T2B200_OK;
}

void t2b200_destroy(t2b200_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  t2_ldpc_free(ctx);
  t2_fft_free(ctx);
  t2_eq_free(ctx);
  t2_ti_free(ctx);
  t2_ts_free(ctx);
  t2_frames_free(ctx);
  t2_comm_free(ctx);
  t2_bch_free(ctx);
  t2_fe_free(ctx);
  if (ctx->d_prbs) cudaFree(ctx->d_prbs);
  if (ctx->d_group_sync) cudaFree(ctx->d_group_sync);
  if (ctx->d_err_flag) cudaFree(ctx->d_err_flag);
  if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
  for (auto& s : ctx->dev) if (s.p) cudaFree(s.p);
  for (auto& s : ctx->pin) if (s.p) cudaFreeHost(s.p);
  if (ctx->s_in) {
    cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_k[i]); cudaEventDestroy(ctx->ev_out[i]); }
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int t2b200_set_stream(t2b200_ctx* ctx, void* cuda_stream)
{
  if (!ctx) return T2B200_ERR_ARG;
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return T2B200_OK;
}

int t2b200_sync(t2b200_ctx* ctx)
{
  if (!ctx) return T2B200_ERR_ARG;
  T2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return t2_check_device_flag(ctx);
}

int t2b200_set_option(t2b200_ctx* ctx, int option, int value)
{
  if (!ctx) return T2B200_ERR_ARG;
  if (option == T2B200_OPT_DEMAP_SATURATE) { ctx->opt_demap_saturate = value != 0; return T2B200_OK; }
  if (option == T2B200_OPT_LDPC_PLAIN_LAUNCH) { ctx->opt_ldpc_plain_launch = value != 0; return T2B200_OK; }
  if (option == T2B200_OPT_BCH_CORRECT) { ctx->opt_bch_correct = value != 0; return T2B200_OK; }
  if (option == T2B200_OPT_STAGE_TIMING) { ctx->opt_stage_timing = value != 0; return T2B200_OK; }
  ctx->err = "t2b200_set_option: unknown option";
  return T2B200_ERR_ARG;
}

const char* t2b200_last_error(const t2b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
long long t2b200_launch_count(const t2b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

// stages not built yet keep their free hooks here so the context teardown stays in one place
__attribute__((weak)) void t2_fft_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_eq_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_ti_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_ts_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_frames_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_comm_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_bch_free(t2b200_ctx*) {}
__attribute__((weak)) void t2_fe_free(t2b200_ctx*) {}
