// Shared pieces of the drop-in stage classes (host/dropin/DSP, host/dropin/DVB_T2): one t2b200 context per process and the
// error policy.  The headers next to this one carry the reference's own file names, class names and public signatures
// (src/DSP/fast_fourier_transform.h, src/DVB_T2/{data_symbol,fc_symbol,time_deinterleaver,llr_demapper,ldpc_decoder,
// bch_decoder}.h), so the reference's unmodified dvbt2_demodulator.cpp / p1_symbol.cpp / p2_symbol.cpp / bb_de_header.cpp
// compile against them; INTEGRATION.md has the recipe.  Every execute() is a t2b200_* call on the GPU; there is no CPU
// fallback: without a usable device the first stage constructed terminates the receiver with a message.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "t2b200.h"

namespace t2b200_dropin {

inline t2b200_ctx* context()
{
  static t2b200_ctx* ctx = [] {
    t2b200_ctx* c = nullptr;
    const char* dev = std::getenv("T2B200_DEVICE");
    if (t2b200_create(dev ? std::atoi(dev) : 0, &c) != T2B200_OK) {
      std::fprintf(stderr, "t2b200: no usable CUDA device -- the GPU stages have no CPU fallback\n");
      std::abort();
    }
    return c;
  }();
  return ctx;
}

inline void check(int rc, const char* what)
{
  if (rc != T2B200_OK) {
    std::fprintf(stderr, "t2b200: %s failed: %s\n", what, t2b200_last_error(context()));
    std::abort();
  }
}

}  // namespace t2b200_dropin

// Qt builds run moc over the stage headers (Q_OBJECT + signals) as they do for the reference's own; a build without moc
// (a Qt-free shim such as the one the parity tests use) defines T2B200_NO_MOC and gets empty bodies for the GUI-only signals.
#ifdef T2B200_NO_MOC
#define T2B200_SIGNAL_BODY {}
#else
#define T2B200_SIGNAL_BODY ;
#endif
