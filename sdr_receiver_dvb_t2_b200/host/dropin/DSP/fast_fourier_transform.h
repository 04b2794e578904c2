// Drop-in for the reference's src/DSP/fast_fourier_transform.h (same class, same three methods): the forward FFT +
// half swap of one OFDM symbol (or the 1K P1 symbol) on the GPU through t2b200_fft instead of FFTW.
//   init(len)  -> the input buffer the caller memcpy's the samples into (fast_fourier_transform.h:54-62)
//   execute()  -> the shifted spectrum, valid until the next call          (fast_fourier_transform.h:64-70)
#ifndef FAST_FOURIER_TRANSFORM_H
#define FAST_FOURIER_TRANSFORM_H

#include <complex>
#include <vector>
#include "t2b200_dropin.h"

typedef std::complex<float> complex;

class fast_fourier_transform
{
public:
    fast_fourier_transform() {}
    ~fast_fourier_transform() {}

    complex* init(int _len_in)
    {
        t2b200_dropin::context();
        len = _len_in;
        in.assign(static_cast<size_t>(len), complex());
        out.assign(static_cast<size_t>(len), complex());
        return in.data();
    }

    complex* execute()
    {
        t2b200_dropin::check(t2b200_fft(t2b200_dropin::context(), len, reinterpret_cast<const float*>(in.data()), 1,
                                        reinterpret_cast<float*>(out.data())), "t2b200_fft");
        return out.data();
    }

private:
    int len = 0;
    std::vector<complex> in, out;
};

#endif // FAST_FOURIER_TRANSFORM_H
