// TEST INFRASTRUCTURE ONLY -- tap around the reference's Farrow resampler.  The reference's own header is included
// unmodified under another class name (REF_FARROW_HH / REF_DECIMATOR_H: its path as a string literal, given by oracle/Makefile) and wrapped: the wrapper
// has the same layout (one member), delegates every call and then reports the call to oracle/ref_chain.cc, which records
// it when a tap is armed.  dvbt2_demodulator.cpp compiles against this header because oracle/tap precedes the reference
// tree on the include path; nothing of the reference is copied.
#ifndef T2B200_TAP_INTERPOLATOR_FARROW_HH
#define T2B200_TAP_INTERPOLATOR_FARROW_HH
#define interpolator_farrow ref_interpolator_farrow
#include REF_FARROW_HH
#undef interpolator_farrow

extern "C" void oracle_tap_farrow(int len_in, const void* in, double resample, int len_out, const void* out);

template<typename T, typename F>
class interpolator_farrow
{
public:
    ref_interpolator_farrow<T, F> impl;          // public: the tap harness reads its state
    void operator()(int len_in_, T* in_, double &arbitrary_resample_, int &len_out_, T* out_)
    {
        impl(len_in_, in_, arbitrary_resample_, len_out_, out_);
        oracle_tap_farrow(len_in_, in_, arbitrary_resample_, len_out_, out_);
    }
};
#endif
