// TEST INFRASTRUCTURE ONLY -- tap around the reference's half-band decimator (see interpolator_farrow.hh next to this
// file for the scheme): the reference's class compiled unmodified under another name, wrapped with the same layout.
#ifndef T2B200_TAP_FILTER_DECIMATOR_H
#define T2B200_TAP_FILTER_DECIMATOR_H
#define filter_decimator ref_filter_decimator
#include REF_DECIMATOR_H
#undef filter_decimator

extern "C" void oracle_tap_decimator(int len_in, const void* in, int len_out, const void* out);

class filter_decimator
{
public:
    ref_filter_decimator impl;          // public: the tap harness reads its state
    void execute(int _len_in, complex* _in, int &_len_out, complex* _out)
    {
        impl.execute(_len_in, _in, _len_out, _out);
        oracle_tap_decimator(_len_in, _in, _len_out, _out);
    }
};
#endif
