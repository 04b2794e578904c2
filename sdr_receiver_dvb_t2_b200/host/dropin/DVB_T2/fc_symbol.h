// Drop-in for the reference's src/DVB_T2/fc_symbol.h (frame-closing symbol): same class, same signatures
// (fc_symbol.h:31-33).  init() takes the tables data_symbol::init left in the pilot generator / address objects
// (fc_symbol.cpp:38-80), execute() is t2b200_equalize (fc_symbol.cpp:82-271).
#ifndef FC_SYMBOL_H
#define FC_SYMBOL_H

#include <QObject>
#include <vector>

#include "dvbt2_definition.h"
#include "pilot_generator.h"
#include "address_freq_deinterleaver.h"
#include "t2b200_dropin.h"

class fc_symbol : public QObject
{
    Q_OBJECT
public:
    explicit fc_symbol(QObject* parent = nullptr) : QObject(parent) {}
    ~fc_symbol() {}

    complex *execute(complex* _ofdm_cell, float &_sample_rate_offset, float &_phase_offset)
    {
        int32_t idx = idx_symbol;
        t2b200_dropin::check(t2b200_equalize(t2b200_dropin::context(), T2B200_SYM_FC, 1, &idx,
                                             reinterpret_cast<const float*>(_ofdm_cell),
                                             reinterpret_cast<float*>(deinterleaved_cell.data()), &_sample_rate_offset,
                                             &_phase_offset), "t2b200_equalize(fc)");
        return deinterleaved_cell.data();
    }

    void init(dvbt2_parameters _dvbt2, pilot_generator *_pilot, address_freq_deinterleaver *_address)
    {
        float amp_sp = 7.0f / 3.0f;                        // fc_symbol.cpp:47-63
        switch (_dvbt2.pilot_pattern) {
        case PP1: case PP2: amp_sp = 4.0f / 3.0f; break;
        case PP3: case PP4: amp_sp = 7.0f / 4.0f; break;
        default: break;
        }
        idx_symbol = _dvbt2.len_frame - 1;                 // fc_symbol.cpp:64
        std::vector<int32_t> map(_pilot->fc_carrier_map, _pilot->fc_carrier_map + _dvbt2.k_total);
        t2b200_dropin::check(t2b200_eq_configure(t2b200_dropin::context(), T2B200_SYM_FC, 1, idx_symbol, _dvbt2.fft_size,
                                                 _dvbt2.k_total, _dvbt2.l_nulls, _dvbt2.n_fc, map.data(), _pilot->fc_pilot_refer,
                                                 _address->h_even_fc, _address->h_odd_fc, amp_sp, 0.0f), "t2b200_eq_configure(fc)");
        deinterleaved_cell.assign(static_cast<size_t>(_dvbt2.n_fc), complex());
    }

signals:
    void replace_spectrograph(const int _len_data, complex* _data) T2B200_SIGNAL_BODY
    void replace_constelation(const int _len_data, complex* _data) T2B200_SIGNAL_BODY
    void replace_oscilloscope(const int _len_data, complex* _data) T2B200_SIGNAL_BODY

private:
    int idx_symbol = 0;
    std::vector<complex> deinterleaved_cell;
};

#endif // FC_SYMBOL_H
