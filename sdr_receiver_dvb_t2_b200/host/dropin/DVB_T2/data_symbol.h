// Drop-in for the reference's src/DVB_T2/data_symbol.h: same class name, same init / execute signatures
// (data_symbol.h:31-34).  init() runs the reference's own table builders exactly as data_symbol.cpp:67-98 does
// (dvbt2_data_parameters_init, pilot_generator::data_generator, address_freq_deinterleaver::data_address_freq_deinterleaver
// -- the objects it is handed stay the reference's) and uploads their tables; execute() is t2b200_equalize on the GPU:
// pilot estimation, angle / amplitude interpolation, derotation, frequency de-interleaving (data_symbol.cpp:108-335), with
// the two feedback floats returned synchronously.
#ifndef DATA_SYMBOL_H
#define DATA_SYMBOL_H

#include <QObject>
#include <vector>

#include "dvbt2_definition.h"
#include "pilot_generator.h"
#include "address_freq_deinterleaver.h"
#include "t2b200_dropin.h"

class data_symbol : public QObject
{
    Q_OBJECT
public:
    explicit data_symbol(QObject* parent = nullptr) : QObject(parent) {}
    ~data_symbol() {}

    complex *execute(int _idx_symbol, complex* _ofdm_cell, float &_sample_rate_offset, float &_phase_offset)
    {
        swap_buffer = !swap_buffer;                        // callee-owned ping-pong output, data_symbol.cpp:140-147
        complex* out = (swap_buffer ? buffer_a : buffer_b).data();
        int32_t idx = _idx_symbol;
        t2b200_dropin::check(t2b200_equalize(t2b200_dropin::context(), T2B200_SYM_DATA, 1, &idx,
                                             reinterpret_cast<const float*>(_ofdm_cell), reinterpret_cast<float*>(out),
                                             &_sample_rate_offset, &_phase_offset), "t2b200_equalize(data)");
        return out;
    }

    void init(dvbt2_parameters &_dvbt2, pilot_generator* _pilot, address_freq_deinterleaver* _address)
    {
        float amp_cp = 8.0f / 3.0f, amp_sp = 7.0f / 3.0f;  // data_symbol.cpp:48-84
        switch (_dvbt2.fft_mode) {
        case FFTSIZE_1K: case FFTSIZE_2K: amp_cp = 4.0f / 3.0f; break;
        case FFTSIZE_4K: amp_cp = (4.0f * sqrtf(2)) / 3.0f; break;
        default: break;
        }
        dvbt2_data_parameters_init(_dvbt2);
        switch (_dvbt2.pilot_pattern) {
        case PP1: case PP2: amp_sp = 4.0f / 3.0f; break;
        case PP3: case PP4: amp_sp = 7.0f / 4.0f; break;
        default: break;
        }
        _pilot->data_generator(_dvbt2);
        _address->data_address_freq_deinterleaver(_dvbt2);
        const int n_sym = _dvbt2.len_frame - _dvbt2.l_fc - _dvbt2.n_p2, k = _dvbt2.k_total;
        std::vector<int32_t> map(static_cast<size_t>(n_sym) * k);
        std::vector<float> ref(map.size());
        for (int s = 0; s < n_sym; ++s)
            for (int i = 0; i < k; ++i) {
                map[static_cast<size_t>(s) * k + i] = _pilot->data_carrier_map[s][i];
                ref[static_cast<size_t>(s) * k + i] = _pilot->data_pilot_refer[s][i];
            }
        t2b200_dropin::check(t2b200_eq_configure(t2b200_dropin::context(), T2B200_SYM_DATA, n_sym, _dvbt2.n_p2, _dvbt2.fft_size, k,
                                                 _dvbt2.l_nulls, _dvbt2.c_data, map.data(), ref.data(), _address->h_even_data,
                                                 _address->h_odd_data, amp_sp, amp_cp), "t2b200_eq_configure(data)");
        buffer_a.assign(static_cast<size_t>(_dvbt2.c_data), complex());
        buffer_b.assign(static_cast<size_t>(_dvbt2.c_data), complex());
    }

signals:
    void replace_spectrograph(const int _len_data, complex* _data) T2B200_SIGNAL_BODY
    void replace_constelation(const int _len_data, complex* _data) T2B200_SIGNAL_BODY
    void replace_oscilloscope(const int _len_data, complex* _data) T2B200_SIGNAL_BODY

private:
    std::vector<complex> buffer_a, buffer_b;
    bool swap_buffer = false;
};

#endif // DATA_SYMBOL_H
