// Drop-in for the reference's src/DVB_T2/llr_demapper.h: same class, constructor and execute() signature
// (llr_demapper.h:30-46).  execute() = t2b200_demap on the GPU for one TI block: derotation, the decision-directed
// precision (ordered float sums reproduced exactly), max-log LLRs with the reference's wrapping int8 cast, bit
// de-interleave + demux (llr_demapper.cpp:132-776).  FECFRAMEs are handed on 32 at a time with the PLP id of each
// (soft_multiplexer_de_twist); a partial batch waits for the next TI block (llr_demapper.cpp:749-765) -- the PLP ids live
// in the object here, not in a stack array that dies with the call.
#ifndef LLR_DEMAPPER_H
#define LLR_DEMAPPER_H

#include <QObject>
#include <QThread>
#include <QWaitCondition>
#include <QMutex>
#include <complex>
#include <vector>

#include "dvbt2_definition.h"
#include "ldpc_decoder.h"
#include "t2b200_dropin.h"

typedef std::complex<float> complex;

class llr_demapper : public QObject
{
    Q_OBJECT
public:
    explicit llr_demapper(QWaitCondition *_signal_in, QMutex* _mutex, QObject *parent = nullptr) :
        QObject(parent), signal_in(_signal_in), mutex_in(_mutex)
    {
        mutex_out = new QMutex;
        signal_out = new QWaitCondition;
        decoder = new ldpc_decoder(signal_out, mutex_out);
    }
    ~llr_demapper() {}
    ldpc_decoder* decoder;

signals:
    void signal_noise_ratio(float _snr) T2B200_SIGNAL_BODY
    void soft_multiplexer_de_twist(int* _idx_plp_simd, l1_postsignalling _l1_post, int _len_out, int8_t* _out);   // -> ldpc_decoder::execute
    void stop_decoder() T2B200_SIGNAL_BODY
    void finished() T2B200_SIGNAL_BODY

public slots:
    void execute(int _ti_block_size, complex* _time_deint_cell, int _plp_id, l1_postsignalling _l1_post)
    {
        const l1_postsignalling_plp& plp = _l1_post.plp[_plp_id];
        const int fec_size = plp.plp_fec_type == FEC_FRAME_NORMAL ? FEC_SIZE_NORMAL : FEC_SIZE_SHORT;
        const int cells_per_fec = fec_size / (2 * (plp.plp_mod + 1));
        int32_t n_fec = _ti_block_size / cells_per_fec;
        llr.resize(static_cast<size_t>(n_fec) * fec_size);
        float snr = 0.0f;
        // the TI block is derotated in place, like the reference does (llr_demapper.cpp:555-557)
        t2b200_dropin::check(t2b200_demap(t2b200_dropin::context(), reinterpret_cast<float*>(_time_deint_cell), 1, &n_fec, plp.plp_mod,
                                          plp.plp_rotation, plp.plp_fec_type, plp.plp_cod, llr.data(), &snr, nullptr, nullptr),
                             "t2b200_demap");
        emit signal_noise_ratio(snr);
        for (int f = 0; f < n_fec; ++f) {
            std::vector<int8_t>& out = swap_buffer ? buffer_a : buffer_b;            // llr_demapper.cpp:753-762
            out.resize(static_cast<size_t>(SIZEOF_SIMD) * fec_size);
            memcpy(out.data() + static_cast<size_t>(blocks) * fec_size, llr.data() + static_cast<size_t>(f) * fec_size,
                   static_cast<size_t>(fec_size));
            idx_plp_simd[blocks] = _plp_id;
            if (++blocks == SIZEOF_SIMD) {
                blocks = 0;
                swap_buffer = !swap_buffer;
                mutex_out->lock();
                emit soft_multiplexer_de_twist(idx_plp_simd, _l1_post, fec_size * SIZEOF_SIMD, out.data());
                signal_out->wait(mutex_out);
                mutex_out->unlock();
            }
        }
    }
    void stop() {}

private:
    QWaitCondition* signal_in;
    QWaitCondition* signal_out;
    QMutex* mutex_in;
    QMutex* mutex_out;
    std::vector<int8_t> llr, buffer_a, buffer_b;
    bool swap_buffer = true;
    int blocks = 0;
    int idx_plp_simd[SIZEOF_SIMD] = {0};
};

#endif // LLR_DEMAPPER_H
