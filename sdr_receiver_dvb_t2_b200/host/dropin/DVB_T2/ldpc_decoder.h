// Drop-in for the reference's src/DVB_T2/ldpc_decoder.h: same class, constructor and execute() signature
// (ldpc_decoder.h:75-91).  execute() = t2b200_ldpc_decode on the GPU with the reference's batch semantics: the 32
// FECFRAMEs of a call are decoded in lock step, at most 25 trials, and a batch that does not converge is dropped with
// the reference's message (ldpc_decoder.cpp:157-301); output one byte per bit, K_ldpc bits per frame.
#ifndef LDPC_DECODER_H
#define LDPC_DECODER_H

#include <QObject>
#include <QThread>
#include <QMutex>
#include <QWaitCondition>
#include <cstdio>
#include <vector>

#include "dvbt2_definition.h"
#include "bch_decoder.h"
#include "t2b200_dropin.h"

#define SIZEOF_SIMD 32
#define TRIALS 25

class ldpc_decoder : public QObject
{
    Q_OBJECT
public:
    explicit ldpc_decoder(QWaitCondition* _signal_in, QMutex* _mutex_in, QObject *parent = nullptr) :
        QObject(parent), signal_in(_signal_in), mutex_in(_mutex_in)
    {
        mutex_out = new QMutex;
        signal_out = new QWaitCondition;
        decoder = new bch_decoder(signal_out, mutex_out);
    }
    ~ldpc_decoder() {}
    bch_decoder* decoder;

signals:
    void bit_bch(int* _idx_plp_simd, l1_postsignalling _l1_post, int _lenout, uint8_t* out);      // -> bch_decoder::execute
    void check(int _lenout, uint8_t* out) T2B200_SIGNAL_BODY
    void stop_decoder() T2B200_SIGNAL_BODY
    void finished() T2B200_SIGNAL_BODY

public slots:
    void execute(int* _idx_plp_simd, l1_postsignalling _l1_post, int _len_in, int8_t *_in)
    {
        const l1_postsignalling_plp& plp = _l1_post.plp[_idx_plp_simd[0]];           // ldpc_decoder.cpp:173-174
        const int code = t2b200_ldpc_code_id(plp.plp_fec_type, plp.plp_cod);
        const int n_ldpc = t2b200_ldpc_n(code), k_ldpc = t2b200_ldpc_k(code);
        const int n = _len_in / n_ldpc;
        std::vector<uint8_t>& out = swap_buffer ? buffer_a : buffer_b;               // ldpc_decoder.cpp:283-298
        out.resize(static_cast<size_t>(n) * k_ldpc);
        trials.resize(static_cast<size_t>(n));
        t2b200_dropin::check(t2b200_ldpc_decode(t2b200_dropin::context(), code, _in, n, out.data(), trials.data(), nullptr, nullptr,
                                                TRIALS, T2B200_LDPC_GROUP32), "t2b200_ldpc_decode");
        if (trials[0] < 0) {                                                         // ldpc_decoder.cpp:264-268
            fprintf(stderr, "LDPC decoder could not recover the codeword! %d\n", trials[0]);
            return;
        }
        swap_buffer = !swap_buffer;
        mutex_out->lock();
        emit bit_bch(_idx_plp_simd, _l1_post, n * k_ldpc, out.data());
        signal_out->wait(mutex_out);
        mutex_out->unlock();
    }
    void stop() {}

private:
    QWaitCondition* signal_in;
    QWaitCondition* signal_out;
    QMutex* mutex_in;
    QMutex* mutex_out;
    std::vector<uint8_t> buffer_a, buffer_b;
    std::vector<int32_t> trials;
    bool swap_buffer = true;
};

#endif // LDPC_DECODER_H
