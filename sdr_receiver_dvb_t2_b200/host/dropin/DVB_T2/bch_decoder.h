// Drop-in for the reference's src/DVB_T2/bch_decoder.h: same class, constructor and execute() signature (bch_decoder.h:28-42).
// execute() = t2b200_bch_descramble on the GPU: keep the first K_bch bits of each of the 32 LDPC information words and XOR
// the BB-scrambler PRBS (bch_decoder.cpp:63-164; like the reference, no BCH error correction), then one
// bit_descramble per BBFRAME into the reference's own bb_de_header.
#ifndef BCH_DECODER_H
#define BCH_DECODER_H

#include <QObject>
#include <QThread>
#include <QMutex>
#include <QWaitCondition>
#include <vector>

#include "dvbt2_definition.h"
#include "bb_de_header.h"
#include "t2b200_dropin.h"

class bch_decoder : public QObject
{
    Q_OBJECT
public:
    explicit bch_decoder(QWaitCondition* _signal_in, QMutex* _mutex_in, QObject *parent = nullptr) :
        QObject(parent), signal_in(_signal_in), mutex_in(_mutex_in)
    {
        mutex_out = new QMutex;
        signal_out = new QWaitCondition;
        deheader = new bb_de_header(signal_out, mutex_out);
    }
    ~bch_decoder() {}
    bb_de_header* deheader;

signals:
    void bit_descramble(int _plp_id, l1_postsignalling _l1_post, int _lenout, uint8_t* out);   // -> bb_de_header::execute
    void check(int _len, uint8_t* out) T2B200_SIGNAL_BODY
    void stop_deheader() T2B200_SIGNAL_BODY
    void finished() T2B200_SIGNAL_BODY

public slots:
    void execute(int* _idx_plp_simd, l1_postsignalling _l1_post, int _len_in, uint8_t* _in)
    {
        const l1_postsignalling_plp& plp = _l1_post.plp[_idx_plp_simd[0]];           // bch_decoder.cpp:72-75
        const int code = t2b200_ldpc_code_id(plp.plp_fec_type, plp.plp_cod);
        const int k_ldpc = t2b200_ldpc_k(code), k_bch = t2b200_ldpc_k_bch(code);
        const int n = _len_in / k_ldpc;
        std::vector<uint8_t>& out = swap_buffer ? buffer_a : buffer_b;               // ping-pong, bch_decoder.cpp:143-160
        swap_buffer = !swap_buffer;
        out.resize(static_cast<size_t>(n) * k_bch);
        t2b200_dropin::check(t2b200_bch_descramble(t2b200_dropin::context(), code, _in, n, out.data()), "t2b200_bch_descramble");
        for (int i = 0; i < n; ++i) {
            mutex_out->lock();
            emit bit_descramble(_idx_plp_simd[i], _l1_post, k_bch, out.data() + static_cast<size_t>(i) * k_bch);
            signal_out->wait(mutex_out);
            mutex_out->unlock();
        }
    }
    void stop() {}

private:
    QWaitCondition* signal_in;
    QWaitCondition* signal_out;
    QMutex* mutex_in;
    QMutex* mutex_out;
    std::vector<uint8_t> buffer_a, buffer_b;
    bool swap_buffer = true;
};

#endif // BCH_DECODER_H
