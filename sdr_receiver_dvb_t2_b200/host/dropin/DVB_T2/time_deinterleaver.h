// Drop-in for the reference's src/DVB_T2/time_deinterleaver.h: same class, constructor, start() and the two slots
// dvbt2_demodulator wires its `data` / `l1_dyn_execute` signals to (time_deinterleaver.h:27-46).  The cells of a T2 frame
// are collected per TI block exactly as the reference's streaming loop walks them (time_deinterleaver.cpp:268-376: the PLP
// whose dynamic start is 0 first, TI blocks of floor / ceil FEC blocks, the next PLP where the slice ends, whatever follows
// the last PLP filling further blocks that are dropped downstream or discarded at the next frame start); each complete
// block goes through t2b200_ti_deinterleave on the GPU -- cell de-interleaver permutation, row / column transpose and the
// unconditional move of Q one cell back -- and on to the demapper.
#ifndef TIME_DEINTERLEAVER_H
#define TIME_DEINTERLEAVER_H

#include <QObject>
#include <QThread>
#include <QMutex>
#include <vector>

#include "DSP/fast_fourier_transform.h"
#include "dvbt2_definition.h"
#include "llr_demapper.h"
#include "t2b200_dropin.h"

class time_deinterleaver : public QObject
{
    Q_OBJECT
public:
    explicit time_deinterleaver(QWaitCondition* _signal_in, QMutex* _mutex, QObject *parent = nullptr) :
        QObject(parent), signal_in(_signal_in), mutex_in(_mutex)
    {
        mutex_out = new QMutex;
        signal_out = new QWaitCondition;
        qam = new llr_demapper(signal_out, mutex_out);
    }
    ~time_deinterleaver() {}

    void start(dvbt2_parameters _dvbt2, l1_presignalling _l1_pre, l1_postsignalling _l1_post)
    {
        dvbt2 = _dvbt2;
        l1_pre = _l1_pre;
        l1_post = _l1_post;
        p2_start_idx_cell = L1_PRE_CELL + l1_pre.l1_post_size;                       // time_deinterleaver.cpp:44
        num_plp = l1_post.num_plp;
        cells_per_fec_block.assign(num_plp, 0);
        n_ti.assign(num_plp, 1);
        blocks.assign(num_plp, std::vector<int>());
        for (int i = 0; i < num_plp; ++i) {
            const l1_postsignalling_plp& p = l1_post.plp[i];
            const int fec_size = p.plp_fec_type == FEC_FRAME_NORMAL ? FEC_SIZE_NORMAL : FEC_SIZE_SHORT;
            cells_per_fec_block[i] = fec_size / (2 * (p.plp_mod + 1));
            n_ti[i] = p.time_il_type == 0 ? p.time_il_length : 1;                    // time_deinterleaver.cpp:117-129
            t2b200_dropin::check(t2b200_ti_configure(t2b200_dropin::context(), i, p.plp_fec_type, p.plp_mod, p.plp_num_blocks_max,
                                                     nullptr), "t2b200_ti_configure");
        }
        flag_start = true;
    }
    llr_demapper* qam;
    volatile int idx_show_plp = 0;

signals:
    void ti_block(int _ti_block_size, complex* _time_deint_cell, int _plp_id, l1_postsignalling _l1_post);   // -> llr_demapper::execute
    void replace_constelation(const int _len_data, complex* _data) T2B200_SIGNAL_BODY
    void stop_qam() T2B200_SIGNAL_BODY
    void finished() T2B200_SIGNAL_BODY

public slots:
    void l1_dyn_execute(l1_postsignalling _l1_post, int _len_in, complex* _ofdm_cell)
    {
        l1_post = _l1_post;                                                           // time_deinterleaver.cpp:268-286
        for (int i = 0; i < num_plp; ++i) {
            const int nb = l1_post.dyn.plp[i].num_blocks, base = nb / n_ti[i];
            blocks[i].assign(l1_post.plp[i].time_il_length, 0);
            for (int j = 0; j < l1_post.plp[i].time_il_length; ++j)
                blocks[i][j] = base + (j >= n_ti[i] - nb % n_ti[i] ? 1 : 0);
        }
        start_t2_frame = true;
        execute(_len_in, _ofdm_cell);
    }
    void execute(int _len_in, complex* _ofdm_cell)
    {
        if (!flag_start) return;
        int num_cells = _len_in;
        complex* ofdm_cell = _ofdm_cell;
        if (start_t2_frame) {                                                         // time_deinterleaver.cpp:300-315
            start_t2_frame = false;
            idx_cell = 0;
            num_cells = _len_in - p2_start_idx_cell;
            ofdm_cell = _ofdm_cell + p2_start_idx_cell;
            for (int i = 0; i < num_plp; ++i) if (l1_post.dyn.plp[i].start == 0) plp_id = i;
            idx_time_il = 0;
            pending.clear();
        }
        while (num_cells > 0) {
            const int ti_block_size = blocks[plp_id][idx_time_il] * cells_per_fec_block[plp_id];
            if (ti_block_size <= 0) return;                                           // a PLP without blocks in this frame
            const int take = std::min(num_cells, ti_block_size - static_cast<int>(pending.size()));
            pending.insert(pending.end(), ofdm_cell, ofdm_cell + take);
            ofdm_cell += take; num_cells -= take; idx_cell += take;
            if (static_cast<int>(pending.size()) < ti_block_size) break;
            std::vector<complex>& out = swap_buffers ? buffer_a : buffer_b;           // ping-pong, time_deinterleaver.cpp:346-353
            swap_buffers = !swap_buffers;
            out.resize(pending.size());
            int32_t n_fec = blocks[plp_id][idx_time_il];
            t2b200_dropin::check(t2b200_ti_deinterleave(t2b200_dropin::context(), plp_id, reinterpret_cast<const float*>(pending.data()), 1,
                                                        &n_fec, reinterpret_cast<float*>(out.data())), "t2b200_ti_deinterleave");
            pending.clear();
            mutex_out->lock();
            emit ti_block(ti_block_size, out.data(), plp_id, l1_post);
            signal_out->wait(mutex_out);
            mutex_out->unlock();
            if (++idx_time_il == l1_post.plp[plp_id].time_il_length) {               // time_deinterleaver.cpp:354-368
                idx_time_il = 0;
                for (int i = 0; i < num_plp; ++i)
                    if (i != plp_id && idx_cell == l1_post.dyn.plp[i].start) { plp_id = i; break; }
            }
        }
    }
    void stop() {}

private:
    QWaitCondition* signal_in;
    QWaitCondition* signal_out;
    QMutex* mutex_in;
    QMutex* mutex_out;
    dvbt2_parameters dvbt2;
    l1_presignalling l1_pre;
    l1_postsignalling l1_post;
    bool flag_start = false, start_t2_frame = true, swap_buffers = true;
    int p2_start_idx_cell = 0, num_plp = 0, plp_id = 0, idx_time_il = 0, idx_cell = 0;
    std::vector<int> cells_per_fec_block, n_ti;
    std::vector<std::vector<int>> blocks;
    std::vector<complex> pending, buffer_a, buffer_b;
};

#endif // TIME_DEINTERLEAVER_H
