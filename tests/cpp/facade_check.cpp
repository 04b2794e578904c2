// Drives the host-side mirror of the reference's stage API (sdr_receiver_dvb_t2_b200/host/t2b200_stages.hpp) the way
// dvbt2_demodulator::symbol_acquisition does (dvbt2_demodulator.cpp:332-385): per OFDM symbol memcpy -> fft->execute()
// -> p2 / data symbol execute -> deinterleaver l1_dyn_execute / execute -> ... -> BBFRAME callback.
// Usage: facade_check <in.bin> <out.bin>   (file layout written by tests/test_facade.py)
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../sdr_receiver_dvb_t2_b200/host/t2b200_stages.hpp"

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (fread(v.data(), sizeof(T), n, f) != n) { perror("read"); exit(2); } return v; }

int main(int argc, char** argv)
{
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  auto h = rd<int32_t>(f, 20);
  t2b200::symbol_mode m{h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]};
  t2b200::plp_config plp{h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17]};
  const int l1_post_size = h[18], num_blocks = h[19];
  const int nd = m.len_frame - m.n_p2 - m.l_fc;
  auto amps = rd<float>(f, 3);                                   // amp_sp, amp_cp, amp_p2
  auto data_map = rd<int32_t>(f, (size_t)nd * m.k_total); auto data_ref = rd<float>(f, (size_t)nd * m.k_total);
  auto p2_map = rd<int32_t>(f, m.k_total); auto p2_ref = rd<float>(f, m.k_total);
  auto he_d = rd<int32_t>(f, m.c_data), ho_d = rd<int32_t>(f, m.c_data), he_p = rd<int32_t>(f, m.c_p2), ho_p = rd<int32_t>(f, m.c_p2);
  auto time = rd<t2b200::complex>(f, (size_t)m.len_frame * m.fft_size);
  fclose(f);

  std::vector<int*> map_rows(nd); std::vector<float*> ref_rows(nd);
  for (int s = 0; s < nd; ++s) { map_rows[s] = data_map.data() + (size_t)s * m.k_total; ref_rows[s] = data_ref.data() + (size_t)s * m.k_total; }

  FILE* out = fopen(argv[2], "wb");
  FILE* ts = argc > 3 ? fopen(argv[3], "wb") : nullptr;       // optional: the datagrams of the bb_de_header mirror, back to back
  int n_frames = 0, n_datagrams = 0;
  try {
    t2b200::context ctx(0);
    t2b200::fast_fourier_transform fft(ctx);
    t2b200::p2_symbol_equalizer p2(ctx);
    t2b200::data_symbol data(ctx);
    t2b200::bb_de_header deheader(ctx, [&](const uint8_t* d, int len) { if (ts) fwrite(d, 1, len, ts); ++n_datagrams; });
    deheader.set_out(plp.id);
    t2b200::fec_chain fec(ctx, [&](int id, int len, uint8_t* bits) {          // emit bit_descramble -> deheader->execute
      fwrite(bits, 1, len, out); ++n_frames;
      deheader.execute(id, len, bits);
    });
    t2b200::complex* in_fft = fft.init(m.fft_size);
    p2.init(m, p2_map.data(), p2_ref.data(), he_p.data(), ho_p.data(), amps[2]);
    data.init(m, map_rows.data(), ref_rows.data(), he_d.data(), ho_d.data(), amps[0], amps[1]);
    fec.start(plp, l1_post_size);
    float sro, ph;
    for (int l = 0; l < m.len_frame; ++l) {
      std::memcpy(in_fft, time.data() + (size_t)l * m.fft_size, sizeof(t2b200::complex) * m.fft_size);
      t2b200::complex* cell = fft.execute();
      if (l == 0) fec.l1_dyn_execute(num_blocks, m.c_p2, p2.execute(0, cell, sro, ph));
      else fec.execute(m.c_data, data.execute(l, cell, sro, ph));
    }
  } catch (const std::exception& e) { std::fprintf(stderr, "facade_check: %s\n", e.what()); fclose(out); if (ts) fclose(ts); return 1; }
  fclose(out);
  if (ts) fclose(ts);
  // the front-end mirror: one chunk of a constant input -> as many decimator outputs as inputs at resample = 0.5, the DC
  // component removed by less than a thousandth (ratio 1e-6), the in-band gain of the half-band filter close to one
  int fe_out = 0; float fe_level = 0.f;
  try {
    t2b200::context ctx(0);
    t2b200::frontend fe(ctx, 8192);
    std::vector<int16_t> i16(4096, 4096), q16(4096, -2048);
    std::vector<t2b200::complex> dec(4200);
    float th1 = 0.f, th2 = 0.f, th3 = 0.f;
    fe_out = fe.execute(4096, i16.data(), q16.data(), 1, 1.0f / (1 << 14), 0.f, 1.f, 0.f, 0.f, 0.5, dec.data(), (int)dec.size(), th1, th2, th3);
    fe_level = dec[3000].real();
  } catch (const std::exception& e) { std::fprintf(stderr, "facade_check (front-end): %s\n", e.what()); return 1; }
  std::printf("bbframes %d datagrams %d frontend %d level %.3f\n", n_frames, n_datagrams, fe_out, fe_level);
  return 0;
}
