// Minimal single-threaded stand-in for the GNU Radio scheduler: drives the blocks in the order of the
// shipped flowgraph (apps/vv009-4kshort.grc of the reference: bbheaderbch -> LDPC -> interleavermod ->
// framemapper -> pilotgen) through their gr::block interface (make / forecast / general_work), with the
// parameters of that flowgraph (4K, short FECFRAME, 256QAM rotated, CR 4/5, PP7, GI 1/32).  The LDPC stage,
// which the flowgraph takes from GNU Radio's gr-dtv, is this module's own ldpc_bb block here.
// Usage: gr_flowgraph_demo [n_t2_frames] [plain|link|link-lazy|auto] [tpb|seq] [c1|c3] [nocheck]   -- prints a checksum of the baseband; needs a CUDA device.
// c3 = BASELINE configuration 3 instead of the shipped flowgraph's: 32K extended, 256QAM rotated, CR 2/3, GI 1/128, PP7, 202 FECFRAMEs.
// With "link" adjacent blocks hand their items over in HBM (dvbt2ll/cuda_link.h); the buffers between the blocks
// are then kept at fixed addresses, as the scheduler's are.
// With "tpb" the blocks run the way GNU Radio's thread-per-block scheduler runs them: one thread per block, the edges
// are rings of three T2-frame slots at fixed addresses, a block works as soon as it has a frame of input and a free
// output slot -- so adjacent blocks overlap and a producer can be two frames ahead of its consumer.
#include <dvbt2ll/bbheaderbch_bb.h>
#include <dvbt2ll/cuda_link.h>
#include <dvbt2ll/ldpc_bb.h>
#include <dvbt2ll/framemapperfint_cc.h>
#include <dvbt2ll/interleavermod_bc.h>
#include <dvbt2ll/pilotgenp1insert_cc.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/dvbt2ll_cuda.h"

using namespace gr::dvbt2ll;

// one general_work() call producing `noutput` items into `out` (a buffer that keeps its address between calls)
template <class In, class Out>
static void run_block(gr::block &b, const std::vector<In> &in, size_t &in_pos, std::vector<Out> &out, int noutput)
{
  gr_vector_int need(1, 0);
  b.forecast(noutput, need);
  if (out.size() < (size_t)noutput) out.resize(noutput);
  gr_vector_int nin(1, (int)(in.size() - in_pos));
  gr_vector_const_void_star ins(1, (const void *)(in.data() + in_pos));
  gr_vector_void_star outs(1, (void *)out.data());
  const int produced = b.general_work(noutput, nin, ins, outs);
  if (produced != noutput) { fprintf(stderr, "block produced %d of %d items\n", produced, noutput); exit(1); }
  in_pos += b.last_consumed();
}

// one edge of the thread-per-block run: SLOTS buffers of one T2 frame's items each, reused round robin
struct Edge {
  enum { SLOTS = 3 };
  std::vector<unsigned char> mem;
  size_t slot_bytes;
  std::mutex m;
  std::condition_variable cv;
  long long produced, consumed;
  explicit Edge(size_t bytes) : mem(bytes * SLOTS + 64), slot_bytes(bytes), produced(0), consumed(0) {}
  void *slot(long long k) { return mem.data() + 64 + (size_t)(k % SLOTS) * slot_bytes; }
  void wait_space() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return produced - consumed < SLOTS; }); }
  void wait_item(long long k) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return produced > k; }); }
  void push() { { std::lock_guard<std::mutex> l(m); produced++; } cv.notify_all(); }
  void pop() { { std::lock_guard<std::mutex> l(m); consumed++; } cv.notify_all(); }
};

// the body of one block's thread: frame f of the input edge -> frame f of the output edge (the first block has no
// input edge: it reads `src_window` bytes of the TS at its stream position through a buffer at a fixed address)
static void block_thread(gr::block *b, Edge *in, int in_items, int src_window, Edge *out, int noutput, int nframes,
                         const unsigned char *ts, size_t ts_bytes, size_t *ts_pos)
{
  for (int f = 0; f < nframes; f++) {
    if (in) in->wait_item(f);
    out->wait_space();
    static std::vector<unsigned char> ts_in;        // source side of the first edge: fixed address, refilled per call
    if (!in) {
      const size_t n = std::min(ts_bytes - *ts_pos, (size_t)src_window);
      if (ts_in.size() < n) ts_in.resize(n);
      memcpy(ts_in.data(), ts + *ts_pos, n);
    }
    gr_vector_int nin(1, in ? in_items : (int)ts_in.size());
    gr_vector_const_void_star ins(1, in ? (const void *)in->slot(f) : (const void *)ts_in.data());
    gr_vector_void_star outs(1, out->slot(f));
    const int produced = b->general_work(noutput, nin, ins, outs);
    if (produced != noutput) { fprintf(stderr, "block produced %d of %d items\n", produced, noutput); exit(1); }
    if (in) in->pop();
    else *ts_pos += b->last_consumed();
    out->push();
  }
}

int main(int argc, char **argv)
{
  const int nframes = argc > 1 ? atoi(argv[1]) : 2;
  const bool tpb = argc > 3 && !strcmp(argv[3], "tpb");
  const bool lazy = argc > 2 && !strcmp(argv[2], "link-lazy");
  if (argc > 2 && !strcmp(argv[2], "auto")) dvbt2ll_set_auto_link(1);       // no link() calls: the blocks find each other's outputs
  const bool linked = lazy || (argc > 2 && !strcmp(argv[2], "link"));
  const bool c3 = argc > 4 && !strcmp(argv[4], "c3");
  const bool check = !(argc > 5 && !strcmp(argv[5], "nocheck"));      // timing runs skip the sink's checksum arithmetic
  const int fecblocks = c3 ? 202 : 8;
  const dvbt2_framesize_t fs = c3 ? FECFRAME_NORMAL : FECFRAME_SHORT;
  const dvbt2_code_rate_t cr = c3 ? C2_3 : C4_5;
  const dvbt2_extended_carrier_t cm = c3 ? CARRIERS_EXTENDED : CARRIERS_NORMAL;
  const dvbt2_fftsize_t fft = c3 ? FFTSIZE_32K : FFTSIZE_4K;
  const dvbt2_guardinterval_t gi = c3 ? GI_1_128 : GI_1_32;
  const int nds = c3 ? 59 : 3;
  // items per FECFRAME / T2 frame of the two configurations (the blocks' own output multiples)
  const int nbch = c3 ? 43200 : 12600, nldpc = c3 ? 64800 : 16200, ncell = c3 ? 8100 : 2025;
  const int mapped_items = c3 ? 1639268 : 18866, frame_samples = c3 ? 1983488 : 31616, ts_per_frame = c3 ? 1084740 : 12352;
  bbheaderbch_bb::sptr bb = bbheaderbch_bb::make(fs, cr, INPUTMODE_NORMAL, INBAND_OFF, fecblocks, 4000000);
  interleavermod_bc::sptr im = interleavermod_bc::make(fs, cr, MOD_256QAM, ROTATION_ON);
  framemapperfint_cc::sptr fm = framemapperfint_cc::make(fs, cr, MOD_256QAM, ROTATION_ON, fecblocks, 3, cm,
      fft, gi, L1_MOD_64QAM, PILOT_PP7, 2, nds, PAPR_OFF, VERSION_111, PREAMBLE_T2_SISO, INPUTMODE_NORMAL, RESERVED_OFF,
      L1_SCRAMBLED_OFF, INBAND_OFF);
  pilotgenp1insert_cc::sptr pg = pilotgenp1insert_cc::make(cm, fft, PILOT_PP7, gi, nds, PAPR_OFF, VERSION_111,
      PREAMBLE_T2_SISO, MISO_TX1, EQUALIZATION_OFF, BANDWIDTH_8_0_MHZ, c3 ? 32768 : 4096);
  ldpc_bb::sptr ldpc = ldpc_bb::make(fs, cr);
  if (linked) {
    if (!link(bb.get(), ldpc.get(), lazy) || !link(ldpc.get(), im.get(), lazy) || !link(im.get(), fm.get(), lazy) || !link(fm.get(), pg.get(), lazy)) {
      fprintf(stderr, "link failed: %s\n", dvbt2ll_last_error());
      return 1;
    }
  }

  // synthetic TS (xorshift32, sync byte every 188 bytes), as in bench.py / the tests
  std::vector<unsigned char> ts((size_t)nframes * ts_per_frame + 1000);
  uint32_t x = 0x12345678u;
  for (size_t i = 0; i < ts.size(); i++) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; ts[i] = (i % 188 == 0) ? 0x47 : (unsigned char)x; }

  size_t ts_pos = 0;
  double acc = 0.0;
  if (tpb) {
    Edge e_bch((size_t)fecblocks * nbch), e_fec((size_t)fecblocks * nldpc), e_cells((size_t)fecblocks * ncell * sizeof(gr_complex)),
         e_mapped((size_t)mapped_items * sizeof(gr_complex)), e_samples((size_t)frame_samples * sizeof(gr_complex));
    std::thread t0(block_thread, bb.get(), (Edge *)0, 0, ts_per_frame + 1000, &e_bch, fecblocks * nbch, nframes, ts.data(), ts.size(), &ts_pos);
    std::thread t1(block_thread, ldpc.get(), &e_bch, fecblocks * nbch, 1, &e_fec, fecblocks * nldpc, nframes, (const unsigned char *)0, (size_t)0, (size_t *)0);
    std::thread t2(block_thread, im.get(), &e_fec, fecblocks * nldpc, 1, &e_cells, fecblocks * ncell, nframes, (const unsigned char *)0, (size_t)0, (size_t *)0);
    std::thread t3(block_thread, fm.get(), &e_cells, fecblocks * ncell, 8, &e_mapped, mapped_items, nframes, (const unsigned char *)0, (size_t)0, (size_t *)0);
    std::thread t4(block_thread, pg.get(), &e_mapped, mapped_items, 8, &e_samples, frame_samples, nframes, (const unsigned char *)0, (size_t)0, (size_t *)0);
    std::chrono::steady_clock::time_point t_first;
    for (int f = 0; f < nframes; f++) {           // the sink
      e_samples.wait_item(f);
      if (f == 0) t_first = std::chrono::steady_clock::now();
      const gr_complex *sm = (const gr_complex *)e_samples.slot(f);
      for (int i = 0; check && i < frame_samples; i++) acc += (f + 1) * (double)std::abs(sm[i]);       // frame-weighted: order matters
      e_samples.pop();
    }
    t0.join(); t1.join(); t2.join(); t3.join(); t4.join();
    if (nframes > 1)
    {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_first).count() / (nframes - 1);
      printf("steady state: %.3f ms per T2 frame (thread per block), %.1f Msamples/s\n", ms, frame_samples / ms * 1e-3);
    }
    printf("T2 frame %d: %d samples, TS consumed so far %zu bytes\n", nframes - 1, frame_samples, ts_pos);
    printf("sum (f+1)|x| = %.6f, kernel launches = %lld\n", acc, dvbt2ll_kernel_launches());
    bb.reset(); ldpc.reset(); im.reset(); fm.reset(); pg.reset();      // before the edges' buffers go
    return 0;
  }
  std::vector<unsigned char> bch, fec;
  std::vector<gr_complex> cells, mapped, samples;
  // the source side of the first edge: a buffer at a fixed address that is refilled with the next stretch of the TS
  // before every call, as a source block writing into the scheduler's ring buffer does
  std::vector<unsigned char> ts_in((size_t)ts_per_frame + 1000);
  std::chrono::steady_clock::time_point t_first;
  for (int f = 0; f < nframes; f++) {
    if (f == 1) t_first = std::chrono::steady_clock::now();
    size_t p = 0;
    memcpy(ts_in.data(), ts.data() + ts_pos, ts_in.size());
    run_block(*bb, ts_in, p, bch, fecblocks * nbch);
    ts_pos += p;
    p = 0;
    run_block(*ldpc, bch, p, fec, fecblocks * nldpc);
    p = 0;
    run_block(*im, fec, p, cells, fecblocks * ncell);
    p = 0;
    run_block(*fm, cells, p, mapped, mapped_items);
    p = 0;
    run_block(*pg, mapped, p, samples, frame_samples);
    for (size_t i = 0; check && i < samples.size(); i++) acc += (f + 1) * (double)std::abs(samples[i]);
    printf("T2 frame %d: %zu samples, TS consumed so far %zu bytes\n", f, samples.size(), ts_pos);
  }
  if (nframes > 1)
  {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_first).count() / (nframes - 1);
    printf("steady state: %.3f ms per T2 frame (one thread), %.1f Msamples/s\n", ms, frame_samples / ms * 1e-3);
  }
  printf("sum (f+1)|x| = %.6f, kernel launches = %lld\n", acc, dvbt2ll_kernel_launches());
  // the blocks (and with them the registrations of these buffers) go before the buffers do
  bb.reset(); ldpc.reset(); im.reset(); fm.reset(); pg.reset();
  return 0;
}
