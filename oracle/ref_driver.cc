/*
 * TEST INFRASTRUCTURE ONLY -- C entry points around the UNMODIFIED reference
 * blocks (compiled in place from /root/reference/lib by oracle/Makefile into
 * oracle/_ref/libdvbt2ll_ref.so).  Used by tests/, tools/gen_tables.py, the
 * golden-vector generator and bench.py's cpu_baseline / --impl reference legs.
 * The product library never links or loads this.
 *
 * The reference is only correct when general_work() is asked for exactly ONE
 * frame of output per call ("one frame per call"):
 * framemapperfint_cc_impl.cc:1967/:2147 and pilotgenp1insert_cc_impl.cc:2801/:2903
 * consume one frame whatever noutput_items says, and interleavermod_bc_impl.h:40-41
 * has a 64800-byte staging array.  ref_block_work() therefore loops frame by frame.
 */
#include <gnuradio/block.h>
#include <gnuradio/fft/fft.h>
#include <volk/volk.h>
#include <sstream>
#include <chrono>
#define private public
#define protected public
#include "bbheaderbch_bb_impl.h"
#include "interleavermod_bc_impl.h"
#include "framemapperfint_cc_impl.h"
#include "pilotgenp1insert_cc_impl.h"
#undef private
#undef protected

#include <chrono>
#include <cstdint>

using namespace gr::dvbt2ll;

extern "C" {

int oracle_shim_quiet = 0;
int oracle_shim_fft_fast = 0;
double oracle_shim_fft_seconds = 0.0;

void ref_set_quiet(int q) { oracle_shim_quiet = q; }
void ref_set_fft_fast(int f) { oracle_shim_fft_fast = f; }
double ref_fft_seconds(int reset) { const double v = oracle_shim_fft_seconds; if (reset) oracle_shim_fft_seconds = 0.0; return v; }

enum { REF_BBHEADERBCH = 1, REF_INTERLEAVERMOD = 2, REF_FRAMEMAPPER = 3, REF_PILOTGEN = 4 };

struct ref_handle {
  int kind;
  gr::block *blk;
  int in_item, out_item;
  boost::shared_ptr<gr::block> keep;
};

static ref_handle *wrap(int kind, boost::shared_ptr<gr::block> p, int in_item, int out_item)
{
  ref_handle *h = new ref_handle();
  h->kind = kind; h->blk = p.get(); h->keep = p; h->in_item = in_item; h->out_item = out_item;
  return h;
}

void *ref_bbheaderbch_new(int framesize, int rate, int mode, int inband, int fecblocks, int tsrate)
{
  return wrap(REF_BBHEADERBCH,
              bbheaderbch_bb::make((dvbt2_framesize_t)framesize, (dvbt2_code_rate_t)rate,
                                   (dvbt2_inputmode_t)mode, (dvbt2_inband_t)inband, fecblocks, tsrate),
              1, 1);
}

void *ref_interleavermod_new(int framesize, int rate, int constellation, int rotation)
{
  return wrap(REF_INTERLEAVERMOD,
              interleavermod_bc::make((dvbt2_framesize_t)framesize, (dvbt2_code_rate_t)rate,
                                      (dvbt2_constellation_t)constellation, (dvbt2_rotation_t)rotation),
              1, 8);
}

void *ref_framemapper_new(int framesize, int rate, int constellation, int rotation, int fecblocks,
                          int tiblocks, int carriermode, int fftsize, int guardinterval,
                          int l1constellation, int pilotpattern, int t2frames, int numdatasyms,
                          int paprmode, int version, int preamble, int inputmode,
                          int reservedbiasbits, int l1scrambled, int inband)
{
  boost::shared_ptr<framemapperfint_cc> blk = framemapperfint_cc::make((dvbt2_framesize_t)framesize, (dvbt2_code_rate_t)rate,
                  (dvbt2_constellation_t)constellation, (dvbt2_rotation_t)rotation, fecblocks, tiblocks,
                  (dvbt2_extended_carrier_t)carriermode, (dvbt2_fftsize_t)fftsize,
                  (dvbt2_guardinterval_t)guardinterval, (dvbt2_l1constellation_t)l1constellation,
                  (dvbt2_pilotpattern_t)pilotpattern, t2frames, numdatasyms, (dvbt2_papr_t)paprmode,
                  (dvbt2_version_t)version, (dvbt2_preamble_t)preamble, (dvbt2_inputmode_t)inputmode,
                  (dvbt2_reservedbiasbits_t)reservedbiasbits, (dvbt2_l1scrambled_t)l1scrambled,
                  (dvbt2_inband_t)inband);
  /* The reference never initialises L1Post::plp_id_dynamic (framemapperfint_cc_impl.cc:240 assigns
   * plp_id a second time instead), so the dynamic PLP_ID field of L1-post is whatever the heap held.
   * The only PLP has PLP_ID 0; pin the field to that so the oracle is deterministic. */
  dynamic_cast<framemapperfint_cc_impl *>(blk.get())->L1_Signalling[0].l1post_data.plp_id_dynamic = 0;
  return wrap(REF_FRAMEMAPPER, blk, 8, 8);
}

void *ref_pilotgen_new(int carriermode, int fftsize, int pilotpattern, int guardinterval,
                       int numdatasyms, int paprmode, int version, int preamble, int misogroup,
                       int equalization, int bandwidth, int vlength)
{
  return wrap(REF_PILOTGEN,
              pilotgenp1insert_cc::make((dvbt2_extended_carrier_t)carriermode, (dvbt2_fftsize_t)fftsize,
                  (dvbt2_pilotpattern_t)pilotpattern, (dvbt2_guardinterval_t)guardinterval, numdatasyms,
                  (dvbt2_papr_t)paprmode, (dvbt2_version_t)version, (dvbt2_preamble_t)preamble,
                  (dvbt2_misogroup_t)misogroup, (dvbt2_equalization_t)equalization,
                  (dvbt2_bandwidth_t)bandwidth, vlength),
              8, 8);
}

void ref_block_free(void *hv) { delete (ref_handle *)hv; }

int ref_block_output_multiple(void *hv) { return ((ref_handle *)hv)->blk->output_multiple(); }

int ref_block_warnings(void *hv) { return ((ref_handle *)hv)->blk->shim_warnings(); }

int ref_block_forecast(void *hv, int noutput)
{
  gr_vector_int req(1, 0);
  ((ref_handle *)hv)->blk->forecast(noutput, req);
  return req[0];
}

/* Runs nframes = noutput / output_multiple frames, ONE general_work() call per frame.
 * Returns items produced; *consumed = total input items consumed.
 * The caller must supply enough input (use ref_block_forecast per frame; the BB header
 * block consumes a data-dependent amount only in HIEFF mode). */
int ref_block_work(void *hv, int noutput, const void *in, int ninput, void *out, int *consumed)
{
  ref_handle *h = (ref_handle *)hv;
  const int om = h->blk->output_multiple();
  const char *ip = (const char *)in;
  char *op = (char *)out;
  int used = 0, produced = 0;
  for (int done = 0; done + om <= noutput; done += om) {
    gr_vector_int ninp(1, ninput - used);
    gr_vector_const_void_star ins(1, (const void *)(ip + (size_t)used * h->in_item));
    gr_vector_void_star outs(1, (void *)(op + (size_t)produced * h->out_item));
    int r = h->blk->general_work(om, ninp, ins, outs);
    if (r < 0) return r;
    produced += r;
    used += h->blk->last_consumed();
  }
  if (consumed) *consumed = used;
  return produced;
}

/* The reference's own (dead-code) LDPC encoder, lib/bbheaderbch_bb_impl.cc:625-646.
 * `frame` holds nbch info bits followed by room for the parity bits (frame_size total). */
int ref_ldpc_calculate(void *hv, unsigned char *frame)
{
  ref_handle *h = (ref_handle *)hv;
  if (h->kind != REF_BBHEADERBCH) return -1;
  (dynamic_cast<bbheaderbch_bb_impl *>(h->blk))->ldpc_calculate(frame);
  return 0;
}

/* ---- integer scalars / arrays of the reference objects, by name (for table pinning) ---- */

#define GETI(field) if (!strcmp(name, #field)) return (long)b->field;

long ref_get_int(void *hv, const char *name)
{
  ref_handle *h = (ref_handle *)hv;
  switch (h->kind) {
    case REF_BBHEADERBCH: {
      bbheaderbch_bb_impl *b = dynamic_cast<bbheaderbch_bb_impl *>(h->blk);
      GETI(kbch) GETI(nbch) GETI(q_val) GETI(bch_code) GETI(num_parity_bits) GETI(frame_size)
      GETI(count) GETI(crc) GETI(fec_block) GETI(extra)
      break; }
    case REF_INTERLEAVERMOD: {
      interleavermod_bc_impl *b = dynamic_cast<interleavermod_bc_impl *>(h->blk);
      GETI(cell_size) GETI(nbch) GETI(q_val) GETI(mod) GETI(frame_size) GETI(cyclic_delay)
      break; }
    case REF_FRAMEMAPPER: {
      framemapperfint_cc_impl *b = dynamic_cast<framemapperfint_cc_impl *>(h->blk);
      GETI(cell_size) GETI(stream_items) GETI(mapped_items) GETI(eta_mod) GETI(t2_frames)
      GETI(t2_frame_num) GETI(N_P2) GETI(C_P2) GETI(N_FC) GETI(C_FC) GETI(C_DATA) GETI(N_post)
      GETI(N_punc) GETI(num_data_symbols) GETI(pn_degree) GETI(ti_blocks) GETI(fec_blocks)
      GETI(FECBlocksPerSmallTIBlock) GETI(FECBlocksPerBigTIBlock) GETI(numBigTIBlocks)
      GETI(numSmallTIBlocks)
      break; }
    case REF_PILOTGEN: {
      pilotgenp1insert_cc_impl *b = dynamic_cast<pilotgenp1insert_cc_impl *>(h->blk);
      GETI(active_items) GETI(num_symbols) GETI(left_nulls) GETI(right_nulls) GETI(guard_interval)
      GETI(N_P2) GETI(C_P2) GETI(N_FC) GETI(C_FC) GETI(C_DATA) GETI(K_EXT) GETI(C_PS) GETI(K_OFFSET)
      GETI(dx) GETI(dy) GETI(miso) GETI(ofdm_fft_size)
      break; }
  }
  return -999999999L;
}

float ref_get_float(void *hv, const char *name)
{
  ref_handle *h = (ref_handle *)hv;
  if (h->kind == REF_PILOTGEN) {
    pilotgenp1insert_cc_impl *b = dynamic_cast<pilotgenp1insert_cc_impl *>(h->blk);
    if (!strcmp(name, "normalization")) return b->normalization;
  }
  return NAN;
}

#define GETA(field, n) if (!strcmp(name, #field)) { *len = (n); return (const int *)b->field; }

/* int arrays; for the pilot block "data_carrier_map:<symbol>" first runs init_pilots(symbol). */
const int *ref_get_int_array(void *hv, const char *name, int *len)
{
  ref_handle *h = (ref_handle *)hv;
  *len = 0;
  if (h->kind == REF_FRAMEMAPPER) {
    framemapperfint_cc_impl *b = dynamic_cast<framemapperfint_cc_impl *>(h->blk);
    GETA(permutations, b->cell_size)
    GETA(Heven, b->C_DATA) GETA(Hodd, b->C_DATA)
    GETA(HevenP2, b->C_P2) GETA(HoddP2, b->C_P2)
    GETA(HevenFC, b->N_FC) GETA(HoddFC, b->N_FC)
  }
  else if (h->kind == REF_PILOTGEN) {
    pilotgenp1insert_cc_impl *b = dynamic_cast<pilotgenp1insert_cc_impl *>(h->blk);
    GETA(prbs, MAX_CARRIERS) GETA(pn_sequence, CHIPS)
    GETA(p2_carrier_map, b->C_PS) GETA(fc_carrier_map, b->C_PS)
    if (!strncmp(name, "data_carrier_map:", 17)) {
      b->init_pilots(atoi(name + 17));
      *len = b->C_PS;
      return b->data_carrier_map;
    }
  }
  return 0;
}

#define GETC(field, n) if (!strcmp(name, #field)) { *len = (n); return (const float *)b->field; }

/* complex<float> arrays, returned as interleaved floats; *len counts complex items. */
const float *ref_get_complex_array(void *hv, const char *name, int *len)
{
  ref_handle *h = (ref_handle *)hv;
  *len = 0;
  if (h->kind == REF_INTERLEAVERMOD) {
    interleavermod_bc_impl *b = dynamic_cast<interleavermod_bc_impl *>(h->blk);
    GETC(m_qpsk, 4) GETC(m_16qam, 16) GETC(m_64qam, 64) GETC(m_256qam, 256)
  }
  else if (h->kind == REF_FRAMEMAPPER) {
    framemapperfint_cc_impl *b = dynamic_cast<framemapperfint_cc_impl *>(h->blk);
    GETC(l1pre_cache, 1840)
    GETC(dummy_randomize, b->mapped_items - b->stream_items - 1840 - (b->N_post / b->eta_mod) - (b->N_FC - b->C_FC))
  }
  else if (h->kind == REF_PILOTGEN) {
    pilotgenp1insert_cc_impl *b = dynamic_cast<pilotgenp1insert_cc_impl *>(h->blk);
    GETC(p1_time, 1024) GETC(p1_timeshft, 1024) GETC(p1_freq, 1024)
    GETC(inverse_sinc, b->ofdm_fft_size)
    GETC(p2_bpsk, 2) GETC(sp_bpsk, 2) GETC(cp_bpsk, 2)
  }
  return 0;
}

/* L1-post cells for a given frame index (reference recomputes them every frame,
 * framemapperfint_cc_impl.cc:2033). out must hold N_post/eta_mod complex items. */
int ref_framemapper_l1post(void *hv, int t2_frame_num, float *out)
{
  ref_handle *h = (ref_handle *)hv;
  if (h->kind != REF_FRAMEMAPPER) return -1;
  framemapperfint_cc_impl *b = dynamic_cast<framemapperfint_cc_impl *>(h->blk);
  b->add_l1post((gr_complex *)out, t2_frame_num);
  return b->N_post / b->eta_mod;
}

/* ---- the standard's constant tables as transcribed in the reference (EN 302 755) ----
 * Used ONLY by tools/gen_tables.py to emit the product's own table file in its own format. */

struct tab_desc { const char *name; const int *data; int rows; int cols; };

#define T2(cls, t, r, c) { #t, &cls::t[0][0], r, c }
#define T1(cls, t, n) { #t, &cls::t[0], 1, n }

static const tab_desc g_tabs[] = {
  T2(bbheaderbch_bb_impl, ldpc_tab_1_2N, 90, 9), T2(bbheaderbch_bb_impl, ldpc_tab_3_5N, 108, 13),
  T2(bbheaderbch_bb_impl, ldpc_tab_2_3N_DVBT2, 120, 14), T2(bbheaderbch_bb_impl, ldpc_tab_3_4N, 135, 13),
  T2(bbheaderbch_bb_impl, ldpc_tab_4_5N, 144, 12), T2(bbheaderbch_bb_impl, ldpc_tab_5_6N, 150, 14),
  T2(bbheaderbch_bb_impl, ldpc_tab_1_3S, 15, 13), T2(bbheaderbch_bb_impl, ldpc_tab_2_5S, 18, 13),
  T2(bbheaderbch_bb_impl, ldpc_tab_1_2S, 20, 9), T2(bbheaderbch_bb_impl, ldpc_tab_3_5S_DVBT2, 27, 13),
  T2(bbheaderbch_bb_impl, ldpc_tab_2_3S, 30, 14), T2(bbheaderbch_bb_impl, ldpc_tab_3_4S, 33, 13),
  T2(bbheaderbch_bb_impl, ldpc_tab_4_5S, 35, 4), T2(bbheaderbch_bb_impl, ldpc_tab_5_6S, 37, 14),
  { "ldpc_tab_1_4S_L1", &framemapperfint_cc_impl::ldpc_tab_1_4S[0][0], 9, 13 },
  { "ldpc_tab_1_2S_L1", &framemapperfint_cc_impl::ldpc_tab_1_2S[0][0], 20, 9 },
  T1(framemapperfint_cc_impl, pre_puncture, 36),
  T1(framemapperfint_cc_impl, post_padding_bqpsk, 20), T1(framemapperfint_cc_impl, post_padding_16qam, 20),
  T1(framemapperfint_cc_impl, post_padding_64qam, 20), T1(framemapperfint_cc_impl, post_puncture_bqpsk, 25),
  T1(framemapperfint_cc_impl, post_puncture_16qam, 25), T1(framemapperfint_cc_impl, post_puncture_64qam, 25),
  { "l1_mux16", &framemapperfint_cc_impl::mux16[0], 1, 8 }, { "l1_mux64", &framemapperfint_cc_impl::mux64[0], 1, 12 },
  T1(framemapperfint_cc_impl, bitperm1keven, 9), T1(framemapperfint_cc_impl, bitperm1kodd, 9),
  T1(framemapperfint_cc_impl, bitperm2keven, 10), T1(framemapperfint_cc_impl, bitperm2kodd, 10),
  T1(framemapperfint_cc_impl, bitperm4keven, 11), T1(framemapperfint_cc_impl, bitperm4kodd, 11),
  T1(framemapperfint_cc_impl, bitperm8keven, 12), T1(framemapperfint_cc_impl, bitperm8kodd, 12),
  T1(framemapperfint_cc_impl, bitperm16keven, 13), T1(framemapperfint_cc_impl, bitperm16kodd, 13),
  T1(framemapperfint_cc_impl, bitperm32k, 14),
  T1(interleavermod_bc_impl, twist16n, 8), T1(interleavermod_bc_impl, twist64n, 12),
  T1(interleavermod_bc_impl, twist256n, 16), T1(interleavermod_bc_impl, twist16s, 8),
  T1(interleavermod_bc_impl, twist64s, 12), T1(interleavermod_bc_impl, twist256s, 8),
  T1(interleavermod_bc_impl, mux16, 8), T1(interleavermod_bc_impl, mux64, 12), T1(interleavermod_bc_impl, mux256, 16),
  T1(interleavermod_bc_impl, mux16_35, 8), T1(interleavermod_bc_impl, mux16_13, 8), T1(interleavermod_bc_impl, mux16_25, 8),
  T1(interleavermod_bc_impl, mux64_35, 12), T1(interleavermod_bc_impl, mux64_13, 12), T1(interleavermod_bc_impl, mux64_25, 12),
  T1(interleavermod_bc_impl, mux256_35, 16), T1(interleavermod_bc_impl, mux256_23, 16),
  T1(interleavermod_bc_impl, mux256s, 8), T1(interleavermod_bc_impl, mux256s_13, 8), T1(interleavermod_bc_impl, mux256s_25, 8),
  T1(pilotgenp1insert_cc_impl, p2_papr_map_1k, 10), T1(pilotgenp1insert_cc_impl, p2_papr_map_2k, 18),
  T1(pilotgenp1insert_cc_impl, p2_papr_map_4k, 36), T1(pilotgenp1insert_cc_impl, p2_papr_map_8k, 72),
  T1(pilotgenp1insert_cc_impl, p2_papr_map_16k, 144), T1(pilotgenp1insert_cc_impl, p2_papr_map_32k, 288),
  T1(pilotgenp1insert_cc_impl, tr_papr_map_1k, 10), T1(pilotgenp1insert_cc_impl, tr_papr_map_2k, 18),
  T1(pilotgenp1insert_cc_impl, tr_papr_map_4k, 36), T1(pilotgenp1insert_cc_impl, tr_papr_map_8k, 72),
  T1(pilotgenp1insert_cc_impl, tr_papr_map_16k, 144), T1(pilotgenp1insert_cc_impl, tr_papr_map_32k, 288),
  T1(pilotgenp1insert_cc_impl, pp1_cp1, 20), T1(pilotgenp1insert_cc_impl, pp1_cp2, 25), T1(pilotgenp1insert_cc_impl, pp1_cp5, 44),
  T1(pilotgenp1insert_cc_impl, pp2_cp1, 20), T1(pilotgenp1insert_cc_impl, pp2_cp2, 22), T1(pilotgenp1insert_cc_impl, pp2_cp3, 2),
  T1(pilotgenp1insert_cc_impl, pp2_cp4, 2), T1(pilotgenp1insert_cc_impl, pp2_cp5, 41), T1(pilotgenp1insert_cc_impl, pp2_cp6, 88),
  T1(pilotgenp1insert_cc_impl, pp3_cp1, 22), T1(pilotgenp1insert_cc_impl, pp3_cp2, 20), T1(pilotgenp1insert_cc_impl, pp3_cp3, 1),
  T1(pilotgenp1insert_cc_impl, pp3_cp5, 44), T1(pilotgenp1insert_cc_impl, pp3_cp6, 49),
  T1(pilotgenp1insert_cc_impl, pp4_cp1, 20), T1(pilotgenp1insert_cc_impl, pp4_cp2, 23), T1(pilotgenp1insert_cc_impl, pp4_cp3, 1),
  T1(pilotgenp1insert_cc_impl, pp4_cp4, 2), T1(pilotgenp1insert_cc_impl, pp4_cp5, 44), T1(pilotgenp1insert_cc_impl, pp4_cp6, 86),
  T1(pilotgenp1insert_cc_impl, pp5_cp1, 19), T1(pilotgenp1insert_cc_impl, pp5_cp2, 23), T1(pilotgenp1insert_cc_impl, pp5_cp3, 3),
  T1(pilotgenp1insert_cc_impl, pp5_cp4, 1), T1(pilotgenp1insert_cc_impl, pp5_cp5, 44),
  T1(pilotgenp1insert_cc_impl, pp6_cp5, 88), T1(pilotgenp1insert_cc_impl, pp6_cp6, 88),
  T1(pilotgenp1insert_cc_impl, pp7_cp1, 15), T1(pilotgenp1insert_cc_impl, pp7_cp2, 30), T1(pilotgenp1insert_cc_impl, pp7_cp3, 5),
  T1(pilotgenp1insert_cc_impl, pp7_cp4, 3), T1(pilotgenp1insert_cc_impl, pp7_cp5, 35), T1(pilotgenp1insert_cc_impl, pp7_cp6, 92),
  T1(pilotgenp1insert_cc_impl, pp8_cp4, 47), T1(pilotgenp1insert_cc_impl, pp8_cp5, 39), T1(pilotgenp1insert_cc_impl, pp8_cp6, 89),
  T1(pilotgenp1insert_cc_impl, pp2_8k, 4), T1(pilotgenp1insert_cc_impl, pp3_8k, 2), T1(pilotgenp1insert_cc_impl, pp4_8k, 2),
  T1(pilotgenp1insert_cc_impl, pp7_8k, 5), T1(pilotgenp1insert_cc_impl, pp8_8k, 5),
  T1(pilotgenp1insert_cc_impl, pp1_16k, 4), T1(pilotgenp1insert_cc_impl, pp2_16k, 2), T1(pilotgenp1insert_cc_impl, pp3_16k, 2),
  T1(pilotgenp1insert_cc_impl, pp4_16k, 2), T1(pilotgenp1insert_cc_impl, pp5_16k, 2), T1(pilotgenp1insert_cc_impl, pp6_16k, 2),
  T1(pilotgenp1insert_cc_impl, pp7_16k, 3), T1(pilotgenp1insert_cc_impl, pp8_16k, 3),
  T1(pilotgenp1insert_cc_impl, pp2_32k, 2), T1(pilotgenp1insert_cc_impl, pp4_32k, 2), T1(pilotgenp1insert_cc_impl, pp6_32k, 4),
  T1(pilotgenp1insert_cc_impl, pp7_32k, 2), T1(pilotgenp1insert_cc_impl, pp8_32k, 6),
  T1(pilotgenp1insert_cc_impl, p1_active_carriers, 384),
};

int ref_num_tables(void) { return (int)(sizeof(g_tabs) / sizeof(g_tabs[0])); }
const char *ref_table_name(int i) { return g_tabs[i].name; }
const int *ref_table(int i, int *rows, int *cols) { *rows = g_tabs[i].rows; *cols = g_tabs[i].cols; return g_tabs[i].data; }

/* byte tables */
const unsigned char *ref_byte_table(const char *name, int *len)
{
  if (!strcmp(name, "pn_sequence_table")) { *len = CHIPS / 8; return pilotgenp1insert_cc_impl::pn_sequence_table; }
  if (!strcmp(name, "s1_modulation_patterns")) { *len = 64; return &pilotgenp1insert_cc_impl::s1_modulation_patterns[0][0]; }
  if (!strcmp(name, "s2_modulation_patterns")) { *len = 512; return &pilotgenp1insert_cc_impl::s2_modulation_patterns[0][0]; }
  *len = 0; return 0;
}

double ref_now(void)
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // extern "C"
