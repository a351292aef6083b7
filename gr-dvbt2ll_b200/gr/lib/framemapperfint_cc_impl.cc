// dvbt2ll::framemapperfint_cc on the GPU: constructor = plan compile (dvbt2ll_framemapperfint_create), forecast()/general_work() = C ABI calls.
// Scheduling contract kept from the reference: set_output_multiple(one frame), forecast as in the reference,
// consume_each(items used), return items produced -- and, unlike the reference, any number of whole frames
// per call is handled correctly (SURVEY.md section 3 "one frame per call").
#include "framemapperfint_cc_impl.h"

namespace gr {
namespace dvbt2ll {

framemapperfint_cc::sptr framemapperfint_cc::make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation, int fecblocks, int tiblocks, dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_guardinterval_t guardinterval, dvbt2_l1constellation_t l1constellation, dvbt2_pilotpattern_t pilotpattern, int t2frames, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_inputmode_t inputmode, dvbt2_reservedbiasbits_t reservedbiasbits, dvbt2_l1scrambled_t l1scrambled, dvbt2_inband_t inband)
{
  return gnuradio::get_initial_sptr(new framemapperfint_cc_impl(framesize, rate, constellation, rotation, fecblocks, tiblocks, carriermode, fftsize, guardinterval, l1constellation, pilotpattern, t2frames, numdatasyms, paprmode, version, preamble, inputmode, reservedbiasbits, l1scrambled, inband));
}

framemapperfint_cc_impl::framemapperfint_cc_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation, int fecblocks, int tiblocks, dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_guardinterval_t guardinterval, dvbt2_l1constellation_t l1constellation, dvbt2_pilotpattern_t pilotpattern, int t2frames, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_inputmode_t inputmode, dvbt2_reservedbiasbits_t reservedbiasbits, dvbt2_l1scrambled_t l1scrambled, dvbt2_inband_t inband)
  : gr::block("framemapperfint_cc", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex)))
{
  d_core.adopt(dvbt2ll_framemapperfint_create(framesize, rate, constellation, rotation, fecblocks, tiblocks, carriermode, fftsize, guardinterval, l1constellation, pilotpattern, t2frames, numdatasyms, paprmode, version, preamble, inputmode, reservedbiasbits, l1scrambled, inband), d_logger, "framemapperfint_cc");
  set_output_multiple(d_core.output_multiple());
}

framemapperfint_cc_impl::~framemapperfint_cc_impl() {}

void framemapperfint_cc_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
  ninput_items_required[0] = d_core.forecast(noutput_items);
}

int framemapperfint_cc_impl::general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                          gr_vector_void_star &output_items)
{
  int consumed = 0;
  const int produced = d_core.work(d_logger, noutput_items, ninput_items[0], input_items[0], output_items[0], &consumed, 0);
  consume_each(consumed);
  return produced;
}

} // namespace dvbt2ll
} // namespace gr
