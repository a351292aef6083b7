// dvbt2ll::bbheaderbch_bb on the GPU: constructor = plan compile (dvbt2ll_bbheaderbch_create), forecast()/general_work() = C ABI calls.
// Scheduling contract kept from the reference: set_output_multiple(one frame), forecast as in the reference,
// consume_each(items used), return items produced -- and, unlike the reference, any number of whole frames
// per call is handled correctly (SURVEY.md section 3 "one frame per call").
#include "bbheaderbch_bb_impl.h"

namespace gr {
namespace dvbt2ll {

bbheaderbch_bb::sptr bbheaderbch_bb::make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_inputmode_t mode, dvbt2_inband_t inband, int fecblocks, int tsrate)
{
  return gnuradio::get_initial_sptr(new bbheaderbch_bb_impl(framesize, rate, mode, inband, fecblocks, tsrate));
}

bbheaderbch_bb_impl::bbheaderbch_bb_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_inputmode_t mode, dvbt2_inband_t inband, int fecblocks, int tsrate)
  : gr::block("bbheaderbch_bb", gr::io_signature::make(1, 1, sizeof(unsigned char)), gr::io_signature::make(1, 1, sizeof(unsigned char)))
{
  d_core.adopt(dvbt2ll_bbheaderbch_create(framesize, rate, mode, inband, fecblocks, tsrate), d_logger, "bbheaderbch_bb");
  set_output_multiple(d_core.output_multiple());
}

bbheaderbch_bb_impl::~bbheaderbch_bb_impl() {}

void bbheaderbch_bb_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
  ninput_items_required[0] = d_core.forecast(noutput_items);
}

int bbheaderbch_bb_impl::general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                          gr_vector_void_star &output_items)
{
  int consumed = 0;
  const int produced = d_core.work(d_logger, noutput_items, ninput_items[0], input_items[0], output_items[0], &consumed, "Transport Stream sync error!");
  consume_each(consumed);
  return produced;
}

} // namespace dvbt2ll
} // namespace gr
