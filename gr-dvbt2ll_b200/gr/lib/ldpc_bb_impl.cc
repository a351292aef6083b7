// dvbt2ll::ldpc_bb on the GPU: constructor = plan compile (dvbt2ll_ldpc_create: EN 302 755 address tables in rotation
// form), forecast()/general_work() = C ABI calls.  Any number of whole FECFRAMEs per call.
#include "ldpc_bb_impl.h"

namespace gr {
namespace dvbt2ll {

ldpc_bb::sptr ldpc_bb::make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate)
{
  return gnuradio::get_initial_sptr(new ldpc_bb_impl(framesize, rate));
}

ldpc_bb_impl::ldpc_bb_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate)
  : gr::block("ldpc_bb", gr::io_signature::make(1, 1, sizeof(unsigned char)), gr::io_signature::make(1, 1, sizeof(unsigned char)))
{
  d_core.adopt(dvbt2ll_ldpc_create(framesize, rate), d_logger, "ldpc_bb");
  set_output_multiple(d_core.output_multiple());
}

ldpc_bb_impl::~ldpc_bb_impl() {}

void ldpc_bb_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
  ninput_items_required[0] = d_core.forecast(noutput_items);
}

int ldpc_bb_impl::general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                               gr_vector_void_star &output_items)
{
  int consumed = 0;
  const int produced = d_core.work(d_logger, noutput_items, ninput_items[0], input_items[0], output_items[0], &consumed, 0);
  consume_each(consumed);
  return produced;
}

} // namespace dvbt2ll
} // namespace gr
