// dvbt2ll::interleavermod_bc on the GPU: constructor = plan compile (dvbt2ll_interleavermod_create), forecast()/general_work() = C ABI calls.
// Scheduling contract kept from the reference: set_output_multiple(one frame), forecast as in the reference,
// consume_each(items used), return items produced -- and, unlike the reference, any number of whole frames
// per call is handled correctly (SURVEY.md section 3 "one frame per call").
#include "interleavermod_bc_impl.h"

namespace gr {
namespace dvbt2ll {

interleavermod_bc::sptr interleavermod_bc::make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation)
{
  return gnuradio::get_initial_sptr(new interleavermod_bc_impl(framesize, rate, constellation, rotation));
}

interleavermod_bc_impl::interleavermod_bc_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation)
  : gr::block("interleavermod_bc", gr::io_signature::make(1, 1, sizeof(unsigned char)), gr::io_signature::make(1, 1, sizeof(gr_complex)))
{
  d_core.adopt(dvbt2ll_interleavermod_create(framesize, rate, constellation, rotation), d_logger, "interleavermod_bc");
  set_output_multiple(d_core.output_multiple());
}

interleavermod_bc_impl::~interleavermod_bc_impl() {}

void interleavermod_bc_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
  ninput_items_required[0] = d_core.forecast(noutput_items);
}

int interleavermod_bc_impl::general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                          gr_vector_void_star &output_items)
{
  int consumed = 0;
  const int produced = d_core.work(d_logger, noutput_items, ninput_items[0], input_items[0], output_items[0], &consumed, 0);
  consume_each(consumed);
  return produced;
}

} // namespace dvbt2ll
} // namespace gr
