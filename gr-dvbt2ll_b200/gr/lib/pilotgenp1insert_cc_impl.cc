// dvbt2ll::pilotgenp1insert_cc on the GPU: constructor = plan compile (dvbt2ll_pilotgenp1insert_create), forecast()/general_work() = C ABI calls.
// Scheduling contract kept from the reference: set_output_multiple(one frame), forecast as in the reference,
// consume_each(items used), return items produced -- and, unlike the reference, any number of whole frames
// per call is handled correctly (SURVEY.md section 3 "one frame per call").
#include "pilotgenp1insert_cc_impl.h"

namespace gr {
namespace dvbt2ll {

pilotgenp1insert_cc::sptr pilotgenp1insert_cc::make(dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_pilotpattern_t pilotpattern, dvbt2_guardinterval_t guardinterval, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_misogroup_t misogroup, dvbt2_equalization_t equalization, dvbt2_bandwidth_t bandwidth, int vlength)
{
  return gnuradio::get_initial_sptr(new pilotgenp1insert_cc_impl(carriermode, fftsize, pilotpattern, guardinterval, numdatasyms, paprmode, version, preamble, misogroup, equalization, bandwidth, vlength));
}

pilotgenp1insert_cc_impl::pilotgenp1insert_cc_impl(dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_pilotpattern_t pilotpattern, dvbt2_guardinterval_t guardinterval, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_misogroup_t misogroup, dvbt2_equalization_t equalization, dvbt2_bandwidth_t bandwidth, int vlength)
  : gr::block("pilotgenp1insert_cc", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex)))
{
  d_core.adopt(dvbt2ll_pilotgenp1insert_create(carriermode, fftsize, pilotpattern, guardinterval, numdatasyms, paprmode, version, preamble, misogroup, equalization, bandwidth, vlength), d_logger, "pilotgenp1insert_cc");
  set_output_multiple(d_core.output_multiple());
}

pilotgenp1insert_cc_impl::~pilotgenp1insert_cc_impl() {}

void pilotgenp1insert_cc_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
  ninput_items_required[0] = d_core.forecast(noutput_items);
}

int pilotgenp1insert_cc_impl::general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                          gr_vector_void_star &output_items)
{
  int consumed = 0;
  const int produced = d_core.work(d_logger, noutput_items, ninput_items[0], input_items[0], output_items[0], &consumed, 0);
  consume_each(consumed);
  return produced;
}

} // namespace dvbt2ll
} // namespace gr
