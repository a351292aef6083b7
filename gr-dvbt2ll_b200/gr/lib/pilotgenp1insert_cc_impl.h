// GPU-backed implementation of dvbt2ll::pilotgenp1insert_cc (replaces reference lib/pilotgenp1insert_cc_impl.{h,cc}).
#ifndef INCLUDED_DVBT2LL_PILOTGENP1INSERT_CC_IMPL_H
#define INCLUDED_DVBT2LL_PILOTGENP1INSERT_CC_IMPL_H

#include <dvbt2ll/pilotgenp1insert_cc.h>

#include "cuda_block.h"

namespace gr {
namespace dvbt2ll {

class pilotgenp1insert_cc_impl : public pilotgenp1insert_cc, public cuda_block_base
{
public:
  pilotgenp1insert_cc_impl(dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_pilotpattern_t pilotpattern, dvbt2_guardinterval_t guardinterval, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_misogroup_t misogroup, dvbt2_equalization_t equalization, dvbt2_bandwidth_t bandwidth, int vlength);
  ~pilotgenp1insert_cc_impl();
  void forecast(int noutput_items, gr_vector_int &ninput_items_required);
  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                   gr_vector_void_star &output_items);

  cuda_block_core &core() { return d_core; }

private:
  cuda_block_core d_core;
};

} // namespace dvbt2ll
} // namespace gr
#endif
