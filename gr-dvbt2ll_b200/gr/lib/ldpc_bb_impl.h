// GPU-backed dvbt2ll::ldpc_bb (no reference counterpart in lib/: the flowgraph uses gr-dtv's dvb_ldpc_bb there).
#ifndef INCLUDED_DVBT2LL_LDPC_BB_IMPL_H
#define INCLUDED_DVBT2LL_LDPC_BB_IMPL_H

#include <dvbt2ll/ldpc_bb.h>

#include "cuda_block.h"

namespace gr {
namespace dvbt2ll {

class ldpc_bb_impl : public ldpc_bb, public cuda_block_base
{
public:
  ldpc_bb_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate);
  ~ldpc_bb_impl();
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
