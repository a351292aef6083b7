// GPU-backed implementation of dvbt2ll::bbheaderbch_bb (replaces reference lib/bbheaderbch_bb_impl.{h,cc}).
#ifndef INCLUDED_DVBT2LL_BBHEADERBCH_BB_IMPL_H
#define INCLUDED_DVBT2LL_BBHEADERBCH_BB_IMPL_H

#include <dvbt2ll/bbheaderbch_bb.h>

#include "cuda_block.h"

namespace gr {
namespace dvbt2ll {

class bbheaderbch_bb_impl : public bbheaderbch_bb, public cuda_block_base
{
public:
  bbheaderbch_bb_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_inputmode_t mode, dvbt2_inband_t inband, int fecblocks, int tsrate);
  ~bbheaderbch_bb_impl();
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
