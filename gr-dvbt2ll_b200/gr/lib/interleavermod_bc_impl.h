// GPU-backed implementation of dvbt2ll::interleavermod_bc (replaces reference lib/interleavermod_bc_impl.{h,cc}).
#ifndef INCLUDED_DVBT2LL_INTERLEAVERMOD_BC_IMPL_H
#define INCLUDED_DVBT2LL_INTERLEAVERMOD_BC_IMPL_H

#include <dvbt2ll/interleavermod_bc.h>

#include "cuda_block.h"

namespace gr {
namespace dvbt2ll {

class interleavermod_bc_impl : public interleavermod_bc, public cuda_block_base
{
public:
  interleavermod_bc_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation);
  ~interleavermod_bc_impl();
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
