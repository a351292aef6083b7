// GPU-backed implementation of dvbt2ll::framemapperfint_cc (replaces reference lib/framemapperfint_cc_impl.{h,cc}).
#ifndef INCLUDED_DVBT2LL_FRAMEMAPPERFINT_CC_IMPL_H
#define INCLUDED_DVBT2LL_FRAMEMAPPERFINT_CC_IMPL_H

#include <dvbt2ll/framemapperfint_cc.h>

#include "cuda_block.h"

namespace gr {
namespace dvbt2ll {

class framemapperfint_cc_impl : public framemapperfint_cc, public cuda_block_base
{
public:
  framemapperfint_cc_impl(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation, int fecblocks, int tiblocks, dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_guardinterval_t guardinterval, dvbt2_l1constellation_t l1constellation, dvbt2_pilotpattern_t pilotpattern, int t2frames, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_inputmode_t inputmode, dvbt2_reservedbiasbits_t reservedbiasbits, dvbt2_l1scrambled_t l1scrambled, dvbt2_inband_t inband);
  ~framemapperfint_cc_impl();
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
