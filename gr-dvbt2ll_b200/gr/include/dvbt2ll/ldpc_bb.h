// dvbt2ll::ldpc_bb -- the LDPC stage of the shipped flowgraph as a block of this module, so the whole chain can stay on
// the GPU.  apps/vv009-4kshort.grc:387-446 wires GNU Radio's in-tree dtv.dvb_ldpc_bb (CPU) between bbheaderbch_bb and
// interleavermod_bc; this block has the same stream contract for the DVB-T2 codes (nbch bits in, 64800 | 16200 bits out,
// one bit per byte, set_output_multiple(frame size)) and takes the two parameters that select a T2 code.  The reference
// restates the encoder (as dead code) at lib/bbheaderbch_bb_impl.cc:533-646.
#ifndef INCLUDED_DVBT2LL_LDPC_BB_H
#define INCLUDED_DVBT2LL_LDPC_BB_H

#include <dvbt2ll/api.h>
#include <dvbt2ll/dvbt2ll_config.h>
#include <gnuradio/block.h>

namespace gr {
namespace dvbt2ll {

class DVBT2LL_API ldpc_bb : virtual public gr::block
{
public:
  typedef boost::shared_ptr<ldpc_bb> sptr;
  static sptr make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate);
};

} // namespace dvbt2ll
} // namespace gr
#endif
