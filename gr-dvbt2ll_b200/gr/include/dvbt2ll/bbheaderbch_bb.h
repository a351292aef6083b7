// Public interface of dvbt2ll::bbheaderbch_bb -- same class name, base class, sptr typedef and make() signature as the
// reference (include/dvbt2ll/bbheaderbch_bb.h:36-49), so SWIG (swig/dvbt2ll_swig.i), the GRC descriptor and existing
// flowgraphs bind to it unchanged.  The implementation behind make() runs on the GPU (lib/bbheaderbch_bb_impl.cc).
#ifndef INCLUDED_DVBT2LL_BBHEADERBCH_BB_H
#define INCLUDED_DVBT2LL_BBHEADERBCH_BB_H

#include <dvbt2ll/api.h>
#include <dvbt2ll/dvbt2ll_config.h>
#include <gnuradio/block.h>

namespace gr {
namespace dvbt2ll {

class DVBT2LL_API bbheaderbch_bb : virtual public gr::block
{
public:
  typedef boost::shared_ptr<bbheaderbch_bb> sptr;
  static sptr make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_inputmode_t mode, dvbt2_inband_t inband, int fecblocks, int tsrate);
};

} // namespace dvbt2ll
} // namespace gr
#endif
