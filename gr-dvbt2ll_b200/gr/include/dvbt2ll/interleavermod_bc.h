// Public interface of dvbt2ll::interleavermod_bc -- same class name, base class, sptr typedef and make() signature as the
// reference (include/dvbt2ll/interleavermod_bc.h:36-49), so SWIG (swig/dvbt2ll_swig.i), the GRC descriptor and existing
// flowgraphs bind to it unchanged.  The implementation behind make() runs on the GPU (lib/interleavermod_bc_impl.cc).
#ifndef INCLUDED_DVBT2LL_INTERLEAVERMOD_BC_H
#define INCLUDED_DVBT2LL_INTERLEAVERMOD_BC_H

#include <dvbt2ll/api.h>
#include <dvbt2ll/dvbt2ll_config.h>
#include <gnuradio/block.h>

namespace gr {
namespace dvbt2ll {

class DVBT2LL_API interleavermod_bc : virtual public gr::block
{
public:
  typedef boost::shared_ptr<interleavermod_bc> sptr;
  static sptr make(dvbt2_framesize_t framesize, dvbt2_code_rate_t rate, dvbt2_constellation_t constellation, dvbt2_rotation_t rotation);
};

} // namespace dvbt2ll
} // namespace gr
#endif
