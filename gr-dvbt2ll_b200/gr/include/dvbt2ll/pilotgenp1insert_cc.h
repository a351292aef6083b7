// Public interface of dvbt2ll::pilotgenp1insert_cc -- same class name, base class, sptr typedef and make() signature as the
// reference (include/dvbt2ll/pilotgenp1insert_cc.h:36-49), so SWIG (swig/dvbt2ll_swig.i), the GRC descriptor and existing
// flowgraphs bind to it unchanged.  The implementation behind make() runs on the GPU (lib/pilotgenp1insert_cc_impl.cc).
#ifndef INCLUDED_DVBT2LL_PILOTGENP1INSERT_CC_H
#define INCLUDED_DVBT2LL_PILOTGENP1INSERT_CC_H

#include <dvbt2ll/api.h>
#include <dvbt2ll/dvbt2ll_config.h>
#include <gnuradio/block.h>

namespace gr {
namespace dvbt2ll {

class DVBT2LL_API pilotgenp1insert_cc : virtual public gr::block
{
public:
  typedef boost::shared_ptr<pilotgenp1insert_cc> sptr;
  static sptr make(dvbt2_extended_carrier_t carriermode, dvbt2_fftsize_t fftsize, dvbt2_pilotpattern_t pilotpattern, dvbt2_guardinterval_t guardinterval, int numdatasyms, dvbt2_papr_t paprmode, dvbt2_version_t version, dvbt2_preamble_t preamble, dvbt2_misogroup_t misogroup, dvbt2_equalization_t equalization, dvbt2_bandwidth_t bandwidth, int vlength);
};

} // namespace dvbt2ll
} // namespace gr
#endif
