// Shared plumbing of the five GNU Radio blocks: each block owns one dvbt2ll_handle of libdvbt2ll_cuda.so and
// forwards forecast() / general_work() to the C ABI (include/dvbt2ll_cuda.h).  Host code stays a plain
// gr::block; all DSP runs in CUDA kernels; there is no CPU fallback (a CUDA failure is logged as FATAL and
// thrown, like the reference's allocation failures, lib/framemapperfint_cc_impl.cc:1121-1124).
#ifndef INCLUDED_DVBT2LL_CUDA_BLOCK_H
#define INCLUDED_DVBT2LL_CUDA_BLOCK_H

#include <gnuradio/block.h>
#include <gnuradio/io_signature.h>

#include <cstdlib>
#include <new>
#include <stdexcept>
#include <string>

#include "../../../include/dvbt2ll_cuda.h"

namespace gr {
namespace dvbt2ll {

class cuda_block_core
{
public:
  cuda_block_core() : d_h(0), d_warned(0) {}
  ~cuda_block_core() { if (d_h) dvbt2ll_destroy(d_h); }

  // takes ownership; a NULL handle means make() was given parameters the plan compiler rejects
  template <class Logger>
  void adopt(dvbt2ll_handle *h, Logger &logger, const char *what)
  {
    if (!h) {
      std::string msg = std::string(what) + ": " + dvbt2ll_last_error();
      GR_LOG_FATAL(logger, msg.c_str());
      throw std::invalid_argument(msg);
    }
    d_h = h;
    // the scheduler's buffers are long-lived and reused call after call: let the library page-lock them on first
    // sight, so both copies of general_work() run as DMA at PCIe rate (DVBT2LL_HOST_REGISTER=0 turns it off)
    const char *e = std::getenv("DVBT2LL_HOST_REGISTER");
    dvbt2ll_set_host_register(d_h, (e && e[0] == '0') ? 0 : 1);
  }
  dvbt2ll_handle *handle() const { return d_h; }
  int output_multiple() const { return dvbt2ll_output_multiple(d_h); }
  int forecast(int noutput) const { return dvbt2ll_forecast(d_h, noutput); }

  // general_work body: returns items produced, sets consumed; logs the reference's warnings
  template <class Logger>
  int work(Logger &logger, int noutput_items, int ninput, const void *in, void *out, int *consumed, const char *sync_msg)
  {
    const int r = dvbt2ll_work(d_h, in, ninput, out, noutput_items, consumed);
    if (r < 0) {
      std::string msg = std::string("dvbt2ll CUDA work failed: ") + dvbt2ll_last_error();
      GR_LOG_FATAL(logger, msg.c_str());
      throw std::runtime_error(msg);
    }
    const int w = dvbt2ll_warnings(d_h);
    for (; sync_msg && d_warned < w; d_warned++) GR_LOG_WARN(logger, sync_msg);
    return r;
  }

private:
  dvbt2ll_handle *d_h;
  int d_warned;
};

// every GPU-backed block of the module exposes its core (dvbt2ll/cuda_link.h: device-resident hand-off)
class cuda_block_base
{
public:
  virtual ~cuda_block_base() {}
  virtual cuda_block_core &core() = 0;
};

} // namespace dvbt2ll
} // namespace gr
#endif
