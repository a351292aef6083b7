/*
 * TEST INFRASTRUCTURE ONLY -- minimal stand-in for the GNU Radio 3.7 runtime
 * headers, just large enough to compile the UNMODIFIED reference sources
 * (/root/reference/lib/<block>_impl.cc) into oracle/_ref/ without GNU Radio, Boost,
 * FFTW or VOLK.  Nothing under oracle/ is linked into the product library.
 *
 * What the reference uses from <gnuradio/block.h> (see e.g.
 * lib/bbheaderbch_bb_impl.cc:43-45,195,738 in the reference tree):
 *   gr::block(name, in_sig, out_sig), forecast(), general_work(),
 *   set_output_multiple(), consume_each(), d_logger, GR_LOG_WARN/FATAL,
 *   gnuradio::get_initial_sptr, gr_complex / gr_complexd / gr_vector_* typedefs.
 */
#ifndef ORACLE_SHIM_GNURADIO_BLOCK_H
#define ORACLE_SHIM_GNURADIO_BLOCK_H

#include <bitset>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <gnuradio/io_signature.h>

typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace gr {

struct shim_logger {
  int warnings;
  int fatals;
  shim_logger() : warnings(0), fatals(0) {}
};

class block
{
public:
  block(const std::string &name, io_signature::sptr, io_signature::sptr)
    : d_name(name), d_output_multiple(1), d_consumed(0), d_logger(&d_log_store) {}
  /* gr::block is a VIRTUAL base of the public block classes; the intermediate class
   * needs a default constructor to exist (real GNU Radio provides a protected one). */
  block() : d_name(""), d_output_multiple(1), d_consumed(0), d_logger(&d_log_store) {}
  virtual ~block() {}

  virtual void forecast(int, gr_vector_int &) {}
  virtual int general_work(int, gr_vector_int &, gr_vector_const_void_star &,
                           gr_vector_void_star &) { return 0; }

  void set_output_multiple(int m) { d_output_multiple = m; }
  int output_multiple() const { return d_output_multiple; }
  void consume_each(int n) { d_consumed = n; }
  int last_consumed() const { return d_consumed; }
  const std::string &name() const { return d_name; }
  int shim_warnings() const { return d_log_store.warnings; }

protected:
  std::string d_name;
  int d_output_multiple;
  int d_consumed;
  shim_logger d_log_store;
  shim_logger *d_logger;
};

} // namespace gr

extern "C" int oracle_shim_quiet;

#define GR_LOG_WARN(logger, msg)                                   \
  do { (logger)->warnings++;                                       \
       if (!oracle_shim_quiet) fprintf(stderr, "[ref WARN] %s\n", msg); } while (0)
#define GR_LOG_FATAL(logger, msg)                                  \
  do { (logger)->fatals++; fprintf(stderr, "[ref FATAL] %s\n", msg); } while (0)

namespace gnuradio {
template <class T>
boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}

#endif
