// Optional, additive: device-resident hand-off between two adjacent GPU-backed blocks of this module that live in one
// process.  After link(a, b), b takes its input straight from HBM whenever the scheduler hands it the very items a
// wrote last (same host address range), skipping b's host-to-device copy; a's output still goes to the host buffer,
// so every other consumer -- and b itself when the ranges do not match -- sees the usual stream.  Flowgraphs that do
// not call it behave exactly like the reference's.
#ifndef INCLUDED_DVBT2LL_CUDA_LINK_H
#define INCLUDED_DVBT2LL_CUDA_LINK_H

#include <dvbt2ll/api.h>
#include <gnuradio/block.h>

namespace gr {
namespace dvbt2ll {

// returns false when either block is not a GPU-backed dvbt2ll block or their item sizes differ.
// lazy_host: a no longer writes its host output buffer while b keeps taking the items from HBM (items b does not take
// that way are written late, nothing is lost) -- only for an edge whose sole reader is b.
DVBT2LL_API bool link(gr::block *producer, gr::block *consumer, bool lazy_host = false);

} // namespace dvbt2ll
} // namespace gr
#endif
