// gr::dvbt2ll::link -- see include/dvbt2ll/cuda_link.h
#include <dvbt2ll/cuda_link.h>

#include "cuda_block.h"

namespace gr {
namespace dvbt2ll {

bool link(gr::block *producer, gr::block *consumer, bool lazy_host)
{
  cuda_block_base *a = dynamic_cast<cuda_block_base *>(producer), *b = dynamic_cast<cuda_block_base *>(consumer);
  if (!a || !b) return false;
  if (dvbt2ll_link(a->core().handle(), b->core().handle()) != 0) return false;
  return !lazy_host || dvbt2ll_link_lazy_host(a->core().handle(), 1) == 0;
}

} // namespace dvbt2ll
} // namespace gr
