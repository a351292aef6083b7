// Export macro of the gnuradio-dvbt2ll module (same name and meaning as the reference's
// include/dvbt2ll/api.h:27-31, so code that includes <dvbt2ll/api.h> keeps compiling).
#ifndef INCLUDED_DVBT2LL_API_H
#define INCLUDED_DVBT2LL_API_H
#include <gnuradio/attributes.h>
#if defined(gnuradio_dvbt2ll_EXPORTS)
#define DVBT2LL_API __GR_ATTR_EXPORT
#else
#define DVBT2LL_API __GR_ATTR_IMPORT
#endif
#endif
