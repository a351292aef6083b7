/* TEST INFRASTRUCTURE ONLY -- export macros used by the reference's api.h. */
#ifndef ORACLE_SHIM_GNURADIO_ATTRIBUTES_H
#define ORACLE_SHIM_GNURADIO_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
