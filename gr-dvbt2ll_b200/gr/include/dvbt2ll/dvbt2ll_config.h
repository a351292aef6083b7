// Configuration enums of the dvbt2ll blocks.  The enumerator NAMES and VALUES are ABI: they are what
// the GRC block descriptors and the SWIG/Python layer pass into make() (reference
// include/dvbt2ll/dvbt2ll_config.h:60-202; note FFTSIZE_* is not in size order and FFTSIZE_16K_T2GI = 11).
#ifndef INCLUDED_DVBT2LL_CONFIG_H
#define INCLUDED_DVBT2LL_CONFIG_H

#define FRAME_SIZE_NORMAL 64800
#define FRAME_SIZE_SHORT 16200

namespace gr {
namespace dvbt2ll {

enum dvbt2_code_rate_t { C1_2 = 0, C3_5, C2_3, C3_4, C4_5, C5_6, C1_3, C2_5 };
enum dvbt2_constellation_t { MOD_QPSK = 0, MOD_16QAM, MOD_64QAM, MOD_256QAM };
enum dvbt2_rotation_t { ROTATION_OFF = 0, ROTATION_ON };
enum dvbt2_framesize_t { FECFRAME_SHORT = 0, FECFRAME_NORMAL };
enum dvbt2_streamtype_t { STREAMTYPE_TS = 0, STREAMTYPE_GS, STREAMTYPE_BOTH };
enum dvbt2_inputmode_t { INPUTMODE_NORMAL = 0, INPUTMODE_HIEFF };
enum dvbt2_extended_carrier_t { CARRIERS_NORMAL = 0, CARRIERS_EXTENDED };
enum dvbt2_preamble_t { PREAMBLE_T2_SISO = 0, PREAMBLE_T2_MISO, PREAMBLE_NON_T2, PREAMBLE_T2_LITE_SISO, PREAMBLE_T2_LITE_MISO };
enum dvbt2_fftsize_t { FFTSIZE_2K = 0, FFTSIZE_8K, FFTSIZE_4K, FFTSIZE_1K, FFTSIZE_16K, FFTSIZE_32K,
                       FFTSIZE_8K_T2GI, FFTSIZE_32K_T2GI, FFTSIZE_16K_T2GI = 11 };
enum dvbt2_guardinterval_t { GI_1_32 = 0, GI_1_16, GI_1_8, GI_1_4, GI_1_128, GI_19_128, GI_19_256 };
enum dvbt2_papr_t { PAPR_OFF = 0, PAPR_ACE, PAPR_TR, PAPR_BOTH };
enum dvbt2_l1constellation_t { L1_MOD_BPSK = 0, L1_MOD_QPSK, L1_MOD_16QAM, L1_MOD_64QAM };
enum dvbt2_pilotpattern_t { PILOT_PP1 = 0, PILOT_PP2, PILOT_PP3, PILOT_PP4, PILOT_PP5, PILOT_PP6, PILOT_PP7, PILOT_PP8 };
enum dvbt2_version_t { VERSION_111 = 0, VERSION_121, VERSION_131 };
enum dvbt2_reservedbiasbits_t { RESERVED_OFF = 0, RESERVED_ON };
enum dvbt2_l1scrambled_t { L1_SCRAMBLED_OFF = 0, L1_SCRAMBLED_ON };
enum dvbt2_misogroup_t { MISO_TX1 = 0, MISO_TX2 };
enum dvbt2_showlevels_t { SHOWLEVELS_OFF = 0, SHOWLEVELS_ON };
enum dvbt2_inband_t { INBAND_OFF = 0, INBAND_ON };
enum dvbt2_equalization_t { EQUALIZATION_OFF = 0, EQUALIZATION_ON };
enum dvbt2_bandwidth_t { BANDWIDTH_1_7_MHZ = 0, BANDWIDTH_5_0_MHZ, BANDWIDTH_6_0_MHZ, BANDWIDTH_7_0_MHZ,
                         BANDWIDTH_8_0_MHZ, BANDWIDTH_10_0_MHZ };

} // namespace dvbt2ll
} // namespace gr

// the reference also exposes the enum types at global scope (dvbt2ll_config.h:207-227)
using gr::dvbt2ll::dvbt2_code_rate_t;
using gr::dvbt2ll::dvbt2_constellation_t;
using gr::dvbt2ll::dvbt2_rotation_t;
using gr::dvbt2ll::dvbt2_framesize_t;
using gr::dvbt2ll::dvbt2_streamtype_t;
using gr::dvbt2ll::dvbt2_inputmode_t;
using gr::dvbt2ll::dvbt2_extended_carrier_t;
using gr::dvbt2ll::dvbt2_preamble_t;
using gr::dvbt2ll::dvbt2_fftsize_t;
using gr::dvbt2ll::dvbt2_guardinterval_t;
using gr::dvbt2ll::dvbt2_papr_t;
using gr::dvbt2ll::dvbt2_l1constellation_t;
using gr::dvbt2ll::dvbt2_pilotpattern_t;
using gr::dvbt2ll::dvbt2_version_t;
using gr::dvbt2ll::dvbt2_reservedbiasbits_t;
using gr::dvbt2ll::dvbt2_l1scrambled_t;
using gr::dvbt2ll::dvbt2_misogroup_t;
using gr::dvbt2ll::dvbt2_showlevels_t;
using gr::dvbt2ll::dvbt2_inband_t;
using gr::dvbt2ll::dvbt2_equalization_t;
using gr::dvbt2ll::dvbt2_bandwidth_t;

#endif
