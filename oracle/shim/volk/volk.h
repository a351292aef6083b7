/* TEST INFRASTRUCTURE ONLY -- the two VOLK kernels the reference calls
 * (lib/pilotgenp1insert_cc_impl.cc:2888 and :2894), as plain float-complex loops. */
#ifndef ORACLE_SHIM_VOLK_H
#define ORACLE_SHIM_VOLK_H
#include <complex>
typedef std::complex<float> lv_32fc_t;
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t *c, const lv_32fc_t *a,
                                              const lv_32fc_t *b, unsigned int n)
{
  for (unsigned int i = 0; i < n; i++) {
    /* explicit (no FMA-contraction-dependent library call): (ar*br - ai*bi, ar*bi + ai*br) */
    float ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
    c[i] = lv_32fc_t(ar * br - ai * bi, ar * bi + ai * br);
  }
}
static inline void volk_32fc_s32fc_multiply_32fc(lv_32fc_t *c, const lv_32fc_t *a,
                                                 const lv_32fc_t s, unsigned int n)
{
  float sr = s.real(), si = s.imag();
  for (unsigned int i = 0; i < n; i++) {
    float ar = a[i].real(), ai = a[i].imag();
    c[i] = lv_32fc_t(ar * sr - ai * si, ar * si + ai * sr);
  }
}
#endif
