/*
 * TEST INFRASTRUCTURE ONLY -- gr::fft::fft_complex stand-in.
 *
 * The reference builds fft::fft_complex(N, false, 1) (backward, UNNORMALISED,
 * exp(+j 2 pi k n / N)) at lib/pilotgenp1insert_cc_impl.cc:1155 and :1222 and
 * calls get_inbuf()/get_outbuf()/execute().  FFTW is not in this image, so:
 *   precision mode (default): radix-2 in DOUBLE with exact twiddles, rounded to
 *       float once at the end -- the parity oracle (error ~1e-16 relative before
 *       rounding, i.e. better than FFTW3f itself);
 *   timing mode (oracle_shim_fft_fast != 0): single-precision radix-4 Stockham,
 *       used only when the reference chain is TIMED as the CPU baseline so the
 *       baseline is not handicapped by the double-precision oracle transform.
 */
#ifndef ORACLE_SHIM_GNURADIO_FFT_H
#define ORACLE_SHIM_GNURADIO_FFT_H
#include <complex>
#include <vector>
#include <cmath>
#include <chrono>

extern "C" int oracle_shim_fft_fast;
extern "C" double oracle_shim_fft_seconds;     /* time spent inside execute(): the FFT's share of the CPU baseline */

namespace gr {
namespace fft {

class fft_complex
{
public:
  fft_complex(int n, bool forward, int /*nthreads*/)
    : d_n(n), d_sign(forward ? -1.0 : 1.0), d_in(n), d_out(n), d_w(n / 2), d_wf(n),
      d_tmp(n), d_tmpf(n), d_tmpf2(n)
  {
    const double PI2 = 6.283185307179586476925286766559;
    for (int k = 0; k < n / 2; k++)
      d_w[k] = std::complex<double>(std::cos(PI2 * k / n), d_sign * std::sin(PI2 * k / n));
    for (int k = 0; k < n; k++)
      d_wf[k] = std::complex<float>((float)std::cos(PI2 * k / n),
                                    (float)(d_sign * std::sin(PI2 * k / n)));
    d_log2 = 0;
    while ((1 << d_log2) < n) d_log2++;
  }
  std::complex<float> *get_inbuf() { return d_in.data(); }
  std::complex<float> *get_outbuf() { return d_out.data(); }
  int inbuf_length() const { return d_n; }
  int outbuf_length() const { return d_n; }

  void execute()
  {
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    if (oracle_shim_fft_fast) execute_fast(); else execute_precise();
    oracle_shim_fft_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }

private:
  void execute_precise()
  {
    const int n = d_n;
    for (int i = 0; i < n; i++) {
      int r = 0;
      for (int b = 0; b < d_log2; b++) r |= ((i >> b) & 1) << (d_log2 - 1 - b);
      d_tmp[r] = std::complex<double>(d_in[i].real(), d_in[i].imag());
    }
    for (int len = 2; len <= n; len <<= 1) {
      int half = len >> 1, step = n / len;
      for (int base = 0; base < n; base += len) {
        for (int k = 0; k < half; k++) {
          std::complex<double> w = d_w[k * step];
          std::complex<double> a = d_tmp[base + k], b = d_tmp[base + k + half];
          double tr = b.real() * w.real() - b.imag() * w.imag();
          double ti = b.real() * w.imag() + b.imag() * w.real();
          d_tmp[base + k] = std::complex<double>(a.real() + tr, a.imag() + ti);
          d_tmp[base + k + half] = std::complex<double>(a.real() - tr, a.imag() - ti);
        }
      }
    }
    for (int i = 0; i < n; i++)
      d_out[i] = std::complex<float>((float)d_tmp[i].real(), (float)d_tmp[i].imag());
  }

  /* single-precision Stockham autosort (decimation in frequency), radix-4 passes and
   * one radix-2 pass when log2(n) is odd.  len = current sub-transform length,
   * s = stride (number of interleaved sub-transforms). */
  void execute_fast()
  {
    const int n = d_n;
    typedef std::complex<float> cf;
    cf *x = d_tmpf.data(), *y = d_tmpf2.data();
    for (int i = 0; i < n; i++) x[i] = d_in[i];
    const float sg = (float)d_sign;
    int len = n, s = 1;
    while (len >= 4) {
      const int n1 = len / 4;
      const int tw = n / len;       /* twiddle index step: W_len^p = W_n^(p*tw) */
      for (int p = 0; p < n1; p++) {
        const cf w1 = d_wf[p * tw], w2 = d_wf[2 * p * tw], w3 = d_wf[3 * p * tw];
        for (int q = 0; q < s; q++) {
          const cf a = x[q + s * p], b = x[q + s * (p + n1)];
          const cf c = x[q + s * (p + 2 * n1)], d = x[q + s * (p + 3 * n1)];
          const cf apc = a + c, amc = a - c, bpd = b + d, bmd = b - d;
          const cf jbmd(-sg * bmd.imag(), sg * bmd.real());   /* (sign j) * (b - d) */
          const cf t1 = amc + jbmd, t2 = apc - bpd, t3 = amc - jbmd;
          y[q + s * (4 * p + 0)] = apc + bpd;
          y[q + s * (4 * p + 1)] = cf(t1.real() * w1.real() - t1.imag() * w1.imag(),
                                      t1.real() * w1.imag() + t1.imag() * w1.real());
          y[q + s * (4 * p + 2)] = cf(t2.real() * w2.real() - t2.imag() * w2.imag(),
                                      t2.real() * w2.imag() + t2.imag() * w2.real());
          y[q + s * (4 * p + 3)] = cf(t3.real() * w3.real() - t3.imag() * w3.imag(),
                                      t3.real() * w3.imag() + t3.imag() * w3.real());
        }
      }
      cf *t = x; x = y; y = t;
      len /= 4; s *= 4;
    }
    if (len == 2) {
      for (int q = 0; q < s; q++) {
        const cf a = x[q], b = x[q + s];
        y[q] = a + b;
        y[q + s] = a - b;
      }
      cf *t = x; x = y; y = t;
    }
    for (int i = 0; i < n; i++) d_out[i] = x[i];
  }

  int d_n;
  int d_log2;
  double d_sign;
  std::vector<std::complex<float> > d_in, d_out;
  std::vector<std::complex<double> > d_w;
  std::vector<std::complex<float> > d_wf;
  std::vector<std::complex<double> > d_tmp;
  std::vector<std::complex<float> > d_tmpf, d_tmpf2;
};

} // namespace fft
} // namespace gr
#endif
