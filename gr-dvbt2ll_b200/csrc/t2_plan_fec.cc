// Plan compiler, FEC side: BB framing / scrambler / BCH tables, LDPC rotation lists, bit
// interleaver + demux source map and constellation LUTs.  Host only, runs once per make().
#include "t2_plan.h"

#include <cmath>
#include <complex>
#include <cstring>

#include "t2_std_tables.inc"

namespace t2 {

// ------------------------------------------------------------------------------------------------
// FEC dimensions.  Values of K_bch / N_bch / q per (frame size, rate): EN 302 755 Tables 6a/6b and
// 7a/7b, as selected in reference lib/bbheaderbch_bb_impl.cc:51-165.
// ------------------------------------------------------------------------------------------------
bool fec_spec(int framesize, int rate, FecSpec *o)
{
  struct Row { int rate, kbch, nbch, q, r; };
  static const Row normal[] = {
    { C1_2, 32208, 32400, 90, 192 }, { C3_5, 38688, 38880, 72, 192 }, { C2_3, 43040, 43200, 60, 160 },
    { C3_4, 48408, 48600, 45, 192 }, { C4_5, 51648, 51840, 36, 192 }, { C5_6, 53840, 54000, 30, 160 } };
  static const Row shortf[] = {
    { C1_3, 5232, 5400, 30, 168 },  { C2_5, 6312, 6480, 27, 168 },   { C1_2, 7032, 7200, 25, 168 },
    { C3_5, 9552, 9720, 18, 168 },  { C2_3, 10632, 10800, 15, 168 }, { C3_4, 11712, 11880, 12, 168 },
    { C4_5, 12432, 12600, 10, 168 }, { C5_6, 13152, 13320, 8, 168 } };
  const Row *tab = framesize == FECFRAME_NORMAL ? normal : shortf;
  const int n = framesize == FECFRAME_NORMAL ? 6 : 8;
  if (framesize != FECFRAME_NORMAL && framesize != FECFRAME_SHORT) return false;
  for (int i = 0; i < n; i++) {
    if (tab[i].rate == rate) {
      o->normal = framesize == FECFRAME_NORMAL;
      o->rate = rate;
      o->nldpc = o->normal ? 64800 : 16200;
      o->kbch = tab[i].kbch; o->nbch = tab[i].nbch; o->q = tab[i].q; o->bch_r = tab[i].r;
      return true;
    }
  }
  return false;
}

int cells_per_fecframe(int framesize, int constellation)
{
  if (constellation < MOD_QPSK || constellation > MOD_256QAM) return 0;
  const int bits = 2 * (constellation + 1);
  return (framesize == FECFRAME_NORMAL ? 64800 : 16200) / bits;
}

void bb_prbs_bits(int n, uint8_t *out)
{
  // 15-bit register, taps 14/15 (EN 302 755 5.2.4), start state 100101010000000 -- in the
  // right-shifting representation the reference uses this is 0x4A80.
  unsigned sr = 0x4A80;
  for (int i = 0; i < n; i++) {
    unsigned b = (sr ^ (sr >> 1)) & 1u;
    out[i] = (uint8_t)b;
    sr >>= 1;
    if (b) sr |= 0x4000;
  }
}

uint32_t crc32_bits(const uint8_t *bits, int n)
{
  uint32_t crc = 0xffffffffu;
  for (int i = 0; i < n; i++) {
    uint32_t fb = (bits[i] ^ (crc >> 31)) & 1u;
    crc <<= 1;
    if (fb) crc ^= 0x04C11DB7u;
  }
  return crc;
}

// ------------------------------------------------------------------------------------------------
// BCH.  Minimal polynomials of EN 302 755 Tables 8a (normal, GF(2^16)) and 8b (short, GF(2^14)) as
// bit masks (bit i = coefficient of x^i).
// ------------------------------------------------------------------------------------------------
static const uint32_t kBchFactorsNormal[12] = {
  0x1002d, 0x10173, 0x10fbd, 0x15a55, 0x11f2f, 0x1f7b5, 0x1af65, 0x17367, 0x10ea1, 0x175a7, 0x13a2d, 0x11ae3 };
static const uint32_t kBchFactorsShort[12] = {
  0x402b, 0x4941, 0x4647, 0x5591, 0x6b55, 0x6389, 0x6ce5, 0x4f21, 0x460f, 0x5a49, 0x5811, 0x65ef };

std::vector<uint8_t> bch_generator(int r)
{
  const uint32_t *f = (r == 168) ? kBchFactorsShort : kBchFactorsNormal;
  const int deg = (r == 168) ? 14 : 16;
  const int nf = r / deg;
  std::vector<uint8_t> g(1, 1);
  for (int k = 0; k < nf; k++) {
    std::vector<uint8_t> h(g.size() + deg, 0);
    for (size_t i = 0; i < g.size(); i++)
      if (g[i])
        for (int j = 0; j <= deg; j++)
          if ((f[k] >> j) & 1u) h[i + j] ^= 1;
    g.swap(h);
  }
  return g;   // g.size() == r + 1
}

namespace {

// r-bit remainder register, left aligned in 192 bits: position j <-> coefficient of x^(r-1-j),
// position j stored in w[j >> 5] at bit 31 - (j & 31).
struct Reg192 {
  uint32_t w[6];
  Reg192() { std::memset(w, 0, sizeof(w)); }
  int top() const { return (int)(w[0] >> 31); }
  void shl1() { for (int i = 0; i < 5; i++) w[i] = (w[i] << 1) | (w[i + 1] >> 31); w[5] <<= 1; }
  void xor_in(const Reg192 &o) { for (int i = 0; i < 6; i++) w[i] ^= o.w[i]; }
  void set(int j) { w[j >> 5] |= 1u << (31 - (j & 31)); }
  int get(int j) const { return (int)((w[j >> 5] >> (31 - (j & 31))) & 1u); }
};

Reg192 generator_taps(const std::vector<uint8_t> &g)
{
  const int r = (int)g.size() - 1;
  Reg192 t;
  for (int i = 0; i < r; i++)
    if (g[i]) t.set(r - 1 - i);
  return t;
}

inline void lfsr_step(Reg192 &R, const Reg192 &taps, int in)
{
  const int fb = in ^ R.top();
  R.shl1();
  if (fb) R.xor_in(taps);
}

} // namespace

std::vector<uint8_t> bch_parity_bits(const uint8_t *msg, int k, const std::vector<uint8_t> &g)
{
  const int r = (int)g.size() - 1;
  const Reg192 taps = generator_taps(g);
  Reg192 R;
  for (int i = 0; i < k; i++) lfsr_step(R, taps, msg[i] & 1);
  std::vector<uint8_t> out(r);
  for (int j = 0; j < r; j++) out[j] = (uint8_t)R.get(j);
  return out;
}

bool build_bb_plan(int framesize, int rate, int mode, int inband, int fecblocks, int tsrate, BbPlan *p,
                   std::string *err)
{
  if (!fec_spec(framesize, rate, &p->fec)) {
    if (err) *err = "bbheaderbch_bb: unsupported (framesize, rate) combination";
    return false;
  }
  if (fecblocks < 1) fecblocks = 1;
  p->mode = mode; p->inband = inband; p->fecblocks = fecblocks; p->tsrate = tsrate;
  const FecSpec &f = p->fec;
  p->payload_bytes = (f.kbch - 80) / 8;

  std::vector<uint8_t> prbs(f.kbch);
  bb_prbs_bits(f.kbch, prbs.data());
  p->scramble.assign(f.kbch / 8, 0);
  for (int i = 0; i < f.kbch; i++) p->scramble[i >> 3] |= (uint8_t)(prbs[i] << (7 - (i & 7)));

  for (int i = 0; i < 256; i++) {
    unsigned c = (unsigned)i;
    for (int b = 0; b < 8; b++) c = (c & 0x80) ? ((c << 1) ^ 0xD5) : (c << 1);
    p->crc8_tab[i] = (uint8_t)c;
  }

  const std::vector<uint8_t> g = bch_generator(f.bch_r);
  const Reg192 taps = generator_taps(g);
  // T0[b] = (b(x) x^r) mod g and T1[b] = (b(x) x^(r+8)) mod g: two message bytes are folded per step with two
  // independent look-ups (slicing by 2)
  p->bch_byte_tab.assign(2 * 256 * 6, 0);
  for (int b = 0; b < 256; b++) {
    Reg192 R;
    for (int i = 7; i >= 0; i--) lfsr_step(R, taps, (b >> i) & 1);
    std::memcpy(&p->bch_byte_tab[b * 6], R.w, sizeof(R.w));
    for (int i = 0; i < 8; i++) lfsr_step(R, taps, 0);
    std::memcpy(&p->bch_byte_tab[(256 + b) * 6], R.w, sizeof(R.w));
  }

  const int msg_bytes = f.kbch / 8;
  // bytes per lane: a multiple of 4 with an odd word count, so that the lanes' 32-bit message loads start on word
  // boundaries (up to the common residue msg_bytes mod 4) and fall into 32 different shared-memory banks
  p->chunk_bytes = ((msg_bytes + 31) / 32 + 3) & ~3;
  if (((p->chunk_bytes >> 2) & 1) == 0) p->chunk_bytes += 4;
  p->lead_zero_bytes = 32 * p->chunk_bytes - msg_bytes;

  // rows of "multiply by x^(8 * chunk_bytes) mod g": image of each register position
  std::vector<Reg192> rows(192);
  for (int i = 0; i < f.bch_r; i++) {
    Reg192 R;
    R.set(i);
    for (int s = 0; s < 8 * p->chunk_bytes; s++) lfsr_step(R, taps, 0);   // R <- R * x mod g
    rows[i] = R;
  }
  // column form, laid out for the ballot evaluation in the kernel: for output word w and lane l the
  // lane evaluates output position j = 32 * w + (31 - l); cols[(w * 32 + l) * 6 + c] = word c of
  // the mask of input positions feeding it.
  p->bch_shift_cols.assign(6 * 32 * 6, 0);
  for (int w = 0; w < 6; w++)
    for (int l = 0; l < 32; l++) {
      const int j = 32 * w + (31 - l);
      Reg192 col;
      if (j < f.bch_r)
        for (int i = 0; i < f.bch_r; i++)
          if (rows[i].get(j)) col.set(i);
      std::memcpy(&p->bch_shift_cols[(w * 32 + l) * 6], col.w, sizeof(col.w));
    }

  // in-band type B: "01", 65 zero bits, TS_RATE (27 bits), 10 zero bits = 104 bits
  uint8_t ib[104];
  std::memset(ib, 0, sizeof(ib));
  ib[1] = 1;
  for (int n = 26; n >= 0; n--) ib[2 + 65 + (26 - n)] = (uint8_t)((tsrate >> n) & 1);
  p->inband_bytes.assign(13, 0);
  for (int i = 0; i < 104; i++) p->inband_bytes[i >> 3] |= (uint8_t)(ib[i] << (7 - (i & 7)));
  return true;
}

// ------------------------------------------------------------------------------------------------
// LDPC
// ------------------------------------------------------------------------------------------------
int ldpc_code_index(int normal, int rate)
{
  const int n = (int)(sizeof(kLdpcCodes) / sizeof(kLdpcCodes[0]));
  for (int i = 0; i < n; i++)
    if (kLdpcCodes[i].normal == normal && kLdpcCodes[i].rate == rate) return i;
  return -1;
}

void ldpc_encode_host(int code_index, const uint8_t *info, int nbch, int nldpc, uint8_t *out)
{
  const LdpcCodeDesc &c = kLdpcCodes[code_index];
  const int P = nldpc - nbch;
  std::memcpy(out, info, nbch);
  uint8_t *par = out + nbch;
  std::memset(par, 0, P);
  const uint16_t *a = c.addr;
  int m = 0;
  for (int g = 0; g < c.groups; g++) {
    for (int n = 0; n < 360; n++, m++) {
      if (!info[m]) continue;
      for (int e = 0; e < c.deg[g]; e++) par[(a[e] + n * c.q) % P] ^= 1;
    }
    a += c.deg[g];
  }
  for (int j = 1; j < P; j++) par[j] ^= par[j - 1];
}

bool build_ldpc_plan(int framesize, int rate, LdpcPlan *p, std::string *err)
{
  if (!fec_spec(framesize, rate, &p->fec)) {
    if (err) *err = "ldpc: unsupported (framesize, rate) combination";
    return false;
  }
  const int ci = ldpc_code_index(p->fec.normal, rate);
  if (ci < 0) { if (err) *err = "ldpc: no address table"; return false; }
  const LdpcCodeDesc &c = kLdpcCodes[ci];
  const int q = c.q;
  p->groups = c.groups;
  if (q != p->fec.q || c.groups * 360 != p->fec.nbch) { if (err) *err = "ldpc: table/spec mismatch"; return false; }
  std::vector<std::vector<uint32_t> > rows(q);
  const uint16_t *a = c.addr;
  for (int g = 0; g < c.groups; g++) {
    for (int e = 0; e < c.deg[g]; e++) {
      const int t = a[e] % q, shift = a[e] / q;
      rows[t].push_back(((uint32_t)shift << 16) | (uint32_t)g);
    }
    a += c.deg[g];
  }
  p->row_ptr.assign(q + 1, 0);
  p->entries.clear();
  p->max_row_deg = 0;
  for (int t = 0; t < q; t++) {
    p->row_ptr[t] = (uint16_t)p->entries.size();
    p->entries.insert(p->entries.end(), rows[t].begin(), rows[t].end());
    if ((int)rows[t].size() > p->max_row_deg) p->max_row_deg = (int)rows[t].size();
  }
  p->row_ptr[q] = (uint16_t)p->entries.size();
  return true;
}

// ------------------------------------------------------------------------------------------------
// Bit interleaver + demux + constellation LUT
// ------------------------------------------------------------------------------------------------
static const double kPi = 3.14159265358979323846;

bool build_map_plan(int framesize, int rate, int constellation, int rotation, MapPlan *p, std::string *err)
{
  if (!fec_spec(framesize, rate, &p->fec)) {
    if (err) *err = "interleavermod_bc: unsupported (framesize, rate) combination";
    return false;
  }
  if (constellation < MOD_QPSK || constellation > MOD_256QAM) {
    if (err) *err = "interleavermod_bc: unknown constellation";
    return false;
  }
  const FecSpec &f = p->fec;
  p->constellation = constellation;
  p->rotation = rotation;
  p->mod = 2 * (constellation + 1);
  p->cell_size = f.nldpc / p->mod;
  p->cyclic_delay = rotation ? 1 : 0;
  const int N = f.nldpc, nbch = f.nbch, q = f.q;
  p->bit_src.assign(N, 0);
  p->ncol = 0;
  std::memset(p->col_of_bit, 0, sizeof(p->col_of_bit));
  std::memset(p->twist_of_col, 0, sizeof(p->twist_of_col));

  if (constellation == MOD_QPSK) {
    // reference :289-314: parity interleaving only for short 1/3 and 2/5; otherwise cells are taken
    // from the natural-order codeword, i.e. natural parity index q*s + t  <-  u index 360*t + s.
    const bool pi = (rate == C1_3 || rate == C2_5);
    for (int i = 0; i < N; i++) {
      if (i < nbch || pi) p->bit_src[i] = (uint16_t)i;
      else {
        const int pidx = i - nbch, t = pidx % q, s = pidx / q;
        p->bit_src[i] = (uint16_t)(nbch + 360 * t + s);
      }
    }
  }
  else {
    const uint8_t *twist = 0, *mux = 0;
    int ncol = 2 * p->mod;
    const bool normal = f.normal != 0;
    if (constellation == MOD_16QAM) {
      twist = normal ? kTwist_16n : kTwist_16s;
      mux = (rate == C3_5 && normal) ? kDemux_16_35 : (rate == C1_3 && !normal) ? kDemux_16_13
          : (rate == C2_5 && !normal) ? kDemux_16_25 : kDemux_16;
    }
    else if (constellation == MOD_64QAM) {
      twist = normal ? kTwist_64n : kTwist_64s;
      mux = (rate == C3_5 && normal) ? kDemux_64_35 : (rate == C1_3 && !normal) ? kDemux_64_13
          : (rate == C2_5 && !normal) ? kDemux_64_25 : kDemux_64;
    }
    else if (normal) {
      twist = kTwist_256n;
      mux = (rate == C3_5) ? kDemux_256_35 : (rate == C2_3) ? kDemux_256_23 : kDemux_256;
    }
    else {
      ncol = p->mod;   // 8 columns, one cell per row (reference :626-677)
      twist = kTwist_256s;
      mux = (rate == C1_3) ? kDemux_256s_13 : (rate == C2_5) ? kDemux_256s_25 : kDemux_256s;
    }
    const int rows = N / ncol;
    p->ncol = ncol;
    for (int e = 0; e < ncol; e++) { p->col_of_bit[mux[e]] = (uint8_t)e; p->twist_of_col[e] = twist[e]; }
    // v[rows*c + (twist[c] + r) % rows] = u[rows*c + r]
    std::vector<int> vsrc(N);
    for (int c = 0; c < ncol; c++)
      for (int r = 0; r < rows; r++) vsrc[rows * c + (twist[c] + r) % rows] = rows * c + r;
    // row read-out and demux: output bit position mux[e] of word d takes v[rows*e + d]
    for (int d = 0; d < rows; d++)
      for (int e = 0; e < ncol; e++) p->bit_src[ncol * d + mux[e]] = (uint16_t)vsrc[rows * e + d];
  }

  // ---- constellation LUT, reproducing the reference's arithmetic (double division -> float, then
  // float-complex multiply by the rotation phasor) so cells are bit-identical, not just within 1 ulp.
  static const double l16[4] = { 3.0, 1.0, -3.0, -1.0 };
  static const double l64[8] = { 7.0, 5.0, 1.0, 3.0, -7.0, -5.0, -1.0, -3.0 };
  static const double l256[16] = { 15.0, 13.0, 9.0, 11.0, 1.0, 3.0, 7.0, 5.0, -15.0, -13.0, -9.0, -11.0, -1.0, -3.0, -7.0, -5.0 };
  const int npts = 1 << p->mod;
  p->lut.assign(npts, cfloat());
  double norm, angle;
  const double *lev;
  switch (constellation) {
    case MOD_QPSK:  norm = std::sqrt(2.0);   angle = 29.0;        lev = 0;    break;
    case MOD_16QAM: norm = std::sqrt(10.0);  angle = 16.8;        lev = l16;  break;
    case MOD_64QAM: norm = std::sqrt(42.0);  angle = 8.6;         lev = l64;  break;
    default:        norm = std::sqrt(170.0); angle = 3.576334375; lev = l256; break;
  }
  const int half = p->mod / 2;
  for (int i = 0; i < npts; i++) {
    if (constellation == MOD_QPSK) {
      p->lut[i].re = (float)(((i & 2) ? -1.0 : 1.0) / norm);
      p->lut[i].im = (float)(((i & 1) ? -1.0 : 1.0) / norm);
    }
    else {
      // real part from bits y0,y2,..., imaginary from y1,y3,... (y0 = MSB of the cell word)
      int ri = 0, ii = 0;
      for (int b = 0; b < half; b++) {
        ri = (ri << 1) | ((i >> (p->mod - 1 - 2 * b)) & 1);
        ii = (ii << 1) | ((i >> (p->mod - 2 - 2 * b)) & 1);
      }
      p->lut[i].re = (float)(lev[ri] / norm);
      p->lut[i].im = (float)(lev[ii] / norm);
    }
  }
  if (rotation) {
    const double a = (2.0 * kPi * angle) / 360.0;
    const std::complex<double> ph = std::exp(std::complex<double>(0.0, a));
    const float c = (float)ph.real(), s = (float)ph.imag();
    for (int i = 0; i < npts; i++) {
      const float x = p->lut[i].re, y = p->lut[i].im;
      const float re = x * c - y * s;
      const float im = x * s + y * c;
      p->lut[i].re = re; p->lut[i].im = im;
    }
  }
  // ---- one table instead of two.  With a = level of the I bits and b = level of the Q bits a point is
  // (a c - b s, a s + b c); the word w~ whose I bits are w's Q bits and whose Q bits are w's I bits with the sign bit
  // flipped has levels (b, -a), so Re lut[w~] = b c + a s = Im lut[w] -- bit for bit, the same two rounded products
  // added in the other order.  The kernels keep "word supplying the imaginary part" as w~ and look both parts up in
  // the real-part table; verified here for every word, otherwise the two-table form stays in use.
  {
    uint32_t im = 0, qm = 0;
    for (int b = 0; b < half; b++) { im |= 1u << (p->mod - 1 - 2 * b); qm |= 1u << (p->mod - 2 - 2 * b); }
    p->im_mask_i = im; p->im_mask_q = qm; p->im_flip = 1u << (p->mod - 2);
    p->im_from_re = 1;
    for (int w = 0; w < npts; w++) {
      const uint32_t wt = ((((uint32_t)w << 1) & im) | (((uint32_t)w >> 1) & qm)) ^ p->im_flip;
      uint32_t x, y;
      std::memcpy(&x, &p->lut[wt].re, 4);
      std::memcpy(&y, &p->lut[w].im, 4);
      if (x != y) p->im_from_re = 0;
    }
  }
  // ---- QPSK: the leading run of cells whose two bits sit where they are in the incoming codeword goes through a
  // byte -> four cell codes table in the kernel (the general path extracts bit by bit through bit_src)
  p->qpsk_lin_cells = 0;
  p->qpsk_par_q = 0;
  p->qpsk_nbch = nbch;
  p->qpsk_lut.clear();
  if (constellation == MOD_QPSK) {
    int n = 0;
    while (n < N && p->bit_src[n] == (uint16_t)n) n++;
    if (n > nbch && n < N) n = nbch;           // (parity bit 0 sits in place in either order)
    p->qpsk_lin_cells = (n / 32) * 16;
    // transposed parity: only where bit_src is exactly "info in place, parity bit q s + t <- nbch + 360 t + s", the
    // natural parity stream fits the info part's words and one thread per 32 x 32 tile suffices
    bool tr = n == nbch && nbch % 8 == 0 && N - nbch <= nbch && ((q + 31) / 32) * 12 <= 128;
    for (int i = nbch; i < N && tr; i++) {
      const int pidx = i - nbch;
      if (p->bit_src[i] != (uint16_t)(nbch + 360 * (pidx % q) + pidx / q)) tr = false;
    }
    if (tr) p->qpsk_par_q = q;
    p->qpsk_lut.assign(512, 0);
    for (int b = 0; b < 256; b++) {
      uint32_t code[4];
      for (int k = 0; k < 4; k++) {
        const uint32_t v = ((uint32_t)b >> (6 - 2 * k)) & 3u;
        const uint32_t vt = p->im_from_re ? ((((v << 1) & p->im_mask_i) | ((v >> 1) & p->im_mask_q)) ^ p->im_flip) & 0xFFu : v;
        code[k] = v | (vt << 8);
      }
      p->qpsk_lut[2 * b] = code[0] | (code[1] << 16);
      p->qpsk_lut[2 * b + 1] = code[2] | (code[3] << 16);
    }
  }
  return true;
}

} // namespace t2
