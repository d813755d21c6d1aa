// hh_fastnormal.cuh — the native-RNG Box-Muller pair, built for the FP64 pipe of sm_100a.
//
// The FP64 pipe (64 lanes per SM) bounds the Heston kernel, so everything that is not a DFMA is pushed off
// it: range reduction and exponent handling are integer bit operations on the Philox words, the
// logarithm and the sine/cosine read small shared-memory tables (LSU pipe) and finish with short
// polynomials, and the square root is one MUFU.RSQ64H seed plus one cubically convergent refinement.
// No int<->double conversion instructions, no slow-path branches.
//
//   bits -> uniforms (restated bit for bit in oracle/hh_oracle.c: hho_normal_pair)
//     y1 = double{hi = 0x3FF00000 | (w1 & 0xFFFFF), lo = w0 | 1},  u1 = 2 - y1   in [2^-52, 1 - 2^-52]
//     n2 = (w3 & 0xFFFFF) << 32 | w2,                              theta = 2 pi n2 2^-52
//   z1 = sqrt(-2 ln u1) cos(theta),  z2 = sqrt(-2 ln u1) sin(theta)
//
// Accuracy: each normal is within ~4e-16 (absolute) of the libm evaluation of the same formulas, checked
// per path against the oracle by tests/test_gpu_european.py.
#pragma once
#include <mutex>

#include "hh_ctx.h"
#include "hh_device.cuh"
#include "hh_tables.h"

namespace hh {

struct FastNormalTables {
  double2 log_tab[tables::kLogBuckets];  // {rcp_i, -2 ln c_i}
  double2 trig_tab[tables::kTrigN];      // {cos, sin}(2 pi j / 256)
  double exp_tab[tables::kExpN];         // 2 k ln 2
};

// device-global copy of the generated tables; one per translation unit (no -rdc), uploaded once per device
static __device__ FastNormalTables g_fast_tables;

static inline cudaError_t upload_fast_tables(int device, cudaStream_t st) {
  static PerDeviceOnce done;   // per device; contexts on different GPUs may arrive here from different threads
  static std::mutex mu;        // guards the static host image below
  if (done.done(device)) return cudaSuccess;
  std::lock_guard<std::mutex> lk(mu);
  if (done.done(device)) return cudaSuccess;
  static FastNormalTables h;  // static: must outlive the async copy
  for (int i = 0; i < tables::kLogBuckets; ++i) h.log_tab[i] = make_double2(tables::kLogTab[i][0], tables::kLogTab[i][1]);
  for (int i = 0; i < tables::kTrigN; ++i) h.trig_tab[i] = make_double2(tables::kTrigTab[i][0], tables::kTrigTab[i][1]);
  for (int i = 0; i < tables::kExpN; ++i) h.exp_tab[i] = tables::kExpTab[i];
  cudaError_t e = cudaMemcpyToSymbolAsync(g_fast_tables, &h, sizeof h, 0, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) done.set(device);
  return e;
}

__device__ __forceinline__ void load_fast_tables(FastNormalTables *s) {
  const double *src = reinterpret_cast<const double *>(&g_fast_tables);
  double *dst = reinterpret_cast<double *>(s);
  constexpr int n = sizeof(FastNormalTables) / sizeof(double);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// Polynomial coefficients that need all 64 bits live in constant memory so that DFMA takes them as a
// c[bank][offset] operand (no per-iteration UMOV / MOV pairs).
struct FastNormalConsts {
  double third, neg_two_fifths, neg_two_thirds;     // -2 ln(1+r) series
  double two_pi_2m52, magic;                        // angle scaling, 1.5 * 2^52
  double s5, s3;                                    // sin: 1/120, -1/6
  double c6, c4;                                    // cos: -1/720, 1/24
  double tiny;                                      // 1e-300
};
struct FastNormalConsts2 {
  double magic;       // 2^52 + 2^43
  double s1, s3, s5;  // sin(C s) = s (s1 + s^2 (s3 + s^2 s5)), C = 2 pi 2^-52 folded into the coefficients
  double c2, c4, c6;  // cos(C s) = 1 + s^2 (c2 + s^2 (c4 + s^2 c6))
  double l5, l3;      // -2 ln(1+r) = r (-2 + r (1 + r (l3 + r (1/2 + r l5)))), l5 = -2/5, l3 = -2/3
};
__constant__ FastNormalConsts2 kFN2 = {4503599627370496.0 + 8796093022208.0,
                                       0x1.921fb54442d18p-50, -0x1.4abbce625be53p-151, 0x1.466bc6775aae2p-254,
                                       -0x1.3bd3cc9be45dep-100, 0x1.03c1f081b5ac4p-202, -0x1.55d3c7e3cbffap-306,
                                       -0.4, -0x1.5555555555555p-1};
__constant__ FastNormalConsts kFN = {0x1.5555555555555p-2, -0.4, -0x1.5555555555555p-1,
                                     0x1.921fb54442d18p-50, 6755399441055744.0,
                                     0x1.1111111111111p-7, -0x1.5555555555555p-3,
                                     -0x1.6c16c16c16c17p-10, 0x1.5555555555555p-5, 1e-300};

// max(x, 0) and max(x, 1e-300) for finite x through the integer pipe (no DSETP on the FP64 pipe):
// a negative double has its sign bit set; non-negative doubles order like their high words.
__device__ __forceinline__ double max0_bits(double x) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, lo & keep);
}
__device__ __forceinline__ double max_tiny_bits(double x) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const bool small = hi < 0x01a56e1f;  // hi word of 1e-300 (= 0x01a56e1fc2f8f359); also catches negatives
  return __hiloint2double(small ? 0x01a56e1f : hi, small ? (int)0xc2f8f359 : lo);
}

__device__ __forceinline__ double rsqrt_seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RSQ64H
  return y;
}

// sqrt(x) for x in [1e-300, 1e300]: seed + one cubic refinement, 6 FP64 instructions, ~1e-16 relative.
__device__ __forceinline__ double fast_sqrt_pos(double x) {
  const double y0 = rsqrt_seed(x);
  const double t = y0 * y0;
  const double e = fma(-x, t, 1.0);
  const double h = fma(e, 0.375, 0.5);
  const double g = y0 * e;
  const double y = fma(h, g, y0);
  return x * y;
}

// The two halves of the Box-Muller pair, exposed separately for kernels that fold the radius and the rotation into
// their own arithmetic (the Heston step multiplies the radius into its square root and the rotation into the
// correlation factor): R2 = -2 ln(u1), and the angle as (table index j, sin(delta), cos(delta) - 1).
__device__ __forceinline__ double fast_neg2log(const FastNormalTables *__restrict__ tb, uint32_t w0, uint32_t w1) {
  const double y1 = __hiloint2double((int)(0x3FF00000u | (w1 & 0xFFFFFu)), (int)(w0 | 1u));
  const double u1 = 2.0 - y1;  // exact
  const uint32_t uh = (uint32_t)__double2hiint(u1);
  const int k = 1023 - (int)(uh >> 20);
  const uint32_t m20 = uh & 0xFFFFFu;
  const int i = (int)((m20 + 0x1000u) >> 13);
  const double f = __hiloint2double((int)(m20 | 0x3FF00000u), __double2loint(u1));
  const double2 lt = tb->log_tab[i];
  const double L = tb->exp_tab[k - (i >= tables::kLogSplit ? 1 : 0)] + lt.y;
  const double r = fma(f, lt.x, -1.0);
  double p = kFN.third;
  p = fma(p, r, kFN.neg_two_fifths);
  p = fma(p, r, 0.5);
  p = fma(p, r, kFN.neg_two_thirds);
  p = fma(p, r, 1.0);
  p = fma(p, r, -2.0);
  return fma(p, r, L);
}
__device__ __forceinline__ uint32_t fast_angle(uint32_t w2, uint32_t w3, double &sn, double &cm) {
  const uint32_t h2 = w3 & 0xFFFFFu;
  const uint32_t j = (h2 + 0x800u) >> 12;
  const int shi = (int)(0x43380000u + h2) - (int)(j << 12);
  const double sd = __hiloint2double(shi, (int)w2) - kFN.magic;
  const double d = sd * kFN.two_pi_2m52;
  const double d2 = d * d;
  const double d3 = d * d2;
  sn = fma(d3, fma(d2, kFN.s5, kFN.s3), d);
  cm = d2 * fma(d2, fma(d2, kFN.c6, kFN.c4), -0.5);
  return j & 255u;
}

__device__ __forceinline__ void fast_normal_pair(const FastNormalTables *__restrict__ tb, uint32_t w0, uint32_t w1,
                                                 uint32_t w2, uint32_t w3, double &z1, double &z2) {
  // ---- R2 = -2 ln(u1) ---------------------------------------------------------------------------------
  const double y1 = __hiloint2double((int)(0x3FF00000u | (w1 & 0xFFFFFu)), (int)(w0 | 1u));
  const double u1 = 2.0 - y1;  // exact
  const uint32_t uh = (uint32_t)__double2hiint(u1);
  const int k = 1023 - (int)(uh >> 20);              // u1 = f 2^-k, f in [1,2), k >= 1
  const uint32_t m20 = uh & 0xFFFFFu;
  const int i = (int)((m20 + 0x1000u) >> 13);        // nearest of the 129 bucket centres 1 + i/128
  const double f = __hiloint2double((int)(m20 | 0x3FF00000u), __double2loint(u1));
  const double2 lt = tb->log_tab[i];
  const double L = tb->exp_tab[k - (i >= tables::kLogSplit ? 1 : 0)] + lt.y;
  const double r = fma(f, lt.x, -1.0);               // exact up to one rounding: rcp_i has 24 bits
  double p = kFN.third;                              // -2 ln(1+r) = r (-2 + r (1 + r (-2/3 + r (1/2 + r (-2/5 + r/3)))))
  p = fma(p, r, kFN.neg_two_fifths);
  p = fma(p, r, 0.5);
  p = fma(p, r, kFN.neg_two_thirds);
  p = fma(p, r, 1.0);
  p = fma(p, r, -2.0);
  const double R2 = fma(p, r, L);
  const double rad = fast_sqrt_pos(R2);

  // ---- (cos, sin)(2 pi n2 2^-52) ---------------------------------------------------------------------------
  const uint32_t h2 = w3 & 0xFFFFFu;                 // high 20 bits of the 52-bit angle
  const uint32_t j = (h2 + 0x800u) >> 12;            // nearest of 256 table angles (j in [0, 256])
  // s = n2 - j 2^44 in [-2^43, 2^43) as a double via the 1.5*2^52 magic constant (exact)
  const int shi = (int)(0x43380000u + h2) - (int)(j << 12);
  const double sd = __hiloint2double(shi, (int)w2) - kFN.magic;
  const double d = sd * kFN.two_pi_2m52;             // delta = s 2^-52 2 pi
  const double2 cs = tb->trig_tab[j & 255u];
  const double d2 = d * d;
  const double d3 = d * d2;
  const double sn = fma(d3, fma(d2, kFN.s5, kFN.s3), d);           // sin(delta)
  const double cm = d2 * fma(d2, fma(d2, kFN.c6, kFN.c4), -0.5);   // cos(delta) - 1
  const double rc = rad * cs.x, rs = rad * cs.y;
  z1 = fma(rc, cm, fma(-rs, sn, rc));
  z2 = fma(rs, cm, fma(rc, sn, rs));
}

// ---- v2: lane-replicated tables for the headline Heston kernel ------------------------------------------------
// ncu on the v1 kernel (profiles/r1_b_*) showed the step loop bound by shared-memory bank conflicts, not by a math pipe:
// four table reads per step at random indices cost ~40 L1 wavefronts per warp-step (9.9-way conflicts on average)
// against 14 for conflict-free access. v2 stores every 16-byte table entry eight times, once per lane of a quarter
// warp (the unit an LDS.128 is served in), so entry e of lane q sits at byte (e * 8 + q) * 16: the eight lanes of a
// quarter warp always hit eight disjoint groups of four banks, whatever their indices. The v2 tables are also
// indexed by TRUNCATION of the random bits (bucket centres at half-integers), which removes the rounding adds and the
// carry handling from the integer pipe (see tools/gen_tables.py).
struct FastTables2 {
  double2 log_tab[tables::kLog2Buckets];  // {rcp_i, -2 ln c_i}
  double2 trig_tab[tables::kTrigN];       // {cos, sin}(2 pi (j + 1/2) / 256)
  double exp_tab[tables::kExp2N];         // 2 (1023 - e') ln 2
};
static __device__ FastTables2 g_fast_tables2;

static inline cudaError_t upload_fast_tables2(int device, cudaStream_t st) {
  static PerDeviceOnce done;
  static std::mutex mu;
  if (done.done(device)) return cudaSuccess;
  std::lock_guard<std::mutex> lk(mu);
  if (done.done(device)) return cudaSuccess;
  static FastTables2 h;
  for (int i = 0; i < tables::kLog2Buckets; ++i) h.log_tab[i] = make_double2(tables::kLog2Tab[i][0], tables::kLog2Tab[i][1]);
  for (int i = 0; i < tables::kTrigN; ++i) h.trig_tab[i] = make_double2(tables::kTrig2Tab[i][0], tables::kTrig2Tab[i][1]);
  for (int i = 0; i < tables::kExp2N; ++i) h.exp_tab[i] = tables::kExp2Tab[i];
  cudaError_t e = cudaMemcpyToSymbolAsync(g_fast_tables2, &h, sizeof h, 0, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) done.set(device);
  return e;
}

constexpr int kRep = 8;                                                      // copies per entry (lanes of a quarter warp)
constexpr int kLogRepBytes = tables::kLog2Buckets * kRep * 16;               // 32 KB
constexpr int kPhaseRepBytes = tables::kTrigN * 2 * kRep * 16;               // 64 KB: {P1,Q1} then {P2,Q2} per angle
constexpr int kExp2Bytes = ((tables::kExp2N * 8 + 15) / 16) * 16;
constexpr uint32_t kLog2Carry = (1u << 20) - ((uint32_t)tables::kLog2Split << 12);  // mantissa + this carries iff bucket >= split

// max(x, 0) for finite x with ONE integer max on the high word: a negative x becomes a positive denormal-sized value
// (< 2^-1022), which rounds away in every use below exactly like 0 does.
__device__ __forceinline__ double max0_hi(double x) {
  return __hiloint2double(max(__double2hiint(x), 0), __double2loint(x));
}
// max(x, ~1e-300) the same way (the low word is kept: any value in [1e-300, 1.000001e-300] serves as the floor)
__device__ __forceinline__ double max_tiny_hi(double x) {
  return __hiloint2double(max(__double2hiint(x), 0x01a56e1f), __double2loint(x));
}

// R2 = -2 ln(u1), u1 = 2 - y1 with y1 = 1.(w1 low 20 bits | w0 | 1): the same uniform as v1 / the oracle.
// log_lane = replicated log table + (lane & 7) * 16 bytes; exp_biased = exp table - kExp2Bias * 8 bytes.
__device__ __forceinline__ double fast_neg2log_v2(const char *__restrict__ log_lane, const char *__restrict__ exp_biased,
                                                  uint32_t w0, uint32_t w1, uint32_t one_hi) {
  const double y1 = __hiloint2double((int)((w1 & 0xFFFFFu) | one_hi), (int)(w0 | 1u));
  const double u1 = 2.0 - y1;  // exact
  const uint32_t uh = (uint32_t)__double2hiint(u1);
  const uint32_t boff = (uh >> 5) & 0x7F80u;                     // bucket * 128 bytes, bucket = top 8 mantissa bits
  const uint32_t eoff = ((uh + kLog2Carry) >> 17) & 0x3FF8u;     // e' * 8 bytes
  const double f = __hiloint2double((int)((uh & 0xFFFFFu) | one_hi), __double2loint(u1));
  const double2 lt = *reinterpret_cast<const double2 *>(log_lane + boff);
  const double L = *reinterpret_cast<const double *>(exp_biased + eoff) + lt.y;
  const double r = fma(f, lt.x, -1.0);                           // |r| <= 2^-9 + 2^-24: degree 5 leaves r^6/3 < 2e-17
  double p = fma(kFN2.l5, r, 0.5);
  p = fma(p, r, kFN2.l3);
  p = fma(p, r, 1.0);
  p = fma(p, r, -2.0);
  return fma(p, r, L);
}
// Angle theta = 2 pi n2 2^-52, n2 = (w3 low 20 bits, w2): returns the byte offset of table angle j = n2 >> 44 in the
// replicated phase table (j * 256) and sin(delta), cos(delta) for delta = theta - 2 pi (j + 1/2) / 256, |delta| <= pi/256.
__device__ __forceinline__ uint32_t fast_angle_v2(uint32_t w2, uint32_t w3, uint32_t magic_hi, double &sn,
                                                   double &cs) {
  const uint32_t poff = (w3 >> 4) & 0xFF00u;
  // 2^52 + (n2 mod 2^44) exactly, minus (2^52 + 2^43): s in [-2^43, 2^43), delta = C s
  const double sd = __hiloint2double((int)((w3 & 0xFFFu) | magic_hi), (int)w2) - kFN2.magic;
  const double s2 = sd * sd;
  sn = sd * fma(s2, fma(s2, kFN2.s5, kFN2.s3), kFN2.s1);
  cs = fma(s2, fma(s2, fma(s2, kFN2.c6, kFN2.c4), kFN2.c2), 1.0);
  return poff;
}

// ---- HH_RNG_PHILOX_64: 64 random bits per Heston step (one Philox4x32-10 block per TWO steps) -------------------------
// Philox is the largest single cost of the headline kernel (18 IMAD.WIDE + 20 LOP3 of its 93 instructions per step), and
// the 52-bit construction above uses 104 of the 128 bits of a block. The opt-in stream halves that cost: step n takes
// words (wa, wb) = (0, 1) of block n/2 when n is even and words (2, 3) when n is odd (counter stream word 2), as
//   angle   the 52-bit construction fed with wa twice: n2 = (wa & 0xFFFFF) << 32 | wa, theta = 2 pi n2 2^-52
//           = 2 pi (rotl(wa, 12) + d) 2^-32 with a sub-grid dither d = (wa & 0xFFFFF) 2^-20 in [0, 1): 2^32 equally likely
//           angles, one per cell of the 2^32 grid (the angle is periodic: no edge cell) — the same instructions as above;
//   radius  y1 = double{hi = 0x3FF00000 | (wb & 0xFFFFF), lo = (wb & 0xFFF00000) | 0x80000},  u1 = 2 - y1
//           = 1 - (rotl(wb, 12) + 1/2) 2^-32: the midpoints of the 2^32 grid, u1 in [2^-33, 1 - 2^-33], |z| <= 6.77.
// Everything after the bits is the same f64 arithmetic (restated in oracle/hh_oracle.c: hho_normal_pair64).

// R2 = -2 ln(u1) from the two words of y1 (see fast_neg2log_v2 for the table layout)
__device__ __forceinline__ double fast_neg2log_words(const char *__restrict__ log_lane, const char *__restrict__ exp_biased,
                                                     uint32_t y_lo, uint32_t y_hi, uint32_t one_hi) {
  const double y1 = __hiloint2double((int)y_hi, (int)y_lo);
  const double u1 = 2.0 - y1;  // exact
  const uint32_t uh = (uint32_t)__double2hiint(u1);
  const uint32_t boff = (uh >> 5) & 0x7F80u;
  const uint32_t eoff = ((uh + kLog2Carry) >> 17) & 0x3FF8u;
  const double f = __hiloint2double((int)((uh & 0xFFFFFu) | one_hi), __double2loint(u1));
  const double2 lt = *reinterpret_cast<const double2 *>(log_lane + boff);
  const double L = *reinterpret_cast<const double *>(exp_biased + eoff) + lt.y;
  const double r = fma(f, lt.x, -1.0);
  double p = fma(kFN2.l5, r, 0.5);
  p = fma(p, r, kFN2.l3);
  p = fma(p, r, 1.0);
  p = fma(p, r, -2.0);
  return fma(p, r, L);
}
__device__ __forceinline__ double fast_neg2log_32(const char *__restrict__ log_lane, const char *__restrict__ exp_biased,
                                                  uint32_t wb, uint32_t one_hi, uint32_t lo_fill) {
  return fast_neg2log_words(log_lane, exp_biased, (wb & 0xFFF00000u) | lo_fill, (wb & 0xFFFFFu) | one_hi, one_hi);
}
// The full Box-Muller pair from the lane-replicated v2 tables (kernels that need z1 and z2 themselves: the LSM path
// generator). trig_lane = replicated {cos, sin}((j + 1/2) 2 pi / 256) table + (lane & 7) * 16 bytes, entry stride 128 B.
__device__ __forceinline__ double fast_sqrt_pos5(double x);
__device__ __forceinline__ void fast_normal_pair_v2(const char *__restrict__ log_lane, const char *__restrict__ exp_biased,
                                                    const char *__restrict__ trig_lane, uint32_t w0, uint32_t w1, uint32_t w2,
                                                    uint32_t w3, uint32_t one_hi, uint32_t magic_hi, double &z1, double &z2) {
  const double R2 = fast_neg2log_v2(log_lane, exp_biased, w0, w1, one_hi);
  double sn, cs;
  const uint32_t poff = fast_angle_v2(w2, w3, magic_hi, sn, cs) >> 1;  // j * 128
  const double2 t = *reinterpret_cast<const double2 *>(trig_lane + poff);
  const double rad = fast_sqrt_pos5(max_tiny_hi(R2));
  const double c = fma(t.x, cs, -(t.y * sn)), s_ = fma(t.y, cs, t.x * sn);
  z1 = rad * c;
  z2 = rad * s_;
}
// The same pair from 64 random bits (HH_RNG_PHILOX_64: 32-bit angle word wa, 32-bit radius word wb; see above)
__device__ __forceinline__ void fast_normal_pair64_v2(const char *__restrict__ log_lane, const char *__restrict__ exp_biased,
                                                      const char *__restrict__ trig_lane, uint32_t wa, uint32_t wb,
                                                      uint32_t one_hi, uint32_t magic_hi, double &z1, double &z2) {
  const double R2 = fast_neg2log_32(log_lane, exp_biased, wb, one_hi, 0x00080000u);
  double sn, cs;
  const uint32_t poff = fast_angle_v2(wa, wa, magic_hi, sn, cs) >> 1;  // j * 128
  const double2 t = *reinterpret_cast<const double2 *>(trig_lane + poff);
  const double rad = fast_sqrt_pos5(max_tiny_hi(R2));
  const double c = fma(t.x, cs, -(t.y * sn)), s_ = fma(t.y, cs, t.x * sn);
  z1 = rad * c;
  z2 = rad * s_;
}
constexpr int kTrigRepBytes = tables::kTrigN * kRep * 16;  // 32 KB

// sqrt(x), x in [1e-300, 1e300]: MUFU.RSQ64H seed y0 (2^-22.9), then s = s0 (1 + e + 3/2 e^2) with s0 = x y0,
// e = 1/2 - s0 (y0 / 2) = (1 - x y0^2) / 2 — cubic, 5 FP64 instructions; y0 / 2 is an exponent decrement on the ALU.
__device__ __forceinline__ double fast_sqrt_pos5(double x) {
  const double y0 = rsqrt_seed(x);
  const double h0 = __hiloint2double(__double2hiint(y0) - 0x00100000, 0);
  const double s0 = x * y0;
  const double e = fma(-s0, h0, 0.5);
  const double p = fma(e * 1.5, e, e);
  return fma(s0, p, s0);
}

// exp(y) - 1 for the per-step exponent of the GBM generator, |y| <= 1/2: y = j/512 + r, exp(y) - 1 =
// (E_j - 1) + E_j (exp(r) - 1) with {E_j, E_j - 1} tabulated in shared memory (513 entries, filled with libm at kernel
// start) and exp(r) - 1 by its series to r^4 (|r| <= 1/1024: remainder r^5/120 < 7.4e-18). 8 FP64 instructions against ~25 for
// libm's exp; larger |y| (huge sigma sqrt(dt) Z) takes the libm path.
constexpr int kExpJ = 256;
constexpr double kExpScale = 512.0, kExpInvScale = 1.0 / 512.0;
struct ExpCoefs {
  double magic, inv6, inv24, inv120, inv720;
};
__constant__ ExpCoefs kExpC = {6755399441055744.0, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720};

static __device__ __noinline__ double expm1_slow_path(double y) { return exp(y) - 1.0; }

// CHECK = false: the caller has PROVEN |y| <= 1/2 on the host (|drift dt| + |sigma sqrt(dt)| z_max, z_max = 8.6 for the
// 52-bit Box-Muller radius), so the range test and its reconvergence bracket leave the loop (4 instructions per call).
template <bool CHECK = true>
__device__ __forceinline__ double fast_expm1_small(const double2 *__restrict__ tab, double y) {
  // |y| <= 1/2 on the integer pipe (also false for NaN); the out-of-line libm path keeps the loop body small
  if (CHECK && (uint32_t)(__double2hiint(y) & 0x7fffffff) > 0x3fe00000u) return expm1_slow_path(y);
  const double t = fma(y, kExpScale, kExpC.magic);  // nearest integer to 512 y in the low word
  const int j = __double2loint(t);
  const double r = fma(t - kExpC.magic, -kExpInvScale, y);  // exact
  const double2 e = tab[j + kExpJ];
  double p = fma(r, kExpC.inv24, kExpC.inv6);  // r^5/120 <= 7.4e-18 is dropped: below half an ulp of exp(y)
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  return fma(e.x, p * r, e.y);
}

constexpr int kExpm1TabBytes = (2 * kExpJ + 1) * 16;
// fills the {E_j, E_j - 1} table with libm (all threads of the block; the caller synchronises)
__device__ __forceinline__ void fill_expm1_table(double2 *tab) {
  for (int e = threadIdx.x; e <= 2 * kExpJ; e += blockDim.x) {
    const double yj = (double)(e - kExpJ) * kExpInvScale;
    tab[e] = make_double2(exp(yj), expm1(yj));
  }
}

// exp(x) over the whole useful range, |x| < 700 (spots from log-spots, once per saved date): x = (256 k + j) ln2/256 + r,
// exp(x) = 2^k T_j exp(r) with T_j = 2^(j/256) in shared memory (2 KB, filled at kernel start), |r| <= ln2/512 and
// exp(r) by its series to r^4 (remainder r^5/120 < 3.8e-17 relative); the argument reduction uses a two-word ln2/256
// under FMA (exact product), and 2^k is an integer add on the exponent field (T_j exp(r) is in [0.99, 2.01) and
// |k| <= 1010, so the result stays normal). 9 FP64 instructions against ~25 for libm's exp; error <= 1.5 ulp
// (tests/test_fast_exp_emulation.py emulates it). |x| >= 700, NaN and Inf take the libm path.
constexpr int kExpFullN = 256;
constexpr int kExpFullBytes = kExpFullN * 8;
struct ExpFullCoefs {
  double magic, scale, neg_hi, neg_lo, inv6, inv24;
};
__constant__ ExpFullCoefs kExpF = {6755399441055744.0, 369.3299304675746, -0.0027076061740622863, -9.058776616587108e-20,
                                   1.0 / 6, 1.0 / 24};

static __device__ __noinline__ double exp_slow_path(double x) { return exp(x); }

__device__ __forceinline__ double fast_exp_full(const double *__restrict__ tab, double x) {
  if ((uint32_t)(__double2hiint(x) & 0x7fffffff) >= 0x4085E000u) return exp_slow_path(x);
  const double t = fma(x, kExpF.scale, kExpF.magic);  // nearest integer n to x 256/ln2 in the low word
  const int n = __double2loint(t);
  const double nf = t - kExpF.magic;
  double r = fma(nf, kExpF.neg_hi, x);
  r = fma(nf, kExpF.neg_lo, r);
  const double e = tab[n & (kExpFullN - 1)];
  double p = fma(r, kExpF.inv24, kExpF.inv6);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  const double v = fma(e * r, p, e);
  return __hiloint2double(__double2hiint(v) + ((n >> 8) << 20), __double2loint(v));
}

// fills T_j = 2^(j/256) (all threads of the block; the caller synchronises)
__device__ __forceinline__ void fill_exp_full_table(double *tab) {
  for (int e = threadIdx.x; e < kExpFullN; e += blockDim.x) tab[e] = exp2((double)e * (1.0 / kExpFullN));
}

}  // namespace hh
