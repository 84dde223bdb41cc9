// Row setup: one thread per (channel, ping) folds the calibration / environment parameters of that
// row into an epb_row record and makes the exact index-space decision
// n_start = first n with R' > 0.  O(C*P) work, negligible next to the O(C*P*R) sample kernels.
//
// Reference arithmetic restated: calibrate/calibrate_ek.py:98-183 (CSv / CSp), range.py:160-201 (TVG
// offsets), calibrate/calibrate_azfp.py:64-104, range.py:60-89, calibrate_ek.py:483-490,:613-637.
#include <stdarg.h>

#include "epb_common.cuh"

// ---------------------------------------------------------------------------------------------------
// error plumbing shared by all translation units
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void epb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int epb_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    epb_set_error("%s: %s", what, cudaGetErrorString(e));
    return EPB_E_CUDA;
  }
  return EPB_OK;
}
int epb_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}
extern "C" const char* epb_last_error(void) { return g_err; }
extern "C" int epb_version(void) { return EPB_VERSION; }

namespace {

using namespace epb;

// smallest n in [0, R] with law_tvg_range(row, n) > 0 (monotone non-decreasing in n); R if none.
__device__ int first_positive(const epb_row& r, int R) {
  int lo = 0, hi = R;  // invariant: all n < lo fail, all n >= hi pass (or hi == R)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (law_tvg_range(r, mid) > 0.0)
      hi = mid;
    else
      lo = mid + 1;  // also taken for NaN
  }
  return lo;
}

__device__ void finish_row(epb_row& r, int R, bool guard) {
  r.r0 = law_range(r, 0);
  r.a = (R > 1) ? (law_range(r, R - 1) - r.r0) / (double)(R - 1) : 0.0;
  r.n_start = guard ? first_positive(r, R) : 0;
  r.reserved = 0;
  // float32 hi/lo block for the FP32-pipe sample kernels (include/epb200.h (3))
  auto split = [](double d, float& h, float& l) {
    h = (float)d;
    l = (float)(d - (double)h);  // NaN / inf propagate into h; l becomes NaN for inf, harmless (h decides)
    if (!(fabs(d) <= 3.0e38)) l = 0.f;
  };
  const double kL = 0.33219280948873623;  // log2(10)/10
  split(r.a, r.a_h, r.a_l);
  split(r.r0, r.r0_h, r.r0_l);
  split((r.r0 - r.off1) - r.off2, r.bp_h, r.bp_l);
  r.two_alpha_f = (float)r.two_alpha;
  r.slog2 = (float)(r.slog * 0.3010299956639812);
  r.fscale_f = (float)r.fscale;
  r.foffK = (float)(r.foff - r.K);
  r.c0 = (float)((r.foff - r.K) * kL);
  r.c1 = (float)(r.fscale * kL);
  r.c2 = (float)(r.two_alpha * kL);
  r.spow = (float)(r.slog / 10.0);
  r.range_last = law_range(r, R - 1);
}

// A thread builds one 192-byte row record; a plain struct store would scatter 192-byte pieces.  The records of a warp
// (32 consecutive rows = 6 KB contiguous) go through shared memory and leave as coalesced 16-byte stores.  Every thread
// of the block must call this (blockDim.x == 128).
__device__ __forceinline__ void store_row_coalesced(epb_row* __restrict__ rows, long long nrows, long long i, const epb_row& r) {
  __shared__ __align__(16) epb_row s_rows[128];
  s_rows[threadIdx.x] = r;
  __syncwarp();
  const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
  const long long first = i - lane;  // row of lane 0
  const long long left = nrows - first;
  const int nvalid = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
  const uint4* src = reinterpret_cast<const uint4*>(s_rows + w0);
  uint4* dst = reinterpret_cast<uint4*>(rows + first);
  const int n16 = nvalid * (int)(sizeof(epb_row) / 16);
  for (int k = lane; k < n16; k += 32) dst[k] = src[k];
  __syncwarp();
}

__global__ void rows_ek_power_kernel(epb_row* rows, long long C, long long P, int R, int sonar, int cal_type,
                                     epb_cp dt_, epb_cp c_, epb_cp al_, epb_cp tau_, epb_cp pt_, epb_cp g_,
                                     epb_cp sa_, epb_cp psi_, epb_cp f_, epb_cp te_, const unsigned char* is_gpt) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long i_store = i;  // rows past the end are computed from the last row and not stored
  if (i >= C * P) i = C * P - 1;
  long long c = i / P, p = i % P;
  double dt = cp_at(dt_, c, p), cw = cp_at(c_, c, p), alpha = cp_at(al_, c, p);
  double tau = cp_at(tau_, c, p), pt = cp_at(pt_, c, p), G = cp_at(g_, c, p);
  double f = cp_at(f_, c, p);
  epb_row r;
  r.law = EPB_LAW_EK | EPB_LAW_NANRANGE;
  r.azfp_N = 0;
  r.p0 = dt;
  r.p1 = cw;
  r.p2 = r.p3 = r.p4 = 0.0;
  double ex60 = __dmul_rn(__dmul_rn(2.0, dt), cw) * 0.5;  // 2*dt*c/2        range.py:176-178
  if (sonar == EPB_SONAR_EX60) {
    r.off1 = ex60;
    r.off2 = 0.0;
  } else {
    r.off1 = __dmul_rn(cw, tau) * 0.25;  // c*tau/4                     range.py:180-184
    r.off2 = (is_gpt && is_gpt[c]) ? ex60 : 0.0;  // GPT: both offsets   range.py:193-199
  }
  r.two_alpha = 2.0 * alpha;
  double lam = cw / f;
  if (cal_type == EPB_CAL_SV) {
    double te = cp_at(te_, c, p), psi = cp_at(psi_, c, p), sa = cp_at(sa_, c, p);
    // calibrate_ek.py:154-171
    r.K = 10.0 * log10(pt) + 2.0 * G + psi + 10.0 * log10(lam * lam * te * cw / (32.0 * CUDART_PI * CUDART_PI)) +
          2.0 * sa;
    r.slog = 20.0;
  } else {
    // calibrate_ek.py:176-183
    r.K = 10.0 * log10(pt) + 2.0 * G + 10.0 * log10(lam * lam / (16.0 * CUDART_PI * CUDART_PI));
    r.slog = 40.0;
  }
  r.fscale = 1.0;
  r.foff = 0.0;
  finish_row(r, R, true);
  store_row_coalesced(rows, C * P, i_store, r);
}

__global__ void rows_azfp_kernel(epb_row* rows, long long C, long long P, int R, int cal_type, epb_cp c_,
                                 epb_cp al_, epb_cp tau_, const double* N, const double* fd, const double* L,
                                 const double* EL, const double* DS, const double* TVR, const double* VTX0,
                                 const double* psi_lin, const double* Sv_offset) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long i_store = i;  // rows past the end are computed from the last row and not stored
  if (i >= C * P) i = C * P - 1;
  long long c = i / P, p = i % P;
  double cw = cp_at(c_, c, p), alpha = cp_at(al_, c, p), tau = cp_at(tau_, c, p);
  epb_row r;
  r.law = EPB_LAW_AZFP;
  r.azfp_N = (int)N[c];
  r.p0 = __ddiv_rn(__dmul_rn(cw, L[c]), __dmul_rn(2.0, fd[c]));
  r.p1 = cw * 0.25;
  r.p2 = fd[c];
  r.p3 = tau;
  r.p4 = (cal_type == EPB_CAL_SV) ? 0.0 : __dmul_rn(cw, tau) * 0.25;
  r.off1 = r.off2 = 0.0;
  r.two_alpha = 2.0 * alpha;
  double SL = TVR[c] + 20.0 * log10(VTX0[c]);
  double a = DS[c];
  r.fscale = 1.0 / (26214.0 * a);  // counts / (26214*DS)      calibrate_azfp.py:70-74
  r.foff = EL[c] - 2.5 / a;
  if (cal_type == EPB_CAL_SV) {
    r.K = SL + 10.0 * log10(0.5 * cw * tau * psi_lin[c]) - Sv_offset[c];
    r.slog = 20.0;
  } else {
    r.K = SL;
    r.slog = 40.0;
  }
  finish_row(r, R, false);  // no R' > 0 guard for AZFP (calibrate_azfp.py:64)
  store_row_coalesced(rows, C * P, i_store, r);
}

__global__ void rows_ek80_complex_kernel(epb_row* rows, long long C, long long P, int R, int cal_type, int bb,
                                         int n_beam, epb_cp dt_, epb_cp c_, epb_cp al_, epb_cp tau_, epb_cp pt_,
                                         epb_cp g_, epb_cp sa_, epb_cp psi_, epb_cp f_, epb_cp te_, epb_cp zet_,
                                         epb_cp zer_, const unsigned char* is_gpt) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long i_store = i;  // rows past the end are computed from the last row and not stored
  if (i >= C * P) i = C * P - 1;
  long long c = i / P, p = i % P;
  double dt = cp_at(dt_, c, p), cw = cp_at(c_, c, p), alpha = cp_at(al_, c, p);
  double tau = cp_at(tau_, c, p), pt = cp_at(pt_, c, p), G = cp_at(g_, c, p), f = cp_at(f_, c, p);
  double zet = cp_at(zet_, c, p), zer = cp_at(zer_, c, p);
  epb_row r;
  r.law = EPB_LAW_EK | EPB_LAW_NANRANGE;
  r.azfp_N = 0;
  r.p0 = dt;
  r.p1 = cw;
  r.p2 = r.p3 = r.p4 = 0.0;
  double ex60 = __dmul_rn(__dmul_rn(2.0, dt), cw) * 0.5;
  r.off1 = __dmul_rn(cw, tau) * 0.25;
  r.off2 = (is_gpt && is_gpt[c]) ? ex60 : 0.0;
  r.two_alpha = 2.0 * alpha;
  double lam = cw / f;
  if (cal_type == EPB_CAL_SV) {
    double te = cp_at(te_, c, p), psi = cp_at(psi_, c, p);
    // calibrate_ek.py:613-625
    r.K = 10.0 * log10(lam * lam * pt * cw / (32.0 * CUDART_PI * CUDART_PI)) + 2.0 * G + 10.0 * log10(te) + psi;
    if (!bb) r.K += 2.0 * cp_at(sa_, c, p);
    r.slog = 20.0;
  } else {
    // calibrate_ek.py:629-636
    r.K = 10.0 * log10(lam * lam * pt / (16.0 * CUDART_PI * CUDART_PI)) + 2.0 * G;
    r.slog = 40.0;
  }
  // prx = n_beam * |mean|^2 / (2*sqrt(2))^2 * (|z_er+z_et|/z_er)^2 / z_et      calibrate_ek.py:483-490
  double zr = fabs(zer + zet) / zer;
  r.fscale = (double)n_beam / 8.0 * zr * zr / zet;
  r.foff = 0.0;
  finish_row(r, R, true);
  store_row_coalesced(rows, C * P, i_store, r);
}

int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

}  // namespace

extern "C" int epb_rows_ek_power(epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R, int sonar, int cal_type,
                                 epb_cp dt, epb_cp c, epb_cp alpha, epb_cp tau, epb_cp pt, epb_cp gain, epb_cp sa,
                                 epb_cp psi, epb_cp freq, epb_cp tau_eff, const unsigned char* is_gpt,
                                 void* stream) {
  EPB_REQUIRE(rows && C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad rows/shape");
  EPB_REQUIRE(dt.ptr && c.ptr && alpha.ptr && tau.ptr && pt.ptr && gain.ptr && freq.ptr, "NULL parameter");
  EPB_REQUIRE(cal_type == EPB_CAL_TS || (sa.ptr && psi.ptr && tau_eff.ptr), "Sv needs sa/psi/tau_eff");
  EPB_REQUIRE(sonar == EPB_SONAR_EX60 || sonar == EPB_SONAR_EX80, "bad sonar");
  rows_ek_power_kernel<<<grid_for(C * P, 128), 128, 0, (cudaStream_t)stream>>>(
      rows, C, P, (int)R, sonar, cal_type, dt, c, alpha, tau, pt, gain, sa, psi, freq, tau_eff, is_gpt);
  return epb_check_launch("epb_rows_ek_power");
}

extern "C" int epb_rows_azfp(epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R, int cal_type, epb_cp c, epb_cp alpha,
                             epb_cp tau, const double* N, const double* f_dig, const double* L, const double* EL,
                             const double* DS, const double* TVR, const double* VTX0, const double* psi_linear,
                             const double* Sv_offset, void* stream) {
  EPB_REQUIRE(rows && C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad rows/shape");
  EPB_REQUIRE(c.ptr && alpha.ptr && tau.ptr && N && f_dig && L && EL && DS && TVR && VTX0, "NULL parameter");
  EPB_REQUIRE(cal_type == EPB_CAL_TS || (psi_linear && Sv_offset), "Sv needs psi/Sv_offset");
  rows_azfp_kernel<<<grid_for(C * P, 128), 128, 0, (cudaStream_t)stream>>>(
      rows, C, P, (int)R, cal_type, c, alpha, tau, N, f_dig, L, EL, DS, TVR, VTX0, psi_linear, Sv_offset);
  return epb_check_launch("epb_rows_azfp");
}

extern "C" int epb_rows_ek80_complex(epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R, int cal_type, int waveform_bb,
                                     int n_beam, epb_cp dt, epb_cp c, epb_cp alpha, epb_cp tau, epb_cp pt,
                                     epb_cp gain, epb_cp sa, epb_cp psi, epb_cp freq_center, epb_cp tau_eff,
                                     epb_cp z_et, epb_cp z_er, const unsigned char* is_gpt, void* stream) {
  EPB_REQUIRE(rows && C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad rows/shape");
  EPB_REQUIRE(dt.ptr && c.ptr && alpha.ptr && tau.ptr && pt.ptr && gain.ptr && freq_center.ptr && z_et.ptr &&
                  z_er.ptr,
              "NULL parameter");
  EPB_REQUIRE(cal_type == EPB_CAL_TS || (psi.ptr && tau_eff.ptr && (waveform_bb || sa.ptr)), "Sv needs psi/tau_eff/sa");
  EPB_REQUIRE(n_beam >= 1 && n_beam <= 4, "n_beam must be 1..4");
  rows_ek80_complex_kernel<<<grid_for(C * P, 128), 128, 0, (cudaStream_t)stream>>>(
      rows, C, P, (int)R, cal_type, waveform_bb, n_beam, dt, c, alpha, tau, pt, gain, sa, psi, freq_center, tau_eff,
      z_et, z_er, is_gpt);
  return epb_check_launch("epb_rows_ek80_complex");
}
