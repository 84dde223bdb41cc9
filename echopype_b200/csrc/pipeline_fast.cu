// Fused pipeline, fast path: host-side dispatch (static conditions, shared-memory sizing, the prepare kernel) of the
// persistent kernel in pipeline_fast_impl.cuh; its instantiations live in pipeline_fast_{f32a,f32b,i16a,i16b}.cu.
#include "pipeline_fast_impl.cuh"

int epb_fast_launch_f32a(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_f32b(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_i16a(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_i16b(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_f32c(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_i16c(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_f32ka(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_f32kb(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_f32kc(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);
int epb_fast_launch_f32w(const void* pr, int T, int G, int noise, int threads, size_t smem, cudaStream_t s);

static int g_grid_reserve = 0;
int epb_grid_reserve() { return g_grid_reserve; }
// SMs the persistent fused kernel leaves free (process-wide; 0 by default): with a ping-sharded plan whose straddle exchange
// runs on a side stream, the NCCL send / recv kernels of step s need somewhere to run while the fused kernel of step s + 1
// holds the rest of the chip.
extern "C" int epb_set_grid_reserve(int sms) {
  g_grid_reserve = sms < 0 ? 0 : sms;
  return 0;
}

namespace {
constexpr int kMaxPingNum = 64;  // two-sweep mode: up to 8 sub-tiles of 8 rows
// sub-tiles per noise tile and rows per sub-tile
inline void sweep_shape(int ping_num, int* S, int* T) {
  *S = (ping_num + kMaxT - 1) / kMaxT;
  *T = (ping_num + *S - 1) / *S;
}
}  // namespace

// Tries to launch the fast path.  Returns 1 when launched (the general kernel must then be launched with the same
// `irregular` flag so that exactly one of the two does the work), 0 when the static conditions do not hold.
// x_i16: x holds int16 raw power counts (-32768 = padding) instead of float32 dB samples.
int epb_pipeline_fast_try(const void* x, int x_i16, const epb_row* rows, const int* xbin, const double* r_edges, int nR,
                          int closed_right, double* acc, float* noise_out, long long C, long long P, long long R,
                          long long nX, int ping_num, int range_sample_num, float noise_max_lin, float snr_lin,
                          double* range_max_out, int sv_input, void* workspace, long long workspace_bytes, cudaStream_t s,
                          float* o_sv, float* o_rng, float* o_svn, float* o_svc) {
  const bool noise = ping_num > 0;
  const bool keep = o_sv || o_rng || o_svn || o_svc;  // full-size outputs: the kKeep instantiations (float32 input only)
  if (keep && (x_i16 || sv_input)) return 0;
  // rows of 4097 .. 8192 samples (float32 input, no full-size outputs): four column groups per thread, single-row tiles;
  // the noise tile then always streams through the ring twice (one row per sub-tile)
  const bool wide = R > 4096;
  if (wide && (x_i16 || keep)) return 0;
  const bool sweep = noise && (ping_num > kMaxT || wide);  // the noise tile streams through the ring twice in sub-tiles
  int S = 1, T = noise ? ping_num : 4;
  if (sweep) sweep_shape(ping_num, &S, &T);
  if (wide) S = noise ? ping_num : 1, T = 1;
  if (ping_num > kMaxPingNum || R % 4 != 0 || R > 8192 || R < 128 || C * nX >= (1LL << 31) || nR > 32000) return 0;
  if (x_i16 && (R % 8 != 0 || sv_input)) return 0;  // 16-byte rows for the bulk copies
  const int xb = x_i16 ? 2 : 4;
  if (noise && range_sample_num < 4) return 0;  // a column group of four may touch at most two range tiles
  const int G = wide ? 4 : (R / 4 > EPB_G1_THREADS) ? EPB_GBIG : 1;  // column groups per thread
  const int threads = (int)(((R / 4 + G - 1) / G + 31) / 32 * 32);
  const int nRt = noise ? (int)((R + range_sample_num - 1) / range_sample_num) : 0;
  // ring: as many tile slots as fit (at least one, at most kMaxTilesInFlight; more than 4 buys nothing)
  // one-group variant: leave room for EPB_G1_BLOCKS resident CTAs per SM (1 KB per CTA is reserved by the driver)
  // leave room for two resident CTAs per SM whenever the register file allows them: one group x <= 512 threads at 64
  // registers, or two groups x <= 256 threads at 128 registers (1 KB per CTA is reserved by the driver)
  const bool two_ctas = (G == 1 && threads <= 512 && EPB_G1_BLOCKS > 1) || (G == 2 && threads <= 256);
  const size_t smem_cap = two_ctas ? (size_t)(227 * 1024) / 2 - 3072 : kSmemMax;
  int nslots = 0;
  for (int n = 4; n >= 1; --n)
    if (fast_smem(R, T, nR, n, nRt, xb, keep) <= smem_cap || (n == 1 && fast_smem(R, T, nR, n, nRt, xb, keep) <= kSmemMax)) {
      nslots = n;
      break;
    }
  if (nslots == 0) return 0;
  const size_t smem = fast_smem(R, T, nR, nslots, nRt, xb, keep);
  FastParams pr;
  pr.x = x, pr.rows = rows, pr.xbin = xbin, pr.edges = r_edges, pr.acc = acc, pr.noise_out = noise_out;
  pr.rmax = range_max_out;
  pr.o_sv = o_sv, pr.o_rng = o_rng, pr.o_svn = o_svn, pr.o_svc = o_svc;
  int* irregular = (int*)workspace;
  pr.irregular = irregular;
  pr.tiles = reinterpret_cast<const TileInfo*>((char*)workspace + 256);
  pr.C = C, pr.P = P, pr.nX = nX;
  pr.PN = noise ? ping_num : T, pr.S = S;
  pr.nPt = (int)((P + pr.PN - 1) / pr.PN);
  pr.ntiles = C * (long long)pr.nPt;
  pr.R = (int)R, pr.nR = nR, pr.rs_num = range_sample_num, pr.closed_right = closed_right, pr.nslots = nslots;
  pr.rt_lanes = 1;
  for (int L : {8, 4, 3, 2})  // the most lanes per range tile with which one pass of the warps covers the row
    if (noise && (long long)(32 / L) * (threads / 32) >= nRt && L <= (range_sample_num + 3) / 4 + 1) {
      pr.rt_lanes = L;
      break;
    }
  pr.rt_tpw = 32 / pr.rt_lanes;
  pr.rt_recip = (65536 + pr.rt_lanes - 1) / pr.rt_lanes;
  pr.noise_max_lin = noise_max_lin;
  pr.snr1 = 1.f + snr_lin;
  pr.sv_input = sv_input;
  const long long ndesc = pr.ntiles * S;
  if (workspace_bytes < 256 + ndesc * (long long)sizeof(TileInfo)) return 0;
  if (cudaMemsetAsync(irregular, 0, sizeof(int), s) != cudaSuccess) return 0;
  prepare_kernel<<<(unsigned)((ndesc + 127) / 128), 128, 0, s>>>(rows, xbin, P, nX, T, pr.PN, S, pr.nPt, ndesc, sv_input,
                                                                 const_cast<TileInfo*>(pr.tiles), irregular);
  int rc = -2;
  if (wide) {
    rc = epb_fast_launch_f32w(&pr, T, G, noise ? 1 : 0, threads, smem, s);
  } else if (keep) {
    if (sweep) {
      rc = epb_fast_launch_f32kc(&pr, T, G, 1, threads, smem, s);
    } else {
      for (auto launcher : {epb_fast_launch_f32ka, epb_fast_launch_f32kb}) {
        rc = launcher(&pr, T, G, noise ? 1 : 0, threads, smem, s);
        if (rc != -2) break;
      }
    }
  } else if (sweep) {
    rc = (x_i16 ? epb_fast_launch_i16c : epb_fast_launch_f32c)(&pr, T, G, 1, threads, smem, s);
  } else {
    for (auto launcher : {x_i16 ? epb_fast_launch_i16a : epb_fast_launch_f32a, x_i16 ? epb_fast_launch_i16b : epb_fast_launch_f32b}) {
      rc = launcher(&pr, T, G, noise ? 1 : 0, threads, smem, s);
      if (rc != -2) break;
    }
  }
  if (rc != 0) cudaMemsetAsync(irregular, 1, sizeof(int), s);  // could not launch: the general kernel does the work
  return 1;
}

long long epb_pipeline_fast_workspace(long long C, long long P, long long R, int ping_num) {
  int PN = ping_num > 0 ? ping_num : 4;
  int S = 1, T = PN;
  if (ping_num > kMaxT && ping_num <= kMaxPingNum) sweep_shape(ping_num, &S, &T);
  if (R > 4096) S = ping_num > 0 ? ping_num : 1, PN = ping_num > 0 ? ping_num : 1;  // one descriptor per row
  return 256 + C * ((P + PN - 1) / PN) * S * (long long)sizeof(TileInfo);
}
