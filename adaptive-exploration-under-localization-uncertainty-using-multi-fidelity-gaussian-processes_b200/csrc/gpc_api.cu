// gpc_api.cu -- C ABI of libgpcore.so (see include/gpcore.h).  Host-side orchestration only:
// every numerical step is one of the CUDA kernels in gpc_factor.cuh / gpc_predict.cuh /
// gpc_ig.cuh; there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gpcore.h"
#include "gpc_factor.cuh"
#include "gpc_grad.cuh"
#include "gpc_ig.cuh"
#include "gpc_gridmean.cuh"
#include "gpc_ozaki.cuh"
#include "gpc_predict.cuh"
#include "gpc_traj.cuh"

#define GPC_VERSION 100

// flag |= 1 when a row's fidelity label is not an integer in [0, F) (multi-fidelity models; emukit raises there)
__global__ void __launch_bounds__(256) k_check_fid(const double* __restrict__ X4, long M, int F, int* __restrict__ flag) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= M) return;
  const double f = X4[i * 4 + 3];
  if (!(f >= 0.0) || f >= (double)F || f != floor(f)) atomicOr(flag, 1);
}

namespace {

// Page-locked host staging memory (cudaHostAlloc): the bounce buffers of gpc_predict for pageable caller arrays.
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  double* d() const { return static_cast<double*>(p); }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  double* d() const { return static_cast<double*>(p); }
};

std::string g_create_error;

}  // namespace

// One captured factorisation (Cholesky + triangular inverse of a given buffer set): ~100 launches on three streams
// replayed as a single CUDA graph, which removes the per-launch host cost and most of the gap between dependent kernels.
struct FactorGraph {
  double *A = nullptr, *X = nullptr, *T = nullptr;
  long n_pad = 0;
  int* status = nullptr;
  cudaGraphExec_t exec = nullptr;
  long launches = 0;
};

struct gpc_handle_s {
  int kind = 0, F = 1, device = 0;
  std::vector<FactorGraph> graphs;
  cudaStream_t stream = nullptr, side = nullptr;   // side: look-ahead stream of the factorisation
  cudaStream_t inv = nullptr;                      // early part of the triangular inverse (runs under the Cholesky tail)
  cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_half = nullptr, ev_inv = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};   // host-buffer predict
  GpcHyp hyp;
  bool have_hyp = false, have_data = false, factored = false;
  long N = 0, n_pad = 0;
  int nb = 0;
  double logdet = 0.0, nlml = 0.0;
  long m_chunk = 65536;   // test / candidate rows per launch batch (measured: 16384 -> 65536 = +5 % on configs[1])
  long launches = 0;
  std::string err;
  std::vector<long> perm;  // internal row i holds the caller's training row perm[i]
  double rho[GPC_MAXF] = {1, 1, 1, 1};
  double sigma_y = 0.0;
  int n_noise = 1;
  // model state
  DevBuf Xt, y, extra, L, X, T, alpha, vec, partial, scal, status, Wm, gpart;
  bool have_extra = false;
  // prediction workspaces
  DevBuf Xs4, Kx, meanpart, sumsq, gradpart, mean, var, Vt, cov, grads, ediag;
  PinBuf pin_in, pin_sx, pin_mean, pin_var;   // gpc_predict: page-locked staging of pageable caller arrays (2 stages each)
  // tcgen05 / INT8 path (gpc_ozaki.cuh): digit images of L^-1 and of the current K* chunk
  DevBuf Bimg, sBv, Aimg, Aimg2, meanpart2, gradpart2;
  cudaEvent_t ev_k[2] = {nullptr, nullptr}, ev_v[2] = {nullptr, nullptr};
  cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr;   // device time of the last information-gain call
  double last_call_ms = 0.0;
  bool have_slices = false;
  int mode = GPC_MODE_INT8;
  int n_sm = 0;
  // information-gain workspaces
  DevBuf gX4, gVt, gS, gSinv, gT, Bt, Zt, cand_off, cand_I, cand_aux, cand_rows, cand_mask, gram, gramZ;
  DevBuf gmT, gmA, gmC, gmc, gmax, gmmean;   // tensor-grid mean: axis tables, A rows, C tile rows, coefficients, axes, result
  DevBuf Vimg, gBimg, gsB, sV1, Pt;   // INT8 information gain: digit images of V (candidates) and of the grid's V, scales
  // hot-kernel timing
  bool hot_timing = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  double hot_ms = 0.0;
  double hot_flops = 0.0;
  long hot_launches = 0;
};

namespace {

int fail(gpc_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail(h, GPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                \
  } while (0)

#define CKL()                                                                                           \
  do {                                                                                                  \
    ++h->launches;                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                               \
    if (e__ != cudaSuccess) return fail(h, GPC_ERR_CUDA, std::string("launch: ") + cudaGetErrorString(e__)); \
  } while (0)

inline long round_up(long v, long m) { return (v + m - 1) / m * m; }

int set_gemm_attrs(gpc_handle h) {
  const int sm = gpcg::SMEM_BYTES;
  CK(cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, gpc64::SMEM_BYTES_DEEP));
  CK(cudaFuncSetAttribute(k_syrk_panel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpc64::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_syrk_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpc64::SMEM_BYTES_DEEP));
  CK(cudaFuncSetAttribute(k_linv_level, cudaFuncAttributeMaxDynamicSharedMemorySize, gpc64::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpvt::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpvt::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_cross_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_gram_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_aat_fro, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_kinv, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_ig_seq_cand, cudaFuncAttributeMaxDynamicSharedMemorySize, GPC_IG_SMEM));
  CK(cudaFuncSetAttribute(k_ig_logdet_cand, cudaFuncAttributeMaxDynamicSharedMemorySize, GPC_IG_SMEM));
  CK(cudaFuncSetAttribute(k_ig_selfgrid_cand, cudaFuncAttributeMaxDynamicSharedMemorySize, GPC_IG_SMEM));
  CK(cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, GPC_POTRF_SMEM));
  CK(cudaFuncSetAttribute(k_gm_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, gpc64::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_SUMSQ, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_F64, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_DIGITS, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_F64, true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_SUMSQ, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_F64, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_DIGITS, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_F64, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_SUMSQ, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_F64, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_DIGITS, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_vt_i8<OUT_F64, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, gpoz::SMEM_BYTES));
  CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device));
  return GPC_OK;
}

// Blocked right-looking Cholesky of the n_pad x n_pad matrix A (in place, lower) followed by the
// triangular inverse X = L^-1 (recursive doubling, scratch T).  d_status: device int, 0 = PD.
// Panels (128 columns) are processed in PAIRS: the trailing matrix is updated once per pair with a
// rank-256 product (half the passes over the trailing matrix and a k-loop twice as long as a
// rank-128 update).  Look-ahead: the update first covers the two block columns of the NEXT pair on
// the main stream; the serial chain of that pair -- diagonal block (one CTA), panel solve, rank-128
// update of the pair's second block column, second diagonal block, second panel solve -- then runs
// on the side stream while the main stream finishes the rank-256 update of the rest.
int factor_matrix_launch(gpc_handle h, double* A, double* X, double* T, long n_pad, int* d_status) {
  const int nb = (int)(n_pad / 128);
  cudaStream_t s = h->stream, s2 = h->side;
  const int half = nb / 2;
  // profiling only (bench.py factor_8192): GPC_FACTOR_PHASE=chol stops after the Cholesky (L^-1 is then NOT formed and
  // nothing downstream is valid) so that the Cholesky alone can be timed
  const char* phase_env = std::getenv("GPC_FACTOR_PHASE");
  const bool chol_only = phase_env && std::strcmp(phase_env, "chol") == 0;
  const bool early = !chol_only && nb >= 8 && (nb & (nb - 1)) == 0;   // power-of-two block count: the recursion splits at nb / 2
  CK(cudaMemsetAsync(d_status, 0, sizeof(int), s));
  k_potrf_diag<<<1, GPC_PD_NT, GPC_POTRF_SMEM, s>>>(A, X, n_pad, 0, d_status);
  CKL();
  for (int p = 0; p + 1 < nb; p += 2) {
    cudaStream_t sc = (p == 0) ? s : s2;   // the stream that factored diagonal block p carries the pair's chain
    const int m = nb - p - 1;              // block rows below diagonal block p
    k_trsm_panel<<<2 * m, gpc64::NT, gpc64::SMEM_BYTES_DEEP, sc>>>(A, X, n_pad, p, 2 * (p + 1));
    CKL();
    k_syrk_panel<true><<<dim3(2, 2 * m), gpc64::NT, gpc64::SMEM_BYTES_DEEP, sc>>>(A, n_pad, p, 2 * (p + 1), 2 * (p + 1), 128);
    CKL();
    k_potrf_diag<<<1, GPC_PD_NT, GPC_POTRF_SMEM, sc>>>(A, X, n_pad, p + 1, d_status);
    CKL();
    const int m2 = nb - p - 2;             // block rows / columns behind the pair
    if (m2 <= 0) {
      if (sc != s) {
        CK(cudaEventRecord(h->ev_side, s2));
        CK(cudaStreamWaitEvent(s, h->ev_side, 0));
      }
      break;
    }
    k_trsm_panel<<<2 * m2, gpc64::NT, gpc64::SMEM_BYTES_DEEP, sc>>>(A, X, n_pad, p + 1, 2 * (p + 2));
    CKL();
    if (early && p + 2 == half) {
      // block columns 0 .. half-1 of L are final: invert the leading half and form T[B, A] = L[B, A] X[A, A]
      // of the top level on the third stream, under the (latency-bound) second half of the Cholesky
      CK(cudaEventRecord(h->ev_half, sc));
      CK(cudaStreamWaitEvent(h->inv, h->ev_half, 0));
      for (int sb = 1; sb < half; sb *= 2)
        for (int phase = 0; phase < 2; ++phase) {
          k_linv_level<<<dim3(2 * sb, 2 * sb, half / (2 * sb)), gpc64::NT, gpc64::SMEM_BYTES, h->inv>>>(A, X, T, n_pad, nb, sb,
                                                                                                   phase, 0);
          CKL();
        }
      k_linv_level<<<dim3(2 * half, 2 * half, 1), gpc64::NT, gpc64::SMEM_BYTES, h->inv>>>(A, X, T, n_pad, nb, half, 0, 0);
      CKL();
      CK(cudaEventRecord(h->ev_inv, h->inv));
    }
    if (sc != s) {
      CK(cudaEventRecord(h->ev_side, s2));
      CK(cudaStreamWaitEvent(s, h->ev_side, 0));
    }
    // rank-256 update with panels p, p+1: the next pair's two block columns first ...
    const int la = m2 < 2 ? m2 : 2;
    k_syrk_panel<true><<<dim3(2 * la, 2 * m2), gpc64::NT, gpc64::SMEM_BYTES_DEEP, s>>>(A, n_pad, p, 2 * (p + 2), 2 * (p + 2), 256);
    CKL();
    CK(cudaEventRecord(h->ev_main, s));
    CK(cudaStreamWaitEvent(s2, h->ev_main, 0));
    k_potrf_diag<<<1, GPC_PD_NT, GPC_POTRF_SMEM, s2>>>(A, X, n_pad, p + 2, d_status);
    CKL();
    // ... then the rest, concurrently with the next pair's chain
    if (m2 > la) {
      k_syrk_panel<false><<<dim3(2 * (m2 - la), 2 * (m2 - la)), gpc64::NT, gpc64::SMEM_BYTES, s>>>(
          A, n_pad, p, 2 * (p + 2 + la), 2 * (p + 2 + la), 256);
      CKL();
    }
  }
  if (nb > 1) {
    CK(cudaEventRecord(h->ev_side, s2));
    CK(cudaStreamWaitEvent(s, h->ev_side, 0));
  }
  if (early) {
    // the trailing half's own levels, then the second phase of the top level
    for (int sb = 1; sb < half; sb *= 2)
      for (int phase = 0; phase < 2; ++phase) {
        k_linv_level<<<dim3(2 * sb, 2 * sb, half / (2 * sb)), gpc64::NT, gpc64::SMEM_BYTES, s>>>(A, X, T, n_pad, nb, sb, phase,
                                                                                           half / (2 * sb));
        CKL();
      }
    CK(cudaStreamWaitEvent(s, h->ev_inv, 0));
    k_linv_level<<<dim3(2 * half, 2 * half, 1), gpc64::NT, gpc64::SMEM_BYTES, s>>>(A, X, T, n_pad, nb, half, 1, 0);
    CKL();
    return GPC_OK;
  }
  if (chol_only) return GPC_OK;
  for (int sb = 1; sb < nb; sb *= 2) {
    const int nodes = (nb + 2 * sb - 1) / (2 * sb);
    for (int phase = 0; phase < 2; ++phase) {
      k_linv_level<<<dim3(2 * sb, 2 * sb, nodes), gpc64::NT, gpc64::SMEM_BYTES, s>>>(A, X, T, n_pad, nb, sb, phase, 0);
      CKL();
    }
  }
  return GPC_OK;
}

// The factorisation as a CUDA graph: captured the first time a buffer set is factored (thread-local capture of the
// three streams of factor_matrix_launch), replayed afterwards.  At N = 2048 the chain is launch-latency bound (16
// serial panels x 4 dependent kernels); beyond GPC_GRAPH_MAX_N the kernels are long enough for plain launches.
#define GPC_GRAPH_MAX_N 4096
int factor_matrix(gpc_handle h, double* A, double* X, double* T, long n_pad, int* d_status) {
  static const bool no_graph = std::getenv("GPC_NO_GRAPH") != nullptr;
  if (n_pad > GPC_GRAPH_MAX_N || no_graph || std::getenv("GPC_FACTOR_PHASE")) return factor_matrix_launch(h, A, X, T, n_pad, d_status);
  for (const FactorGraph& g : h->graphs)
    if (g.A == A && g.X == X && g.T == T && g.n_pad == n_pad && g.status == d_status) {
      CK(cudaGraphLaunch(g.exec, h->stream));
      h->launches += g.launches;
      return GPC_OK;
    }
  const long l0 = h->launches;
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
    const int rc = factor_matrix_launch(h, A, X, T, n_pad, d_status);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == GPC_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc == GPC_OK && e == cudaSuccess && exec) {
      FactorGraph g;
      g.A = A; g.X = X; g.T = T; g.n_pad = n_pad; g.status = d_status; g.exec = exec; g.launches = h->launches - l0;
      if (h->graphs.size() >= 8) {            // buffers were re-allocated a few times: drop the oldest capture
        cudaGraphExecDestroy(h->graphs.front().exec);
        h->graphs.erase(h->graphs.begin());
      }
      h->graphs.push_back(g);
      h->launches = l0;
      CK(cudaGraphLaunch(exec, h->stream));
      h->launches += g.launches;
      return GPC_OK;
    }
    cudaGetLastError();                       // capture or instantiation failed: plain launches
    h->launches = l0;
  } else {
    cudaGetLastError();
  }
  return factor_matrix_launch(h, A, X, T, n_pad, d_status);
}

// out = M^T x (+ bias) via two-stage partial sums.
int trmv_t(gpc_handle h, const double* M, const double* x, const double* bias, double bias_scale, double sign,
           double* out) {
  const int nb = h->nb;
  k_trmv_t_partial<<<dim3(nb, nb), 512, 0, h->stream>>>(M, h->n_pad, x, h->partial.d(), h->n_pad);
  CKL();
  k_colsum_partial<<<(unsigned)((h->n_pad + 255) / 256), 256, 0, h->stream>>>(h->partial.d(), nb, h->n_pad, bias,
                                                                            bias_scale, sign, out);
  CKL();
  return GPC_OK;
}

// alpha = L^-T L^-1 y with the explicit inverse as the solver and one step of iterative
// refinement against L itself per triangular system (restores a backward-stable residual).
int solve_alpha(gpc_handle h) {
  const long np = h->n_pad;
  cudaStream_t s = h->stream;
  double* z = h->vec.d();
  double* r = z + np;
  double* dz = z + 2 * np;
  double* al = h->alpha.d();
  const double *L = h->L.d(), *X = h->X.d(), *y = h->y.d();
  const unsigned gr = (unsigned)((np + 7) / 8), ga = (unsigned)((np + 255) / 256);
  k_trmv_n<<<gr, 256, 0, s>>>(X, np, np, y, nullptr, 0.0, 1.0, z);  CKL();
  k_trmv_n<<<gr, 256, 0, s>>>(L, np, np, z, y, 1.0, -1.0, r);        CKL();
  k_trmv_n<<<gr, 256, 0, s>>>(X, np, np, r, nullptr, 0.0, 1.0, dz);  CKL();
  k_axpy<<<ga, 256, 0, s>>>(z, dz, np);                              CKL();
  int rc;
  if ((rc = trmv_t(h, X, z, nullptr, 0.0, 1.0, al))) return rc;
  if ((rc = trmv_t(h, L, al, z, 1.0, -1.0, r))) return rc;
  if ((rc = trmv_t(h, X, r, nullptr, 0.0, 1.0, dz))) return rc;
  k_axpy<<<ga, 256, 0, s>>>(al, dz, np);                             CKL();
  return GPC_OK;
}

// Fidelity labels of caller rows (x, y, z, fid): integers in [0, F) for a multi-fidelity model (emukit raises on
// anything else); single-fidelity models ignore the column.
int check_fidelity_rows(gpc_handle h, const double* X4, long n, const char* what) {
  if (h->F == 1 || !X4) return GPC_OK;
  const double F = (double)h->F;
  for (long i = 0; i < n; ++i) {
    const double f = X4[i * 4 + 3];
    if (!(f >= 0.0) || f >= F || f != std::floor(f))
      return fail(h, GPC_ERR_ARG, std::string(what) + ": fidelity index must be an integer in [0, F)");
  }
  return GPC_OK;
}

int require_factor(gpc_handle h) {
  if (!h) return GPC_ERR_ARG;
  if (!h->factored) return fail(h, GPC_ERR_STATE, "model not factored: call gpc_factor first");
  return GPC_OK;
}

int ensure_pred_ws(gpc_handle h, long m_pad, bool need_vt, bool need_grad) {
  const long np = h->n_pad;
  const int nchunks = (int)((np + KS_COLS - 1) / KS_COLS);
  CK(h->Kx.ensure((size_t)m_pad * np * 8));
  CK(h->meanpart.ensure((size_t)nchunks * m_pad * 8));
  CK(h->sumsq.ensure((size_t)(2 * h->nb) * m_pad * 8));
  if (need_vt) CK(h->Vt.ensure((size_t)m_pad * np * 8));
  if (need_grad) CK(h->gradpart.ensure((size_t)nchunks * 3 * m_pad * 8));
  return GPC_OK;
}

template <bool WITH_GRAD, bool STORE_K>
int launch_kstar(gpc_handle h, const double* dXs4, long M, long m_pad) {
  const long np = h->n_pad;
  const int nchunks = (int)((np + KS_COLS - 1) / KS_COLS);
  k_kstar<WITH_GRAD, STORE_K><<<dim3((unsigned)(m_pad / KS_ROWS), nchunks), 256, 0, h->stream>>>(
      h->hyp, h->Xt.d(), h->alpha.d(), h->N, np, dXs4, M, m_pad, h->Kx.d(), h->meanpart.d(),
      WITH_GRAD ? h->gradpart.d() : nullptr);
  CKL();
  return GPC_OK;
}

// V^T = A X^T for an [m_pad][ld] operand A and an explicit lower-triangular inverse X (ld x ld):
// the dominant DMMA contraction, timed with CUDA events on the handle's stream when enabled.
template <bool STORE_V, bool SUMSQ>
int launch_vt_on(gpc_handle h, const double* A, const double* X, long ld, long m_pad, double* Vt, double* sumsq) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->hot_timing) {
    if (h->ev_used + 2 > h->ev.size()) {
      for (int i = 0; i < 64; ++i) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->ev.push_back(e);
      }
    }
    e0 = h->ev[h->ev_used++];
    e1 = h->ev[h->ev_used++];
    CK(cudaEventRecord(e0, h->stream));
  }
  const int nb2 = (int)(ld / gpvt::BN);
  k_vt<STORE_V, SUMSQ><<<dim3((unsigned)((nb2 + 1) / 2), (unsigned)(m_pad / gpvt::BM)), gpvt::NT, gpvt::SMEM_BYTES,
                         h->stream>>>(A, X, ld, nb2, m_pad, Vt, sumsq);
  CKL();
  if (h->hot_timing) {
    CK(cudaEventRecord(e1, h->stream));
    // 2 * m_pad * 64 * sum_jb (jb+1) 64, minus the quarter of every diagonal tile that is skipped
    h->hot_flops += (double)m_pad * (double)ld * (double)(ld + 32);
  }
  return GPC_OK;
}

template <bool STORE_V, bool SUMSQ>
int launch_vt(gpc_handle h, long m_pad) {
  return launch_vt_on<STORE_V, SUMSQ>(h, h->Kx.d(), h->X.d(), h->n_pad, m_pad, h->Vt.d(), h->sumsq.d());
}

// INT8 digits of the rows of L^-1 (built lazily once per factorisation) and the K* scale.
int ensure_slices(gpc_handle h) {
  if (h->have_slices) return GPC_OK;
  const long np = h->n_pad;
  CK(h->Bimg.ensure((size_t)np * np * gpoz::S));
  CK(h->sBv.ensure((size_t)np * 8));
  k_slice_rows<<<(unsigned)(np / 8), 256, 0, h->stream>>>(h->X.d(), np, np, static_cast<int8_t*>(h->Bimg.p), h->sBv.d(), np, 1);
  CKL();
  h->have_slices = true;
  return GPC_OK;
}

double kstar_scale(gpc_handle h) {
  double kmax = 0.0;
  for (int i = 0; i < h->F; ++i) kmax = std::fmax(kmax, h->hyp.kdiag[i]);  // |k(a, b)| <= max prior variance
  return kmax > 0.0 ? kmax * gpoz::SCALE_HEADROOM : 1.0;  // |k| / sA <= 0.4975 (Cauchy-Schwarz: |k(a, b)| <= max prior variance)
}

// Scale of the digits of V = L^-1 K*: sum_i V_i^2 <= k(x*, x*), so |V| <= sqrt(max prior variance).
double v_scale(gpc_handle h) {
  double kmax = 0.0;
  for (int i = 0; i < h->F; ++i) kmax = std::fmax(kmax, h->hyp.kdiag[i]);
  return kmax > 0.0 ? std::sqrt(kmax) * gpoz::SCALE_HEADROOM : 1.0;  // |V| / sV <= 0.4975
}

// Digit levels of the INT8 contraction: 6 (21 digit GEMMs, FP64 results) or 4 in the FP32-tolerance mode (10 GEMMs).
inline int i8_levels(gpc_handle h) { return h->mode == GPC_MODE_INT8_F32 ? 4 : (h->mode == GPC_MODE_INT8_L5 ? 5 : gpoz::S); }

template <int OUT, bool FULLK>
int launch_i8(gpc_handle h, const VtI8Args& a, double fp64_equiv_flops) {
  const int grid = a.n_items < h->n_sm ? a.n_items : h->n_sm;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->hot_timing) {
    if (h->ev_used + 2 > h->ev.size()) {
      for (int i = 0; i < 64; ++i) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->ev.push_back(e);
      }
    }
    e0 = h->ev[h->ev_used++];
    e1 = h->ev[h->ev_used++];
    CK(cudaEventRecord(e0, h->stream));
  }
  if (a.nlev == gpoz::S) k_vt_i8<OUT, FULLK, gpoz::S><<<grid, gpoz::NT, gpoz::SMEM_BYTES, h->stream>>>(a);
  else if (a.nlev == 5) k_vt_i8<OUT, FULLK, 5><<<grid, gpoz::NT, gpoz::SMEM_BYTES, h->stream>>>(a);
  else k_vt_i8<OUT, FULLK, 4><<<grid, gpoz::NT, gpoz::SMEM_BYTES, h->stream>>>(a);
  CKL();
  if (h->hot_timing) {
    CK(cudaEventRecord(e1, h->stream));
    h->hot_flops += fp64_equiv_flops;  // FP64-equivalent; x 21 digit GEMMs in int8 ops
  }
  return GPC_OK;
}

// V = K* X^T on the INT8 tensor cores for one chunk whose digit image is Aimg (triangular schedule).
// OUT_SUMSQ: sumsq; OUT_F64: Vt [m_pad][n_pad]; OUT_DIGITS: Vimg, the digit image of V (scale v_scale).
template <int OUT>
int launch_vt_i8(gpc_handle h, const int8_t* Aimg, long m_pad, double* Vt, double* sumsq, int8_t* Vimg = nullptr) {
  const long np = h->n_pad;
  VtI8Args a{};
  a.Aimg = Aimg;
  a.Bimg = static_cast<const int8_t*>(h->Bimg.p);
  a.nkb = (int)(np / 64);
  a.nb2 = (int)(np / gpoz::TN);
  a.b_mt = 0;
  a.b_p = gpoz::B_SLICE;
  a.b_k = (long)gpoz::S * gpoz::B_SLICE;
  a.b_j = (long)a.nkb * a.b_k;
  a.sB = h->sBv.d();
  a.sA = kstar_scale(h);
  a.ld_out = np;
  a.m_pad = m_pad;
  a.n_items = (int)(m_pad / gpoz::TM) * ((a.nb2 + 1) / 2);
  a.out = Vt;
  a.dig = Vimg;
  a.nkb_out = a.nkb;
  a.dig_mul = gpoz::DIGIT_MUL / v_scale(h);
  a.sumsq = sumsq;
  a.nlev = i8_levels(h);
  return launch_i8<OUT, false>(h, a, (double)m_pad * (double)np * (double)(np + 64));
}

// Gram[tile] (128 x 128, compact) = V_tile V_tile^T from the digit image of V: the two 64-row halves of the
// tile's own A image are the B operand.  sV1 = device vector of >= 128 copies of v_scale.
int launch_gram_i8(gpc_handle h, const int8_t* Vimg, long m_pad, const double* sV1, double* gram) {
  const long np = h->n_pad;
  VtI8Args a{};
  a.Aimg = Vimg;
  a.Bimg = Vimg;
  a.nkb = (int)(np / 64);
  a.nb2 = 2;
  a.b_p = gpoz::A_SLICE;
  a.b_k = (long)gpoz::S * gpoz::A_SLICE;
  a.b_mt = (long)a.nkb * a.b_k;
  a.b_j = gpoz::B_SLICE;
  a.sB = sV1;
  a.sA = v_scale(h);
  a.ld_out = 128;
  a.m_pad = m_pad;
  a.n_items = (int)(m_pad / gpoz::TM) * 2;
  a.out = gram;
  a.nlev = i8_levels(h);
  return launch_i8<OUT_F64, true>(h, a, (double)m_pad * 128.0 * (double)np);
}

// P [m_pad][ldp] = V Vg^T from the digit images of V (A layout) and of the rows of Vg (B layout, scales sBg).
int launch_cross_i8(gpc_handle h, const int8_t* Vimg, long m_pad, const int8_t* Bg, const double* sBg, long g_pad,
                    double* P) {
  const long np = h->n_pad;
  VtI8Args a{};
  a.Aimg = Vimg;
  a.Bimg = Bg;
  a.nkb = (int)(np / 64);
  a.nb2 = (int)(g_pad / 64);
  a.b_mt = 0;
  a.b_p = gpoz::B_SLICE;
  a.b_k = (long)gpoz::S * gpoz::B_SLICE;
  a.b_j = (long)a.nkb * a.b_k;
  a.sB = sBg;
  a.sA = v_scale(h);
  a.ld_out = g_pad;
  a.m_pad = m_pad;
  a.n_items = (int)(m_pad / gpoz::TM) * a.nb2;
  a.out = P;
  a.nlev = i8_levels(h);
  return launch_i8<OUT_F64, true>(h, a, (double)m_pad * (double)g_pad * (double)np);
}

// Posterior mean + variance of M device-resident test rows on the tcgen05 path, software-pipelined
// over m_chunk-row chunks on two streams: the K* digit assembly of chunk i+1 (FP64 / integer
// pipes, side stream) runs while the INT8 tensor cores contract chunk i (main stream).  The two
// kernels need different resources, so they share the SMs: k_vt_i8 holds one CTA per SM, the
// assembly CTAs fill the registers and shared memory that are left.
int predict_i8_pipeline(gpc_handle h, const double* dXs4, long M, double* dmean, double* dvar, unsigned flags,
                        const double* d_sx, long sx_rows) {
  const long np = h->n_pad, mc = h->m_chunk;
  const long mp_max = round_up(M < mc ? M : mc, 128);
  const int nchunks = (int)((np + KI_COLS - 1) / KI_COLS);
  int rc;
  if ((rc = ensure_slices(h))) return rc;
  DevBuf* Ab[2] = {&h->Aimg, &h->Aimg2};
  DevBuf* Mb[2] = {&h->meanpart, &h->meanpart2};
  DevBuf* Gb[2] = {&h->gradpart, &h->gradpart2};
  const int nbuf = (M > mc) ? 2 : 1;
  for (int b = 0; b < nbuf; ++b) {
    CK(Ab[b]->ensure((size_t)mp_max * np * gpoz::S));
    CK(Mb[b]->ensure((size_t)nchunks * mp_max * 8));
    if (d_sx) CK(Gb[b]->ensure((size_t)nchunks * 3 * mp_max * 8));
  }
  CK(h->sumsq.ensure((size_t)(4 * h->nb) * mp_max * 8));   // k_vt_i8 leaves two partial sums per 64-column tile
  // the assembly runs on the LOW-priority stream: when k_vt_i8 of chunk i and the assembly of chunk i+1 become
  // runnable together, the contraction's persistent CTAs are placed first and the assembly fills in behind them
  cudaStream_t s1 = h->stream, s2 = h->inv;
  CK(cudaEventRecord(h->ev_main, s1));       // the side stream starts after everything queued so far
  CK(cudaStreamWaitEvent(s2, h->ev_main, 0));
  const double sA = kstar_scale(h);
  // profiling only (profiles/tools/phase_power.py): GPC_I8_PHASE=kstar / vt runs one of the two kernels of the pipeline
  // alone (results are then meaningless) so that clocks and board power can be sampled per kernel
  const char* phase_env = std::getenv("GPC_I8_PHASE");
  const bool run_kstar = !phase_env || std::strcmp(phase_env, "vt") != 0;
  const bool run_vt = !phase_env || std::strcmp(phase_env, "kstar") != 0;
  long i = 0;
  for (long o = 0; o < M; o += mc, ++i) {
    const long m = (M - o) < mc ? (M - o) : mc;
    const long m_pad = round_up(m, 128);
    const int b = (int)(i & 1);
    const double* xs = dXs4 + o * 4;
    if (i >= 2) CK(cudaStreamWaitEvent(s2, h->ev_v[b], 0));  // chunk i-2 no longer reads this buffer
    const dim3 grid((unsigned)(m_pad / 128), (unsigned)nchunks);
    if (!run_kstar) {
    } else if (d_sx)
      k_kstar_i8<true><<<grid, 128, 0, s2>>>(h->hyp, h->Xt.d(), h->alpha.d(), h->N, np, xs, m, m_pad, sA,
                                             static_cast<int8_t*>(Ab[b]->p), Mb[b]->d(), Gb[b]->d());
    else
      k_kstar_i8<false><<<grid, 128, 0, s2>>>(h->hyp, h->Xt.d(), h->alpha.d(), h->N, np, xs, m, m_pad, sA,
                                              static_cast<int8_t*>(Ab[b]->p), Mb[b]->d(), nullptr);
    CKL();
    CK(cudaEventRecord(h->ev_k[b], s2));
    CK(cudaStreamWaitEvent(s1, h->ev_k[b], 0));
    if (run_vt && (rc = launch_vt_i8<OUT_SUMSQ>(h, static_cast<const int8_t*>(Ab[b]->p), m_pad, nullptr, h->sumsq.d()))) return rc;
    k_finalize_pred<<<(unsigned)((m + 255) / 256), 256, 0, s1>>>(
        h->hyp, xs, m, m_pad, Mb[b]->d(), nchunks, h->sumsq.d(), 4 * h->nb, Gb[b]->d(),
        d_sx ? d_sx + (sx_rows == 1 ? 0 : o * 3) : nullptr, sx_rows, dmean ? dmean + o : nullptr, dvar + o, flags);
    CKL();
    CK(cudaEventRecord(h->ev_v[b], s1));
  }
  return GPC_OK;
}

inline bool use_i8(gpc_handle h, const double* dvar, unsigned flags) {
  return dvar && !(flags & GPC_MEAN_ONLY) && h->mode != GPC_MODE_FP64 && h->n_pad <= gpoz::MAX_K;
}

// One chunk of the posterior on the FP64 path (device pointers; M <= m_chunk).  d_sx != NULL adds
// the NIGP test-input-noise term (needs the mean gradients: the gradient variant of k_kstar runs).
int predict_chunk(gpc_handle h, const double* dXs4, long M, double* dmean, double* dvar, unsigned flags,
                  const double* d_sx = nullptr, long sx_rows = 0) {
  const long m_pad = round_up(M, 128);
  int rc;
  const bool want_var = dvar && !(flags & GPC_MEAN_ONLY);
  const int nchunks = (int)((h->n_pad + KS_COLS - 1) / KS_COLS);
  if ((rc = ensure_pred_ws(h, m_pad, false, d_sx != nullptr))) return rc;
  if (want_var) {
    if (d_sx) rc = launch_kstar<true, true>(h, dXs4, M, m_pad);
    else rc = launch_kstar<false, true>(h, dXs4, M, m_pad);
    if (rc) return rc;
    if ((rc = launch_vt<false, true>(h, m_pad))) return rc;
  } else {
    if ((rc = launch_kstar<false, false>(h, dXs4, M, m_pad))) return rc;
  }
  k_finalize_pred<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(
      h->hyp, dXs4, M, m_pad, h->meanpart.d(), nchunks, h->sumsq.d(), 2 * h->nb, h->gradpart.d(),
      want_var ? d_sx : nullptr, sx_rows, dmean, want_var ? dvar : nullptr, flags);
  CKL();
  return GPC_OK;
}

// All chunks of a device-resident test set.
int predict_rows(gpc_handle h, const double* dXs4, long M, double* dmean, double* dvar, unsigned flags,
                 const double* d_sx, long sx_rows) {
  if (use_i8(h, dvar, flags)) return predict_i8_pipeline(h, dXs4, M, dmean, dvar, flags, d_sx, sx_rows);
  int rc;
  for (long o = 0; o < M; o += h->m_chunk) {
    const long m = (M - o) < h->m_chunk ? (M - o) : h->m_chunk;
    if ((rc = predict_chunk(h, dXs4 + o * 4, m, dmean ? dmean + o : nullptr, dvar ? dvar + o : nullptr, flags,
                            d_sx ? d_sx + (sx_rows == 1 ? 0 : o * 3) : nullptr, sx_rows)))
      return rc;
  }
  return GPC_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int gpc_version(void) { return GPC_VERSION; }

const char* gpc_last_error(gpc_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static void destroy_handle(gpc_handle h);

int gpc_create(int kind, int F, int device, gpc_handle* out) {
  gpc_handle h = nullptr;
  if (!out) return fail(nullptr, GPC_ERR_ARG, "out == NULL");
  *out = nullptr;
  if (kind < GPC_SF_RBF || kind > GPC_NIGP) return fail(nullptr, GPC_ERR_ARG, "unknown model kind");
  const bool mf = (kind == GPC_MF_AR1_RBF || kind == GPC_MF_AR1_MAT32);
  if (F < 1 || F > GPC_MAXF || (!mf && F != 1)) return fail(nullptr, GPC_ERR_SHAPE, "bad fidelity count");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, GPC_ERR_CUDA,
                std::string("no CUDA device (gpcore has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, GPC_ERR_ARG, "bad device ordinal");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, GPC_ERR_CUDA, cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10)
    return fail(nullptr, GPC_ERR_CUDA, "gpcore is built for sm_100a (Blackwell B200) only");
  h = new gpc_handle_s();
  h->kind = kind;
  h->F = F;
  h->device = device;
  // the side stream carries the latency-critical chain of the factorisation (diagonal blocks, panel solves):
  // it outranks the main stream (bulk updates), which outranks the early-inverse stream
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const int prio_mid = (prio_hi < prio_lo - 1) ? prio_hi + 1 : prio_lo;
  e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_mid);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->inv, cudaStreamNonBlocking, prio_lo);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_cmp[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_half, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_inv, cudaEventDisableTiming);

  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_v[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    std::string m = cudaGetErrorString(e);
    destroy_handle(h);            // streams / events created so far
    return fail(nullptr, GPC_ERR_CUDA, m);
  }
  int rc = set_gemm_attrs(h);
  if (rc) {
    g_create_error = h->err;
    destroy_handle(h);
    return rc;
  }
  *out = h;
  return GPC_OK;
}

// Releases everything a handle owns; safe on a partially constructed handle (every member starts as nullptr).
static void destroy_handle(gpc_handle h) {
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = {&h->gmT, &h->gmA, &h->gmC, &h->gmc, &h->gmax, &h->gmmean, &h->Vimg, &h->gBimg, &h->gsB, &h->sV1, &h->Pt, &h->Xt, &h->y, &h->extra, &h->L, &h->X, &h->T, &h->alpha, &h->vec, &h->partial, &h->scal,
                    &h->status, &h->Wm, &h->gpart, &h->Xs4, &h->Kx, &h->meanpart, &h->sumsq, &h->gradpart, &h->mean, &h->var, &h->Vt,
                    &h->cov, &h->grads, &h->ediag, &h->Bimg, &h->sBv, &h->Aimg, &h->Aimg2, &h->meanpart2, &h->gradpart2, &h->gX4, &h->gVt, &h->gS, &h->gSinv, &h->gT, &h->Bt, &h->Zt,
                    &h->cand_off, &h->cand_I, &h->cand_aux, &h->cand_rows, &h->cand_mask, &h->gram, &h->gramZ};
  for (DevBuf* b : bufs) b->release();
  h->pin_in.release(); h->pin_sx.release(); h->pin_mean.release(); h->pin_var.release();
  for (FactorGraph& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
    if (h->ev_v[i]) cudaEventDestroy(h->ev_v[i]);
  }
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->ev_side) cudaEventDestroy(h->ev_side);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_cmp[i]) cudaEventDestroy(h->ev_cmp[i]);
    if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
  }
  if (h->ev_call0) cudaEventDestroy(h->ev_call0);
  if (h->ev_call1) cudaEventDestroy(h->ev_call1);
  if (h->ev_half) cudaEventDestroy(h->ev_half);
  if (h->ev_inv) cudaEventDestroy(h->ev_inv);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->inv) cudaStreamDestroy(h->inv);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int gpc_destroy(gpc_handle h) {
  if (!h) return GPC_OK;
  destroy_handle(h);
  return GPC_OK;
}

int gpc_set_hypers(gpc_handle h, const double* p, int n, double jitter) {
  if (!h || !p) return GPC_ERR_ARG;
  GpcHyp g;
  memset(&g, 0, sizeof(g));
  g.F = h->F;
  g.jitter = jitter;
  const int F = h->F;
  auto bad = [&](const char* m) { return fail(h, GPC_ERR_SHAPE, m); };
  if (h->kind == GPC_SF_RBF || h->kind == GPC_SF_MAT32) {
    if (n != 5) return bad("single-fidelity hypers: expected [variance, lx, ly, lz, noise_var]");
    g.base = (h->kind == GPC_SF_MAT32);
    g.var[0] = p[0];
    for (int d = 0; d < 3; ++d) g.inv_l[0][d] = 1.0 / p[1 + d];
    g.noise[0] = p[4];
    g.coef[0][0] = 1.0;
  } else if (h->kind == GPC_NIGP) {
    if (n != 5) return bad("NIGP hypers: expected [lx, ly, lz, sigma_f, sigma_y]");
    g.base = 0;
    for (int d = 0; d < 3; ++d) g.inv_l[0][d] = 1.0 / sqrt(1.0 / (1.0 / (p[d] * p[d])));  // inv_l=True round trip
    h->sigma_y = p[4];
    g.var[0] = p[3];           // NIGP.py:18 passes sigma_f as the kernel variance
    g.noise[0] = p[4] * p[4];  // sigma_y is a standard deviation (NIGP.py:41)
    g.coef[0][0] = 1.0;
  } else {
    const int n1 = 4 * F + (F - 1) + 1, nF = 4 * F + (F - 1) + F;
    if (n != n1 && n != nF) return bad("multi-fidelity hypers: expected 4F + (F-1) + (1|F) values");
    g.base = (h->kind == GPC_MF_AR1_MAT32);
    for (int m = 0; m < F; ++m) {
      g.var[m] = p[4 * m];
      for (int d = 0; d < 3; ++d) g.inv_l[m][d] = 1.0 / p[4 * m + 1 + d];
    }
    const double* rho = p + 4 * F;
    for (int l = 0; l < F - 1; ++l) h->rho[l] = rho[l];
    h->n_noise = (n == n1) ? 1 : F;
    for (int i = 0; i < F; ++i)
      for (int m = 0; m <= i; ++m) {
        double c = 1.0;
        for (int l = m; l < i; ++l) c *= rho[l];
        g.coef[i][m] = c;
      }
    const double* nz = p + 4 * F + F - 1;
    for (int i = 0; i < F; ++i) g.noise[i] = (n == n1) ? nz[0] : nz[i];
  }
  for (int i = 0; i < F; ++i) {
    double v = 0.0;
    for (int m = 0; m <= i; ++m) v += g.coef[i][m] * g.coef[i][m] * g.var[m];
    g.kdiag[i] = v;
  }
  for (int m = 0; m < F; ++m) {
    if (!(g.var[m] >= 0.0) || !std::isfinite(g.var[m])) return bad("kernel variance must be finite and >= 0");
    for (int d = 0; d < 3; ++d)
      if (!std::isfinite(g.inv_l[m][d])) return bad("lengthscales must be non-zero and finite");
  }
  h->hyp = g;
  h->have_hyp = true;
  h->factored = false;
  return GPC_OK;
}

int gpc_set_data(gpc_handle h, const double* X4, const double* y, const double* extra, long N) {
  if (!h || !X4 || !y) return GPC_ERR_ARG;
  if (N < 1) return fail(h, GPC_ERR_SHAPE, "N must be >= 1");
  CK(cudaSetDevice(h->device));
  const long np = round_up(N, 128);
  // Multi-fidelity models keep their training rows sorted by fidelity (stable), so that a block of
  // consecutive columns of K* shares one fidelity and the AR1 sum needs no per-lane branching.
  // Every posterior quantity is invariant under this permutation; gpc_get_alpha undoes it.
  h->perm.resize((size_t)N);
  for (long i = 0; i < N; ++i) h->perm[(size_t)i] = i;
  if (h->F > 1) {
    for (long i = 0; i < N; ++i) {
      const double f = X4[i * 4 + 3];
      if (!(f >= 0.0) || f >= (double)h->F || f != std::floor(f))
        return fail(h, GPC_ERR_ARG, "fidelity index must be an integer in [0, F)");
    }
    std::stable_sort(h->perm.begin(), h->perm.end(),
                     [&](long a, long b) { return X4[a * 4 + 3] < X4[b * 4 + 3]; });
  }
  std::vector<double> soa((size_t)4 * np, 0.0), yp((size_t)np, 0.0);
  for (long i = 0; i < N; ++i) {
    const long src = h->perm[(size_t)i];
    for (int c = 0; c < 4; ++c) soa[(size_t)c * np + i] = X4[src * 4 + c];
    yp[i] = y[src];
  }
  CK(h->Xt.ensure((size_t)4 * np * 8));
  CK(h->y.ensure((size_t)np * 8));
  CK(h->alpha.ensure((size_t)np * 8));
  CK(h->vec.ensure((size_t)3 * np * 8));
  CK(h->scal.ensure(64));
  CK(h->status.ensure(64));
  CK(cudaMemcpyAsync(h->Xt.p, soa.data(), (size_t)4 * np * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->y.p, yp.data(), (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
  h->have_extra = extra != nullptr;
  std::vector<double> ep;
  if (extra) {
    ep.assign((size_t)np, 0.0);
    for (long i = 0; i < N; ++i) ep[(size_t)i] = extra[h->perm[(size_t)i]];
    CK(h->extra.ensure((size_t)np * 8));
    CK(cudaMemcpyAsync(h->extra.p, ep.data(), (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  h->N = N;
  h->n_pad = np;
  h->nb = (int)(np / 128);
  h->have_data = true;
  h->factored = false;
  return GPC_OK;
}

int gpc_factor(gpc_handle h, double* nlml, double* logdet) {
  if (!h) return GPC_ERR_ARG;
  if (!h->have_hyp || !h->have_data) return fail(h, GPC_ERR_STATE, "set hypers and data before gpc_factor");
  CK(cudaSetDevice(h->device));
  const long np = h->n_pad;
  const int nb = h->nb;
  CK(h->L.ensure((size_t)np * np * 8));
  CK(h->X.ensure((size_t)np * np * 8));
  CK(h->T.ensure((size_t)np * np * 8));
  CK(h->partial.ensure((size_t)nb * np * 8));
  h->factored = false;
  k_assemble_train<<<dim3(nb, nb), 256, 0, h->stream>>>(h->hyp, h->Xt.d(), h->have_extra ? h->extra.d() : nullptr,
                                                        h->L.d(), h->N, np);
  CKL();
  int rc = factor_matrix(h, h->L.d(), h->X.d(), h->T.d(), np, static_cast<int*>(h->status.p));
  if (rc) return rc;
  // alpha and the log-det are queued behind the factorisation without waiting for its status (a non-PD matrix only
  // makes them garbage, which is never returned): ONE host synchronisation per gpc_factor
  if ((rc = solve_alpha(h))) return rc;
  k_logdet_fit<<<1, 1024, 0, h->stream>>>(h->L.d(), np, h->N, h->y.d(), h->alpha.d(), h->scal.d());
  CKL();
  int st = 0;
  double sc[2];
  CK(cudaMemcpyAsync(&st, h->status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(sc, h->scal.p, 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (st != 0) {
    char buf[128];
    snprintf(buf, sizeof buf, "covariance not positive definite (pivot block %d)", st - 1);
    return fail(h, GPC_ERR_NOT_PD, buf);
  }
  if (!std::isfinite(sc[0]) || !std::isfinite(sc[1]))
    return fail(h, GPC_ERR_NOT_PD, "non-finite factor (covariance not positive definite)");
  h->logdet = sc[0];
  h->nlml = 0.5 * sc[1] + 0.5 * sc[0] + 0.5 * (double)h->N * log(2.0 * M_PI);
  h->factored = true;
  h->have_slices = false;
  if (nlml) *nlml = h->nlml;
  if (logdet) *logdet = h->logdet;
  return GPC_OK;
}

int gpc_nlml_grad(gpc_handle h, double* grad, int n, double* diagW) {
  int rc = require_factor(h);
  if (rc) return rc;
  if (!grad) return GPC_ERR_ARG;
  const int F = h->F;
  const bool mf = (h->kind == GPC_MF_AR1_RBF || h->kind == GPC_MF_AR1_MAT32);
  const int n_noise = mf ? h->n_noise : 1;
  const int want = mf ? 4 * F + (F - 1) + n_noise : 5;
  if (n != want) return fail(h, GPC_ERR_SHAPE, "gpc_nlml_grad: gradient length must match the hyper-parameter vector");
  CK(cudaSetDevice(h->device));
  const long np = h->n_pad;
  const int nb = h->nb;
  cudaStream_t s = h->stream;
  CK(h->Wm.ensure((size_t)np * np * 8));
  CK(h->T.ensure((size_t)np * np * 8));
  CK(h->gpart.ensure(((size_t)nb * nb * GPC_NGK + GPC_NGK + (size_t)np) * 8));
  GpcGradTab tab;
  memset(&tab, 0, sizeof(tab));
  for (int l = 0; l < F - 1; ++l)
    for (int i = 0; i < F; ++i)
      for (int m = 0; m <= i; ++m)
        if (m <= l && l < i) {
          double c = 1.0;
          for (int q = m; q < i; ++q)
            if (q != l) c *= h->rho[q];
          tab.dcoef[l][i][m] = c;
        }
  k_transpose<<<dim3((unsigned)(np / 32), (unsigned)(np / 32)), 256, 0, s>>>(h->X.d(), h->T.d(), np);
  CKL();
  k_kinv<<<dim3(nb, nb), gpcg::NTHREADS, gpcg::SMEM_BYTES, s>>>(h->T.d(), np, nb, h->Wm.d());
  CKL();
  double* part = h->gpart.d();
  double* gsum = part + (size_t)nb * nb * GPC_NGK;
  double* dW = gsum + GPC_NGK;
  k_nlml_grad<<<dim3(nb, nb), 256, 0, s>>>(h->hyp, tab, h->Xt.d(), h->alpha.d(), h->Wm.d(), h->N, np, nb, part, dW);
  CKL();
  k_reduce_partial<<<GPC_NGK, 256, 0, s>>>(part, (long)nb * nb, gsum);
  CKL();
  double gk[GPC_NGK];
  std::vector<double> dw((size_t)h->N), xf;
  CK(cudaMemcpyAsync(gk, gsum, sizeof(gk), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(dw.data(), dW, (size_t)h->N * 8, cudaMemcpyDeviceToHost, s));
  if (mf) {
    xf.resize((size_t)h->N);
    CK(cudaMemcpyAsync(xf.data(), h->Xt.d() + 3 * np, (size_t)h->N * 8, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  // d/d noise_f = 1/2 sum_{i at fidelity f} W_ii
  double tr_f[GPC_MAXF] = {0, 0, 0, 0};
  for (long i = 0; i < h->N; ++i) tr_f[mf ? (int)xf[(size_t)i] : 0] += 0.5 * dw[(size_t)i];
  if (mf) {
    for (int m = 0; m < F; ++m)
      for (int q = 0; q < 4; ++q) grad[4 * m + q] = gk[4 * m + q];
    for (int l = 0; l < F - 1; ++l) grad[4 * F + l] = gk[4 * GPC_MAXF + l];
    if (n_noise == 1) {
      double t = 0.0;
      for (int f = 0; f < F; ++f) t += tr_f[f];
      grad[4 * F + F - 1] = t;
    } else {
      for (int f = 0; f < F; ++f) grad[4 * F + F - 1 + f] = tr_f[f];
    }
  } else if (h->kind == GPC_NIGP) {
    for (int d = 0; d < 3; ++d) grad[d] = gk[1 + d];
    grad[3] = gk[0];                                 // sigma_f is the kernel variance
    grad[4] = 2.0 * h->sigma_y * tr_f[0];            // d/d sigma_y of sigma_y^2 on the diagonal
  } else {
    for (int q = 0; q < 4; ++q) grad[q] = gk[q];
    grad[4] = tr_f[0];
  }
  if (diagW)
    for (long i = 0; i < h->N; ++i) diagW[h->perm[(size_t)i]] = dw[(size_t)i];
  return GPC_OK;
}

long gpc_padded_n(gpc_handle h) { return h ? h->n_pad : 0; }

int gpc_get_alpha(gpc_handle h, double* alpha) {
  int rc = require_factor(h);
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  std::vector<double> tmp((size_t)h->N);
  CK(cudaMemcpyAsync(tmp.data(), h->alpha.p, (size_t)h->N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (long i = 0; i < h->N; ++i) alpha[h->perm[(size_t)i]] = tmp[(size_t)i];  // back to the caller's row order
  return GPC_OK;
}

static int get_lower(gpc_handle h, const double* dsrc, double* out) {
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy2DAsync(out, (size_t)h->N * 8, dsrc, (size_t)h->n_pad * 8, (size_t)h->N * 8, (size_t)h->N,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (long i = 0; i < h->N; ++i)
    for (long j = i + 1; j < h->N; ++j) out[i * h->N + j] = 0.0;
  return GPC_OK;
}

int gpc_get_chol(gpc_handle h, double* L) {
  int rc = require_factor(h);
  return rc ? rc : get_lower(h, h->L.d(), L);
}

int gpc_get_linv(gpc_handle h, double* Linv) {
  int rc = require_factor(h);
  return rc ? rc : get_lower(h, h->X.d(), Linv);
}

int gpc_factor_state_dev(gpc_handle h, double** L, double** Linv, double** alpha, long* n_pad) {
  if (!h) return GPC_ERR_ARG;
  if (!h->have_hyp || !h->have_data) return fail(h, GPC_ERR_STATE, "set hypers and data first");
  CK(cudaSetDevice(h->device));
  const long np = h->n_pad;
  CK(h->L.ensure((size_t)np * np * 8));
  CK(h->X.ensure((size_t)np * np * 8));
  CK(h->partial.ensure((size_t)h->nb * np * 8));
  if (L) *L = h->L.d();
  if (Linv) *Linv = h->X.d();
  if (alpha) *alpha = h->alpha.d();
  if (n_pad) *n_pad = np;
  return GPC_OK;
}

int gpc_adopt_factor(gpc_handle h, double logdet) {
  if (!h) return GPC_ERR_ARG;
  if (!h->have_hyp || !h->have_data || !h->L.p || !h->X.p)
    return fail(h, GPC_ERR_STATE, "gpc_adopt_factor needs hypers, data and gpc_factor_state_dev buffers");
  h->logdet = logdet;
  h->factored = true;
  h->have_slices = false;
  return GPC_OK;
}

int gpc_kernel_matrix(gpc_handle h, const double* Xa4, long na, const double* Xb4, long nb_, double* K) {
  if (!h || !Xa4 || !K) return GPC_ERR_ARG;
  if (!h->have_hyp) return fail(h, GPC_ERR_STATE, "set hypers first");
  if (!Xb4) { Xb4 = Xa4; nb_ = na; }
  if (na < 1 || nb_ < 1) return fail(h, GPC_ERR_SHAPE, "empty input");
  if (check_fidelity_rows(h, Xa4, na, "kernel rows") || check_fidelity_rows(h, Xb4, nb_, "kernel rows")) return GPC_ERR_ARG;
  CK(cudaSetDevice(h->device));
  DevBuf a, b, k;
  cudaError_t e;
  int rc = GPC_OK;
  if ((e = a.ensure((size_t)na * 32)) || (e = b.ensure((size_t)nb_ * 32)) || (e = k.ensure((size_t)na * nb_ * 8))) {
    rc = fail(h, GPC_ERR_CUDA, cudaGetErrorString(e));
  } else {
    cudaMemcpyAsync(a.p, Xa4, (size_t)na * 32, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(b.p, Xb4, (size_t)nb_ * 32, cudaMemcpyHostToDevice, h->stream);
    k_kernel_matrix<<<dim3((unsigned)((nb_ + 31) / 32), (unsigned)((na + 7) / 8)), 256, 0, h->stream>>>(
        h->hyp, a.d(), na, b.d(), nb_, k.d());
    ++h->launches;
    cudaMemcpyAsync(K, k.p, (size_t)na * nb_ * 8, cudaMemcpyDeviceToHost, h->stream);
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(h, GPC_ERR_CUDA, cudaGetErrorString(e));
  }
  a.release(); b.release(); k.release();
  return rc;
}

int gpc_predict_dev(gpc_handle h, const double* dXs4, long M, double* dmean, double* dvar, unsigned flags) {
  int rc = require_factor(h);
  if (rc) return rc;
  if (M < 0 || !dXs4) return fail(h, GPC_ERR_SHAPE, "bad test set");
  if (M == 0) return GPC_OK;
  CK(cudaSetDevice(h->device));
  return predict_rows(h, dXs4, M, dmean, dvar, flags, nullptr, 0);
}

// Host-pointer posterior.  The test rows travel in stages of up to GPC_STAGE_ROWS rows; two device stages alternate.
// The copy stream (in order) carries  in(0), in(1), out(0), in(2), out(1), ...: the host -> device copy of stage s + 1
// and the device -> host copy of stage s - 1 run while the compute stream works on stage s.
//
// Caller arrays are usually PAGEABLE (plain NumPy arrays): a cudaMemcpyAsync on pageable memory blocks the host (for a
// device -> host copy until the data has arrived), which would serialise the pipeline.  So pageable arrays are staged
// through the handle's own page-locked buffers: the host thread memcpy's stage s + 1 into its pinned input buffer and
// drains the pinned results of stage s - 1 into the caller's arrays while the GPU computes stage s; every copy the GPU
// sees is a plain cudaMemcpyAsync on pinned memory.  Arrays the caller has pinned itself (cudaHostAlloc /
// cudaHostRegister, detected with cudaPointerGetAttributes) are used in place.
#define GPC_STAGE_ROWS (1L << 17)
static bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

static int predict_host(gpc_handle h, const double* Xs4, long M, const double* sx, long sx_rows, double* mean,
                        double* var, unsigned flags) {
  int rc = require_factor(h);
  if (rc) return rc;
  if (M < 0 || (M > 0 && !Xs4)) return fail(h, GPC_ERR_SHAPE, "bad test set");
  if (sx && sx_rows != 1 && sx_rows != M) return fail(h, GPC_ERR_SHAPE, "input-noise array must be (1 x 3) or (M x 3)");
  if (sx && (h->F != 1 || h->hyp.base != 0))
    return fail(h, GPC_ERR_ARG, "input-noise correction is defined for the single-fidelity squared-exponential kernel");
  if (M == 0) return GPC_OK;
  CK(cudaSetDevice(h->device));
  const long stage = M <= GPC_STAGE_ROWS ? round_up(M, 128) : GPC_STAGE_ROWS;
  const int nbuf = M > stage ? 2 : 1;
  CK(h->Xs4.ensure((size_t)nbuf * stage * 32));
  CK(h->mean.ensure((size_t)nbuf * stage * 8));
  CK(h->var.ensure((size_t)nbuf * stage * 8));
  cudaStream_t sc = h->stream, sio = h->side;
  const bool per_row_sx = sx && sx_rows != 1;
  const bool want_var = var && !(flags & GPC_MEAN_ONLY);
  const bool stage_in = !is_pinned(Xs4) || (per_row_sx && !is_pinned(sx));
  const bool stage_out = !is_pinned(mean) || (want_var && !is_pinned(var));
  if (stage_in) {
    CK(h->pin_in.ensure((size_t)nbuf * stage * 32));
    if (per_row_sx) CK(h->pin_sx.ensure((size_t)nbuf * stage * 24));
  }
  if (stage_out) {
    if (mean) CK(h->pin_mean.ensure((size_t)nbuf * stage * 8));
    if (want_var) CK(h->pin_var.ensure((size_t)nbuf * stage * 8));
  }
  if (sx) {
    CK(h->ediag.ensure((size_t)(per_row_sx ? nbuf * stage : 1) * 24));
    if (!per_row_sx) CK(cudaMemcpyAsync(h->ediag.p, sx, 24, cudaMemcpyHostToDevice, sc));
  }
  int* d_badfid = static_cast<int*>(h->status.p) + 4;
  CK(cudaMemsetAsync(d_badfid, 0, sizeof(int), sc));
  CK(cudaEventRecord(h->ev_cmp[0], sc));            // the copy stream starts behind everything queued so far
  CK(cudaStreamWaitEvent(sio, h->ev_cmp[0], 0));
  const long nstage = (M + stage - 1) / stage;
  // any error below leaves copies into caller-owned memory in flight: quiesce both streams before returning
  auto bail = [&](int code) {
    cudaStreamSynchronize(sio);
    cudaStreamSynchronize(sc);
    return code;
  };
#define CKB(call)                                                                                            \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess) return bail(fail(h, GPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__))); \
  } while (0)
  auto copy_in = [&](long si) -> int {
    const long s0 = si * stage, ms = (M - s0) < stage ? (M - s0) : stage;
    const int b = (int)(si % nbuf);
    const double* src = Xs4 + s0 * 4;
    const double* srcx = per_row_sx ? sx + s0 * 3 : nullptr;
    if (stage_in) {
      // the pinned buffer last fed stage si - 2: that copy completed long ago, but make it explicit
      if (si >= 2) CKB(cudaEventSynchronize(h->ev_h2d[b]));
      double* pin = h->pin_in.d() + (size_t)b * stage * 4;
      std::memcpy(pin, src, (size_t)ms * 32);
      src = pin;
      if (per_row_sx) {
        double* pinx = h->pin_sx.d() + (size_t)b * stage * 3;
        std::memcpy(pinx, srcx, (size_t)ms * 24);
        srcx = pinx;
      }
    }
    CKB(cudaMemcpyAsync(h->Xs4.d() + (size_t)b * stage * 4, src, (size_t)ms * 32, cudaMemcpyHostToDevice, sio));
    if (per_row_sx)
      CKB(cudaMemcpyAsync(h->ediag.d() + (size_t)b * stage * 3, srcx, (size_t)ms * 24, cudaMemcpyHostToDevice, sio));
    CKB(cudaEventRecord(h->ev_h2d[b], sio));
    return GPC_OK;
  };
  // results of stage si: pinned bounce buffer -> caller arrays, once its device -> host copy has completed
  auto drain = [&](long si) -> int {
    if (!stage_out) return GPC_OK;
    const long s0 = si * stage, ms = (M - s0) < stage ? (M - s0) : stage;
    const int b = (int)(si % nbuf);
    CKB(cudaEventSynchronize(h->ev_d2h[b]));
    if (mean) std::memcpy(mean + s0, h->pin_mean.d() + (size_t)b * stage, (size_t)ms * 8);
    if (want_var) std::memcpy(var + s0, h->pin_var.d() + (size_t)b * stage, (size_t)ms * 8);
    return GPC_OK;
  };
  if ((rc = copy_in(0))) return rc;
  for (long si = 0; si < nstage; ++si) {
    const long s0 = si * stage, ms = (M - s0) < stage ? (M - s0) : stage;
    const int b = (int)(si % nbuf);
    double* dX = h->Xs4.d() + (size_t)b * stage * 4;
    double* dm = h->mean.d() + (size_t)b * stage;
    double* dv = h->var.d() + (size_t)b * stage;
    const double* de = sx ? (per_row_sx ? h->ediag.d() + (size_t)b * stage * 3 : h->ediag.d()) : nullptr;
    CKB(cudaStreamWaitEvent(sc, h->ev_h2d[b], 0));                       // inputs of this stage have landed
    if (si >= 2) CKB(cudaStreamWaitEvent(sc, h->ev_d2h[b], 0));          // results of stage s - 2 have left these buffers
    if (h->F > 1) {
      k_check_fid<<<(unsigned)((ms + 255) / 256), 256, 0, sc>>>(dX, ms, h->F, d_badfid);
      ++h->launches;
    }
    if ((rc = predict_rows(h, dX, ms, mean ? dm : nullptr, want_var ? dv : nullptr, flags, de, sx_rows))) return bail(rc);
    CKB(cudaEventRecord(h->ev_cmp[b], sc));
    if (si + 1 < nstage && (rc = copy_in(si + 1))) return rc;           // queued BEFORE out(s): it must not wait for compute(s)
    CKB(cudaStreamWaitEvent(sio, h->ev_cmp[b], 0));
    double* om = stage_out ? h->pin_mean.d() + (size_t)b * stage : mean + s0;
    double* ov = stage_out ? h->pin_var.d() + (size_t)b * stage : var + s0;
    if (mean) CKB(cudaMemcpyAsync(om, dm, (size_t)ms * 8, cudaMemcpyDeviceToHost, sio));
    if (want_var) CKB(cudaMemcpyAsync(ov, dv, (size_t)ms * 8, cudaMemcpyDeviceToHost, sio));
    CKB(cudaEventRecord(h->ev_d2h[b], sio));
    if (si >= 1 && (rc = drain(si - 1))) return rc;                     // while the GPU computes stage si
  }
  if ((rc = drain(nstage - 1))) return rc;
  int badfid = 0;
  CKB(cudaMemcpyAsync(&badfid, d_badfid, sizeof(int), cudaMemcpyDeviceToHost, sc));
  CKB(cudaStreamSynchronize(sio));
  CKB(cudaStreamSynchronize(sc));
#undef CKB
  if (badfid) return fail(h, GPC_ERR_ARG, "test rows: fidelity index must be an integer in [0, F)");
  return GPC_OK;
}

int gpc_predict(gpc_handle h, const double* Xs4, long M, double* mean, double* var, unsigned flags) {
  return predict_host(h, Xs4, M, nullptr, 0, mean, var, flags);
}

int gpc_predict_noisy(gpc_handle h, const double* Xs4, long M, const double* sx, long sx_rows, double* mean,
                      double* var, unsigned flags) {
  if (!sx) return fail(h, GPC_ERR_ARG, "sx == NULL");
  return predict_host(h, Xs4, M, sx, sx_rows, mean, var, flags);
}

// dmean != NULL: device output (asynchronous on the handle's stream), else the result is copied to `mean` (host).
static int grid_mean_impl(gpc_handle h, const double* ax, long nx, const double* ay, long ny, const double* az, long nz,
                          double fid, double* mean, double* dmean) {
  int rc = require_factor(h);
  if (rc) return rc;
  if (!ax || !ay || !az || (!mean && !dmean) || nx < 1 || ny < 1 || nz < 1) return fail(h, GPC_ERR_SHAPE, "bad grid axes");
  if (h->hyp.base != 0) return fail(h, GPC_ERR_ARG, "the tensor-grid mean needs a squared-exponential kernel (it separates per axis)");
  if (h->F > 1 && ((int)fid < 0 || (int)fid >= h->F)) return fail(h, GPC_ERR_ARG, "fidelity index out of range");
  if (nx > 65535 || ny > 65535 || nz > 65472) return fail(h, GPC_ERR_SHAPE, "an axis has more than 65535 points");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const long np = h->n_pad, npairs = nx * ny;
  const long nzp = round_up(nz, 64), nxp = nx, nyp = ny;
  const long R = std::min<long>(round_up(npairs, 64), 8192);          // (ix, iy) pairs per pass
  const int fi = h->F > 1 ? (int)fid : 0;
  CK(h->gmax.ensure((size_t)(nx + ny + nz) * 8));
  CK(h->gmT.ensure((size_t)(nxp + nyp + nzp) * np * 8));
  CK(h->gmA.ensure((size_t)R * np * 8));
  CK(h->gmC.ensure((size_t)R * nzp * 8));
  CK(h->gmc.ensure((size_t)np * 8));
  if (!dmean) {
    CK(h->gmmean.ensure((size_t)npairs * nz * 8));
    dmean = h->gmmean.d();
  }
  double* dax = h->gmax.d();
  double* day = dax + nx;
  double* daz = day + ny;
  CK(cudaMemcpyAsync(dax, ax, (size_t)nx * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(day, ay, (size_t)ny * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(daz, az, (size_t)nz * 8, cudaMemcpyHostToDevice, s));
  double* Tx = h->gmT.d();
  double* Ty = Tx + nxp * np;
  double* Tz = Ty + nyp * np;
  const int mtop = fi < h->F - 1 ? fi : h->F - 1;
  const unsigned nb256 = (unsigned)((np + 255) / 256);
  for (long p0 = 0; p0 < npairs; p0 += R) {
    for (int m = 0; m <= mtop; ++m) {
      if (p0 == 0 || mtop > 0) {   // one term: the tables are built once; several terms: per pass and term
        k_gm_table<<<dim3(nb256, (unsigned)nxp), 256, 0, s>>>(h->hyp, h->Xt.d(), np, h->N, dax, (int)nx, 0, m, Tx);
        k_gm_table<<<dim3(nb256, (unsigned)nyp), 256, 0, s>>>(h->hyp, h->Xt.d(), np, h->N, day, (int)ny, 1, m, Ty);
        k_gm_table<<<dim3(nb256, (unsigned)nzp), 256, 0, s>>>(h->hyp, h->Xt.d(), np, h->N, daz, (int)nz, 2, m, Tz);
        k_gm_coef<<<nb256, 256, 0, s>>>(h->hyp, h->Xt.d(), h->alpha.d(), np, h->N, fi, m, h->gmc.d());
        CKL();
        h->launches += 3;
      }
      k_gm_form<<<dim3(nb256, (unsigned)R), 256, 0, s>>>(h->gmc.d(), Tx, Ty, np, p0, npairs, (int)ny, h->gmA.d());
      CKL();
      k_gm_gemm<<<dim3((unsigned)(nzp / 64), (unsigned)(R / 64)), gpc64::NT, gpc64::SMEM_BYTES, s>>>(
          h->gmA.d(), Tz, np, h->gmC.d(), nzp, m == 0 ? 0.0 : 1.0);
      CKL();
    }
    k_gm_store<<<(unsigned)(4 * h->n_sm), 256, 0, s>>>(h->gmC.d(), nzp, p0, npairs, (int)nz, R, dmean);
    CKL();
  }
  if (mean) {
    CK(cudaMemcpyAsync(mean, dmean, (size_t)npairs * nz * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return GPC_OK;
}

int gpc_predict_grid_mean(gpc_handle h, const double* ax, long nx, const double* ay, long ny, const double* az, long nz,
                          double fid, double* mean) {
  if (!mean) return fail(h, GPC_ERR_SHAPE, "mean is NULL");
  return grid_mean_impl(h, ax, nx, ay, ny, az, nz, fid, mean, nullptr);
}

int gpc_predict_grid_mean_dev(gpc_handle h, const double* ax, long nx, const double* ay, long ny, const double* az,
                              long nz, double fid, double* dmean) {
  if (!dmean) return fail(h, GPC_ERR_SHAPE, "dmean is NULL");
  return grid_mean_impl(h, ax, nx, ay, ny, az, nz, fid, nullptr, dmean);
}

int gpc_predict_cov(gpc_handle h, const double* Xs4, long M, double* mean, double* cov, const double* extra_diag,
                    unsigned flags) {
  int rc = require_factor(h);
  if (rc) return rc;
  if (M < 1 || !Xs4 || !cov) return fail(h, GPC_ERR_SHAPE, "bad test set");
  if (M > 32768) return fail(h, GPC_ERR_SHAPE, "full covariance limited to M <= 32768");
  if ((rc = check_fidelity_rows(h, Xs4, M, "test rows"))) return rc;
  CK(cudaSetDevice(h->device));
  const long m_pad = round_up(M, 128);
  if ((rc = ensure_pred_ws(h, m_pad, true, false))) return rc;
  CK(h->Xs4.ensure((size_t)m_pad * 32));
  CK(h->cov.ensure((size_t)M * M * 8));
  CK(h->mean.ensure((size_t)m_pad * 8));
  CK(cudaMemcpyAsync(h->Xs4.p, Xs4, (size_t)M * 32, cudaMemcpyHostToDevice, h->stream));
  if (extra_diag) {
    CK(h->ediag.ensure((size_t)M * 8));
    CK(cudaMemcpyAsync(h->ediag.p, extra_diag, (size_t)M * 8, cudaMemcpyHostToDevice, h->stream));
  }
  if ((rc = launch_kstar<false, true>(h, h->Xs4.d(), M, m_pad))) return rc;
  if ((rc = launch_vt<true, false>(h, m_pad))) return rc;
  const int mt = (int)(m_pad / 128);
  k_cov<<<dim3(mt, mt), gpcg::NTHREADS, gpcg::SMEM_BYTES, h->stream>>>(
      h->hyp, h->Vt.d(), h->n_pad, h->Xs4.d(), M, extra_diag ? h->ediag.d() : nullptr, h->cov.d(), M, 0, flags);
  CKL();
  if (mean) {
    const int nchunks = (int)((h->n_pad + KS_COLS - 1) / KS_COLS);
    k_finalize_pred<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(
        h->hyp, h->Xs4.d(), M, m_pad, h->meanpart.d(), nchunks, h->sumsq.d(), h->nb, nullptr, nullptr, 0, h->mean.d(),
        nullptr, flags);
    CKL();
    CK(cudaMemcpyAsync(mean, h->mean.p, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaMemcpyAsync(cov, h->cov.p, (size_t)M * M * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return GPC_OK;
}

int gpc_mean_grad(gpc_handle h, const double* Xs4, long M, double* mean, double* grads) {
  int rc = require_factor(h);
  if (rc) return rc;
  if (M < 1 || !Xs4 || !grads) return fail(h, GPC_ERR_SHAPE, "bad test set");
  if (h->F != 1 || h->hyp.base != 0)
    return fail(h, GPC_ERR_ARG, "mean gradients are defined for the single-fidelity squared-exponential kernel");
  CK(cudaSetDevice(h->device));
  const long mc = h->m_chunk;
  CK(h->Xs4.ensure((size_t)mc * 32));
  CK(h->mean.ensure((size_t)mc * 8));
  CK(h->grads.ensure((size_t)mc * 24));
  const int nchunks = (int)((h->n_pad + KS_COLS - 1) / KS_COLS);
  for (long o = 0; o < M; o += mc) {
    const long m = (M - o) < mc ? (M - o) : mc;
    const long m_pad = round_up(m, 128);
    if ((rc = ensure_pred_ws(h, m_pad, false, true))) return rc;
    CK(cudaMemcpyAsync(h->Xs4.p, Xs4 + o * 4, (size_t)m * 32, cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_kstar<true, false>(h, h->Xs4.d(), m, m_pad))) return rc;
    k_finalize_pred<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(
        h->hyp, h->Xs4.d(), m, m_pad, h->meanpart.d(), nchunks, h->sumsq.d(), h->nb, nullptr, nullptr, 0, h->mean.d(),
        nullptr, 0u);
    CKL();
    k_finalize_grad<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(h->gradpart.d(), nchunks, m, m_pad,
                                                                         h->grads.d());
    CKL();
    if (mean) CK(cudaMemcpyAsync(mean + o, h->mean.p, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(grads + o * 3, h->grads.p, (size_t)m * 24, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return GPC_OK;
}

int gpc_spd_stats(gpc_handle h, const double* cov, long M, const double* e, double* quad, double* fro_inv,
                  double* logdet) {
  if (!h || !cov) return GPC_ERR_ARG;
  if (M < 1 || M > 32768) return fail(h, GPC_ERR_SHAPE, "gpc_spd_stats: 1 <= M <= 32768");
  CK(cudaSetDevice(h->device));
  const long mp = round_up(M, 128);
  const int mt = (int)(mp / 128);
  CK(h->gS.ensure((size_t)mp * mp * 8));
  CK(h->gSinv.ensure((size_t)mp * mp * 8));
  CK(h->gT.ensure((size_t)mp * mp * 8));
  CK(h->cand_aux.ensure(64));
  CK(h->scal.ensure(64));
  CK(h->gram.ensure((size_t)(mt * mt + 2 * mp) * 8));
  cudaStream_t s = h->stream;
  CK(cudaMemcpy2DAsync(h->gS.p, (size_t)mp * 8, cov, (size_t)M * 8, (size_t)M * 8, (size_t)M, cudaMemcpyHostToDevice, s));
  k_pad_identity<<<dim3((unsigned)((mp + 255) / 256), (unsigned)mp), 256, 0, s>>>(h->gS.d(), M, mp);
  CKL();
  int* dstat = reinterpret_cast<int*>(static_cast<char*>(h->cand_aux.p) + 32);
  int rc = factor_matrix(h, h->gS.d(), h->gSinv.d(), h->gT.d(), mp, dstat);
  if (rc) return rc;
  k_logdet_fit<<<1, 1024, 0, s>>>(h->gS.d(), mp, M, nullptr, nullptr, h->scal.d());
  CKL();
  double* part = h->gram.d();
  double* ev = part + (size_t)mt * mt;
  double* zv = ev + mp;
  k_aat_fro<<<dim3(mt, mt), gpcg::NTHREADS, gpcg::SMEM_BYTES, s>>>(h->gSinv.d(), mp, mt, part);
  CKL();
  if (e) {
    CK(cudaMemsetAsync(ev, 0, (size_t)mp * 8, s));
    CK(cudaMemcpyAsync(ev, e, (size_t)M * 8, cudaMemcpyHostToDevice, s));
    k_trmv_n<<<(unsigned)((mp + 7) / 8), 256, 0, s>>>(h->gSinv.d(), mp, mp, ev, nullptr, 0.0, 1.0, zv);
    CKL();
    k_sumsq_vec<<<1, 1024, 0, s>>>(zv, M, h->scal.d() + 2);
    CKL();
  }
  int st = 0;
  double sc[3] = {0, 0, 0};
  std::vector<double> hp((size_t)mt * mt);
  CK(cudaMemcpyAsync(&st, dstat, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(sc, h->scal.p, 24, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hp.data(), part, (size_t)mt * mt * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (st != 0 || !std::isfinite(sc[0])) return fail(h, GPC_ERR_NOT_PD, "gpc_spd_stats: matrix not positive definite");
  double f2 = 0.0;
  for (int a = 0; a < mt; ++a)
    for (int b = 0; b <= a; ++b) f2 += (a == b ? 1.0 : 2.0) * hp[(size_t)a * mt + b];
  f2 -= (double)(mp - M);  // the identity padding block
  if (logdet) *logdet = sc[0];
  if (quad) *quad = e ? sc[2] : 0.0;
  if (fro_inv) *fro_inv = std::sqrt(f2 > 0.0 ? f2 : 0.0);
  return GPC_OK;
}

int gpc_traj_points(gpc_handle h, long C, const long* edge_off, const double* edge_xy, const long* prim_off,
                    const double* prims, double variance_rate, double meas_rate, int dense, int with_var, double t_off,
                    const double* fid_levels, int max_pts, double* pts, double* fid, long* counts) {
  if (!h || !edge_off || !edge_xy || !prim_off || !prims || !pts || !counts) return GPC_ERR_ARG;
  if (C < 0 || max_pts < 1) return fail(h, GPC_ERR_SHAPE, "bad path set");
  if (!(meas_rate > 0.0)) return fail(h, GPC_ERR_ARG, "meas_rate must be positive");
  if (C == 0) return GPC_OK;
  CK(cudaSetDevice(h->device));
  const long E = edge_off[C], P = prim_off[E];
  for (long c = 0; c < C; ++c)
    if (edge_off[c + 1] - edge_off[c] > GPC_TRAJ_MAXEDGE) return fail(h, GPC_ERR_SHAPE, "a path has more than 32 edges");
  DevBuf d_eo, d_xy, d_po, d_pr, d_wp, d_pts, d_fid, d_cnt;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](DevBuf& b, size_t n) { if (e == cudaSuccess) e = b.ensure(n ? n : 8); };
  alloc(d_eo, (size_t)(C + 1) * 8); alloc(d_xy, (size_t)E * 32); alloc(d_po, (size_t)(E + 1) * 8);
  alloc(d_pr, (size_t)P * 32); alloc(d_wp, (size_t)(P + E) * 32); alloc(d_pts, (size_t)C * max_pts * 40);
  alloc(d_fid, (size_t)C * max_pts * 8); alloc(d_cnt, (size_t)C * 8);
  int rc = GPC_OK;
  if (e != cudaSuccess) {
    rc = fail(h, GPC_ERR_CUDA, cudaGetErrorString(e));
  } else {
    cudaStream_t s = h->stream;
    cudaMemcpyAsync(d_eo.p, edge_off, (size_t)(C + 1) * 8, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_xy.p, edge_xy, (size_t)E * 32, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_po.p, prim_off, (size_t)(E + 1) * 8, cudaMemcpyHostToDevice, s);
    if (P) cudaMemcpyAsync(d_pr.p, prims, (size_t)P * 32, cudaMemcpyHostToDevice, s);
    cudaMemsetAsync(d_pts.p, 0, (size_t)C * max_pts * 40, s);
    cudaMemsetAsync(d_fid.p, 0, (size_t)C * max_pts * 8, s);
    k_traj_points<<<(unsigned)C, 128, 0, s>>>(static_cast<const long*>(d_eo.p), d_xy.d(), static_cast<const long*>(d_po.p),
                                             d_pr.d(), d_wp.d(), variance_rate, meas_rate, dense, with_var, t_off,
                                             fid_levels ? fid_levels[0] : 0.0, fid_levels ? fid_levels[1] : 0.0,
                                             fid_levels ? 1 : 0, max_pts, d_pts.d(), fid ? d_fid.d() : nullptr,
                                             static_cast<long*>(d_cnt.p));
    ++h->launches;
    cudaMemcpyAsync(pts, d_pts.p, (size_t)C * max_pts * 40, cudaMemcpyDeviceToHost, s);
    if (fid) cudaMemcpyAsync(fid, d_fid.p, (size_t)C * max_pts * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(counts, d_cnt.p, (size_t)C * 8, cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(h, GPC_ERR_CUDA, cudaGetErrorString(e));
  }
  DevBuf* all[] = {&d_eo, &d_xy, &d_po, &d_pr, &d_wp, &d_pts, &d_fid, &d_cnt};
  for (DevBuf* b : all) b->release();
  if (rc == GPC_OK)
    for (long c = 0; c < C; ++c)
      if (counts[c] > max_pts) return fail(h, GPC_ERR_SHAPE, "a path produced more points than max_pts (counts[] holds the sizes)");
  return rc;
}

void* gpc_stream(gpc_handle h) { return h ? (void*)h->stream : nullptr; }
long gpc_launch_count(gpc_handle h) { return h ? h->launches : 0; }

int gpc_set_mode(gpc_handle h, int mode) {
  if (!h) return GPC_ERR_ARG;
  if (mode != GPC_MODE_FP64 && mode != GPC_MODE_INT8 && mode != GPC_MODE_INT8_F32 && mode != GPC_MODE_INT8_L5)
    return fail(h, GPC_ERR_ARG, "unknown mode");
  h->mode = mode;
  return GPC_OK;
}

int gpc_get_mode(gpc_handle h) { return h ? h->mode : -1; }

int gpc_set_chunk(gpc_handle h, long m_chunk) {
  if (!h) return GPC_ERR_ARG;
  if (m_chunk < 128 || m_chunk % 128) return fail(h, GPC_ERR_SHAPE, "chunk must be a positive multiple of 128");
  h->m_chunk = m_chunk;
  return GPC_OK;
}

int gpc_enable_hot_timing(gpc_handle h, int on) {
  if (!h) return GPC_ERR_ARG;
  h->hot_timing = on != 0;
  return GPC_OK;
}

int gpc_last_call_device_ms(gpc_handle h, double* ms) {
  if (!h || !ms) return GPC_ERR_ARG;
  *ms = h->last_call_ms;
  return GPC_OK;
}

int gpc_hot_kernel_time(gpc_handle h, double* ms_total, long* launches, double* flops, int reset) {
  if (!h) return GPC_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    h->hot_ms += ms;
    ++h->hot_launches;
  }
  h->ev_used = 0;
  if (ms_total) *ms_total = h->hot_ms;
  if (launches) *launches = h->hot_launches;
  if (flops) *flops = h->hot_flops;
  if (reset) { h->hot_ms = 0.0; h->hot_launches = 0; h->hot_flops = 0.0; }
  return GPC_OK;
}

}  // extern "C"

#include "gpc_ig_api.inc"
