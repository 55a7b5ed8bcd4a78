/*
 * gpcore.h -- C ABI of libgpcore.so, the B200 (sm_100a) dense Gaussian-process inference core.
 *
 * The reference (colem404/Adaptive-Exploration-...-Multi-fidelity-Gaussian-Processes) is pure
 * Python and has no FFI layer; its boundary is the duck-typed model protocol
 * (fit / predict / predict_covariance / set_XY / set_data / CalcCost).  Each entry point below
 * names the reference call it replaces (paths relative to the reference checkout).  The Python
 * host shim (package `gpcore`) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns an int status (GPC_OK == 0); gpc_last_error() gives the message;
 *   - all matrices are C-contiguous float64; "X4" rows are (x, y, z, fidelity index) -- unused
 *     spatial columns must be 0, single-fidelity models ignore the 4th column;
 *   - pointers are HOST pointers unless the function name ends in _dev (then they are device
 *     pointers on the handle's device and the call is asynchronous on gpc_stream());
 *   - the caller owns every buffer it passes; the handle owns all device state;
 *   - a handle is not thread-safe: one handle per host thread (the reference only ever calls
 *     its GP from the planner thread, PhysicalExperimentCode/GraceExplorationExperiments_MFGP.py:767);
 *   - a non positive-definite covariance is a recoverable status (GPC_ERR_NOT_PD), mirroring the
 *     LinAlgError the reference catches at NIGP.py:156 and ...MFGP.py:392.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns GPC_ERR_CUDA.
 */
#ifndef GPCORE_H
#define GPCORE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpc_handle_s* gpc_handle;

enum gpc_status {
  GPC_OK = 0,
  GPC_ERR_NOT_PD = 1, /* Cholesky met a non-positive pivot */
  GPC_ERR_SHAPE = 2,  /* bad N / M / k / F / parameter-vector length */
  GPC_ERR_CUDA = 3,   /* CUDA runtime error (message in gpc_last_error) */
  GPC_ERR_STATE = 4,  /* call order: data/hypers/factor missing */
  GPC_ERR_ARG = 5
};

enum gpc_kind {
  GPC_SF_RBF = 0,       /* GPy GPRegression + RBF(ARD)       GPTrainers.py:80-84 */
  GPC_SF_MAT32 = 1,     /* GPy GPRegression + Matern32(ARD)  PhysicalExperimentCode/...SFGP.py:610 */
  GPC_MF_AR1_RBF = 2,   /* emukit LinearMultiFidelityKernel over RBF    GPTrainers.py:62-66 */
  GPC_MF_AR1_MAT32 = 3, /* ... over Matern32   PhysicalExperimentCode/...MFGP.py:656 */
  GPC_NIGP = 4          /* NIGP.py SE-ARD with heteroscedastic diagonal; sigma_f is the kernel variance */
};

/* gpc_predict flags */
#define GPC_INCLUDE_NOISE 1u /* add the likelihood variance (GPy predict default)                */
#define GPC_CLIP_DIAG 2u     /* clip the latent marginal variance at 1e-15 (GPy)                  */
#define GPC_CLIP_COV 4u      /* clip the whole covariance element-wise at 1e-10 (emukit wrapper)  */
#define GPC_NIGP_FLOOR 8u    /* NIGP.py:327-333: var = max(var + 1e-12, 1e-12); cov += 1e-12 I    */
#define GPC_MEAN_ONLY 16u    /* skip the variance (var may be NULL)                               */

/* gpc_set_mode: how the posterior variance contraction V = L^-1 K* runs
 *   GPC_MODE_FP64  FP64 DMMA (mma.sync.m8n8k4.f64), bounded by the 37 TFLOP/s FP64 pipe of sm_100a
 *   GPC_MODE_INT8  (default) Ozaki splitting on the tcgen05 INT8 tensor cores: both operands are cut
 *                  into 6 balanced base-256 digits, the 21 digit GEMMs are exact in int32 (TMEM), their
 *                  recombination carries 2^-48 of the operand scales -- results agree with the FP64 path
 *                  to ~1e-12 relative, far inside the 1e-9 parity tolerance (tests/test_gpu_int8.py).
 *                  Used for N <= 16384 (int32 accumulator range); larger problems run in FP64. */
#define GPC_MODE_FP64 0
#define GPC_MODE_INT8 1
/*   GPC_MODE_INT8_F32  the optional reduced-precision mode of the north star ("1e-4 for an optional FP32 mode"): the same
 *                  INT8 contraction with only the 4 most significant digits of each operand (digit pairs p + q <= 3:
 *                  10 digit GEMMs instead of 21, the low-order slices are neither copied nor multiplied).  Balanced
 *                  digits truncate to nearest, so the results carry ~2^-30 of the operand scales: variances to ~1e-7
 *                  relative -- three orders inside an FP32 evaluation of the same cancellation-prone difference --
 *                  at about half the contraction time.  The mean is unaffected (it is formed in FP64 either way). */
#define GPC_MODE_INT8_F32 2
/*   GPC_MODE_INT8_L5   five of the six digit levels (15 digit GEMMs, -29 % tensor work): posterior variances to ~5e-10
 *                  normwise at the BASELINE sizes -- inside the 1e-9 tolerance by a factor of two only, which is why it
 *                  is NOT the default -- and information gain outside 1e-9 (~3e-8).  An explicit speed / margin
 *                  trade for callers that only need the posterior. */
#define GPC_MODE_INT8_L5 3

/* gpc_ig_seq flags */
#define GPC_IG_FIRST_PREADDED 1u /* GraceRIGV3.py:454-455: point 0 is appended before it is predicted */

/* ---- life cycle ------------------------------------------------------------------------- */
/* F = number of fidelities (1 for SF / NIGP, <= 4).  device = CUDA ordinal. */
int gpc_create(int kind, int F, int device, gpc_handle* out);
int gpc_destroy(gpc_handle h);
const char* gpc_last_error(gpc_handle h); /* h may be NULL: last creation error */
int gpc_version(void);

/* ---- model state -------------------------------------------------------------------------- */
/* Flat hyper-parameter vector in the reference's own `param_array` order:
 *   SF   (n=5):  variance, lx, ly, lz, noise_var                    (...SFGP.py:620)
 *   MF   (n=4F + F-1 + (1|F)): per fidelity (variance, lx, ly, lz), rho_1..rho_{F-1},
 *         noise | noise_0..noise_{F-1}                               (...MFGP.py:670)
 *   NIGP (n=5):  lx, ly, lz, sigma_f (kernel variance!), sigma_y (std) (NIGP.py:137-141)
 * jitter is added to the training diagonal: 1e-8 for the GPy models (exact_gaussian_inference),
 * 1e-8 inside NIGP's NLML (NIGP.py:152-154), 0 in NIGP.predict (NIGP.py:287). */
int gpc_set_hypers(gpc_handle h, const double* flat, int n, double jitter);

/* Training set: X4 is N x 4, y is N, extra_noise_diag (N, may be NULL) is the per-point variance
 * added to the diagonal (NIGP's v_i, NIGP.py:147,286).  Replaces set_XY / set_data. */
int gpc_set_data(gpc_handle h, const double* X4, const double* y, const double* extra_noise_diag, long N);

/* Assemble K + noise, Cholesky-factor it, solve for alpha, build L^-1.  Outputs may be NULL.
 * nlml = 0.5 y'alpha + 0.5 logdet + 0.5 N log(2 pi)  (NIGP.py:159-161; GPy -log_marginal). */
int gpc_factor(gpc_handle h, double* nlml, double* logdet);

/* Analytic gradient of the NLML of the current factorisation with respect to the flat
 * hyper-parameter vector (same order and length n as gpc_set_hypers):
 *   d NLML / d theta = 1/2 sum_ij (Ky^-1 - alpha alpha^T)_ij dKy_ij / d theta.
 * Replaces the numerical differentiation of the reference's fits (NIGP.py:231-239 calls the
 * objective 2D+3 times per L-BFGS-B step; GPTrainers.py:68,84,94 optimise inside GPy).
 * diagW (N, may be NULL) receives diag(Ky^-1 - alpha alpha^T) in the caller's row order -- what
 * NIGP needs for the sigma_x gradient (d v_i / d sigma_x_d = 2 sigma_x_d grad_id^2). */
int gpc_nlml_grad(gpc_handle h, double* grad, int n, double* diagW);

/* Copies of the factor state for tests and for the multi-GPU broadcast.  alpha is returned in the
 * caller's training-row order; L and L^-1 are those of the internal order, which for
 * multi-fidelity models is the training set stably sorted by fidelity index. */
int gpc_get_alpha(gpc_handle h, double* alpha /* N */);
int gpc_get_chol(gpc_handle h, double* L /* N x N, lower, row-major */);
int gpc_get_linv(gpc_handle h, double* Linv /* N x N, lower, row-major */);
long gpc_padded_n(gpc_handle h);
/* Device pointers of the replicated factor state (padded) so that rank 0 can
 * NCCL-broadcast it; after receiving, the other ranks call gpc_adopt_factor(). */
int gpc_factor_state_dev(gpc_handle h, double** L, double** Linv, double** alpha, long* n_pad);
int gpc_adopt_factor(gpc_handle h, double logdet);

/* ---- kernel matrix (NIGP.py:11-20 SE_ARD_kernel; gpy_model.kern.K, GraceRIGV3.py:515) ------ */
/* K is na x nb; Xb4 == NULL means K(Xa, Xa).  No noise is added. */
int gpc_kernel_matrix(gpc_handle h, const double* Xa4, long na, const double* Xb4, long nb, double* K);

/* ---- posterior (NIGP.py:269-333; GPy predict; emukit wrapper predict) ----------------------- */
int gpc_predict(gpc_handle h, const double* Xs4, long M, double* mean, double* var, unsigned flags);
int gpc_predict_dev(gpc_handle h, const double* dXs4, long M, double* dmean, double* dvar, unsigned flags);
/* NIGP.predict(Xs, Xs_input_noise=sx) (NIGP.py:304-324): var += sum_d (d mean / d x_d)^2 sx_d^2,
 * fused on the device; sx is (1 x 3) when sx_rows == 1, else (M x 3) input-noise std devs. */
int gpc_predict_noisy(gpc_handle h, const double* Xs4, long M, const double* sx, long sx_rows,
                      double* mean, double* var, unsigned flags);
/* Posterior MEAN on a tensor grid (squared-exponential kernels only): grid point (i, j, k) = (ax[i], ay[j], az[k])
 * at fidelity index fid, mean[(i * ny + j) * nz + k] -- np.meshgrid(ax, ay, az, indexing="ij") raveled in C order.
 * Every test set of the reference is such a grid (exploreSimSettings.py:116-119, the planner's fieldGrid); the
 * cross-covariance separates per axis, so the mean is (nx ny) x nz x N GEMMs on the FP64 tensor cores (2 M N flop)
 * instead of M N kernel evaluations.  Same values as gpc_predict's mean to O(ulp) per product of three exps. */
int gpc_predict_grid_mean(gpc_handle h, const double* ax, long nx, const double* ay, long ny, const double* az,
                          long nz, double fid, double* mean);
/* The same with a DEVICE output pointer (axes stay host arrays: they are tiny); asynchronous on gpc_stream(h). */
int gpc_predict_grid_mean_dev(gpc_handle h, const double* ax, long nx, const double* ay, long ny, const double* az,
                              long nz, double fid, double* dmean);
/* Full M x M posterior covariance (GPy predict(full_cov=1), emukit predict_covariance,
 * NIGP.predict(return_cov=1)); mean may be NULL.  extra_diag (M, may be NULL) is added to the
 * diagonal (NIGP's test-input noise term, NIGP.py:321-324). */
int gpc_predict_cov(gpc_handle h, const double* Xs4, long M, double* mean, double* cov,
                    const double* extra_diag, unsigned flags);
/* Posterior mean and its input gradients, NIGP.py:55-64 / :307-311:
 * grads[n][d] = sum_j alpha_j k(x_n, x_j) (-(x_nd - x_jd) / l_d^2);  grads is M x 3. */
int gpc_mean_grad(gpc_handle h, const double* Xs4, long M, double* mean, double* grads);

/* ---- information gain (GraceRIGV3.py:443-562, PhysicalExperimentCode/GraceRIGV3.py:571-678) - */
/* Candidates are ragged: candidate c owns rows offsets[c] .. offsets[c+1]-1 of Xc4 (<= 64 rows;
 * an empty candidate scores 0).
 * Sequential IG (calcPathInfoSF2/SF3, calculatePathInfoEmu):
 *   I_c = sum_i log(1 + s_i / sig_n), s_i = noise-inclusive predictive variance of point i
 *   (queried at fidelity pred_fid when pred_fid >= 0, else at its own fidelity) given the data and
 *   the candidate's earlier points appended with zero targets.
 * row_mask (one byte per row of Xc4, may be NULL = every row conditioned, none pre-appended):
 *   bit0 = later rows condition on this row; bit1 = the row is appended before it is predicted
 *   itself (the windowed variants calcPathInfoSF/SF4, calculatePathInfoEmu once > 100 points).
 * best (may be NULL) receives argmax_c I_c (first index on ties, -1 if C == 0). */
int gpc_ig_seq(gpc_handle h, const double* Xc4, const long* offsets, long C, double sig_n,
               int pred_fid, unsigned flags, const unsigned char* row_mask, double* I_out, long* best);
/* Log-det IG on a fixed grid (calcPathInfoSFBatch / calculatePathInfoEmuBatch):
 *   I_c = 0.5 (logdet S_prior(grid) - logdet S_post(grid | data u X_c)),
 * S = noise-inclusive predictive covariance.  The raw value is returned (the reference's
 * max(.,0) / det==0 guards live in the Python shim).  logdet_prior (may be NULL) receives
 * logdet S_prior.  G <= 4096. */
int gpc_ig_logdet(gpc_handle h, const double* grid4, long G, const double* Xc4, const long* offsets,
                  long C, double* I_out, double* logdet_prior, long* best);
/* The same with flags.  GPC_CLIP_COV reproduces emukit's GPyMultiOutputWrapper.predict_covariance, which
 * returns np.clip(cov, 1e-10, inf): calculatePathInfoEmuBatch (PhysicalExperimentCode/GraceRIGV3.py:599-618)
 * takes BOTH determinants of element-wise clipped G x G matrices (every negative posterior covariance
 * between two grid points becomes 1e-10), so each candidate costs one G x G factorisation instead of
 * a k x k determinant-lemma update; logdet_prior is then the log-det of the clipped prior.  NaN for a
 * candidate whose clipped matrix is not positive definite.  Needs k_max * round_up(G, 32) <= ~22000. */
int gpc_ig_logdet_ex(gpc_handle h, const double* grid4, long G, const double* Xc4, const long* offsets,
                     long C, unsigned flags, double* I_out, double* logdet_prior, long* best);
/* "Self-grid" log-det IG (calculatePathInfoEmu2, GraceRIGV3.py:505-523): the grid is the candidate
 * itself queried at pred_fid (>= 0; < 0 = each point's own fidelity),
 *   I_c = 0.5 (logdet K(Xp) - logdet S_post(Xp | data u X_c)),
 * K(Xp) = prior kernel matrix (no noise), S_post = noise-inclusive predictive covariance, clipped
 * element-wise at 1e-10 when flags has GPC_CLIP_COV (emukit predict_covariance).  NaN when a
 * matrix is numerically not positive definite (the reference takes log(det) of an LU there). */
int gpc_ig_selfgrid(gpc_handle h, const double* Xc4, const long* offsets, long C, int pred_fid,
                    unsigned flags, double* I_out, long* best);
/* The IG calls time their dominant kernel (the same L^-1 K* contraction) through the
 * gpc_hot_kernel_time hooks below. */

/* ---- candidate generation (GraceRIGV3.py:235-294 evaluateTraj, :373-427 edgePointsToTrajPoints /
 *      pathToTrajPoints, :508-512 fidelity labels) -------------------------------------------------- */
/* Path c owns edges edge_off[c] .. edge_off[c+1]-1 (<= 32); edge e has start/end node positions
 * edge_xy[e] = (x0, y0, x1, y1) and the motion primitives prim_off[e] .. prim_off[e+1]-1, each
 * (type, a, b, c): 0 spiral (dz, _, speed), 1 glide (pitch, dz, speed), 2 swim (dist, speed), 3 flat dive
 * (dz, speed).  dense != 0 resamples every edge at meas_rate (np.arange / np.interp), otherwise the
 * way-points themselves are returned; rows equal after rounding to 4 decimals are dropped keeping the
 * first.  pts is C x max_pts x 5 (x, y, z, t, var); fid (C x max_pts, may be NULL) is the fidelity
 * index from fid_levels (2 if var < fl[0], 1 if fl[0] < var < fl[1], else 0); counts[c] = points of path c.
 * GPC_ERR_SHAPE when a path yields more than max_pts points (counts[] then tells the sizes). */
int gpc_traj_points(gpc_handle h, long C, const long* edge_off, const double* edge_xy, const long* prim_off,
                    const double* prims, double variance_rate, double meas_rate, int dense, int with_var,
                    double t_off, const double* fid_levels, int max_pts, double* pts, double* fid, long* counts);

/* ---- evaluator (GPTrainers.py:121-137) -------------------------------------------------------- */
/* For a symmetric positive-definite M x M matrix cov (host, row-major) and a vector e (M, may be
 * NULL): quad = e^T inv(cov) e, fro_inv = ||inv(cov)||_F, logdet = log det cov -- through one
 * Cholesky + triangular inverse on the device instead of np.linalg.inv.  The covariance-weighted
 * MSE of the reference is quad / fro_inv / M.  Any handle kind works; its model state is untouched. */
int gpc_spd_stats(gpc_handle h, const double* cov, long M, const double* e, double* quad, double* fro_inv,
                  double* logdet);

/* ---- measurement hooks ---------------------------------------------------------------------- */
void* gpc_stream(gpc_handle h);                 /* cudaStream_t all kernels are launched on     */
long gpc_launch_count(gpc_handle h);            /* kernels launched by this handle so far        */
int gpc_set_chunk(gpc_handle h, long m_chunk);  /* test points per launch batch (default 65536)  */
int gpc_set_mode(gpc_handle h, int mode);       /* GPC_MODE_FP64 | GPC_MODE_INT8                   */
int gpc_get_mode(gpc_handle h);
/* Device time of the dominant kernel (the L^-1 K* DMMA contraction) accumulated with CUDA events
 * on gpc_stream() since the last reset, the number of its launches and the floating-point
 * operations those launches executed (2 x multiply-adds over the padded triangular k-range). */
int gpc_hot_kernel_time(gpc_handle h, double* ms_total, long* launches, double* flops, int reset);
/* Device time (CUDA events on gpc_stream(), first to last kernel, the staged uploads of the candidate rows
 * included) of the most recent information-gain call on this handle: the rate with the inputs already in HBM. */
int gpc_last_call_device_ms(gpc_handle h, double* ms);
int gpc_enable_hot_timing(gpc_handle h, int on);

#ifdef __cplusplus
}
#endif
#endif /* GPCORE_H */
