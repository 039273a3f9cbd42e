// jb_dist_*: the multi-GPU forms of the hot path behind the C ABI — ONE host process drives
// every GPU of the box, so a `foreign import ccall` caller (/root/reference/Numeric/Jalla/Foreign/BLAS.chs:155,
// Foreign/LAPACKE.chs:691-694,703-706,733-736,741-744) reaches them like any other symbol.  Per device: its own
// streams and workspace pool; all devices peer-mapped, and every inter-GPU transfer is a copy-engine peer copy over
// NVLink ordered by events (cudaMemcpy2DAsync on side streams) — no collective kernel takes SMs from the persistent GEMM.
//
//   jb_dist_dgemm          2-D blocks on a Pr x Pc grid (SUMMA): device (i,j) receives A[i, K-block j] and B[K-block i, j]
//                          from the host once, pulls the other K panels from its row / column peers, multiplies panel by
//                          panel starting with what it already holds, and returns its C block.
//   jb_dist_dgetrf/getrs   right-looking LU with partial pivoting on a block-cyclic layout whose process grid is 1 x P
//                          (column blocks of width nb dealt round-robin): a panel is factorised where it lives with the
//                          full column height local — the pivot search needs no exchange — then it and its pivots are
//                          broadcast over NVLink; every device interchanges, solves and updates its own columns; the next
//                          panel is factorised one step ahead (look-ahead 1) on a second stream of its owner.
//                          On NVSwitch every peer is one hop, so the 1 x P grid trades nothing for the cheap pivoting.
//   jb_dist_dpotrf/potrs   the same layout for Cholesky (panel = diagonal block + the solve of the rows below it).
//   jb_dist_dgesv_batched  batch slices per device, no communication.
// Operands are host (pinned or pageable) or any UVA pointer; results return to them (LAPACK semantics).  The factors also
// stay resident in the handle, so a following ?getrs / ?potrs on the same matrix does not move them again.
#include "jb_common.cuh"
#include <vector>
#include <algorithm>
#include <chrono>

namespace jb {

struct DevGuard {
  int prev = 0;
  explicit DevGuard(int d) { cudaGetDevice(&prev); cudaSetDevice(d); }
  ~DevGuard() { cudaSetDevice(prev); }
};

struct DevCtx {
  int dev = 0;
  cudaStream_t s = nullptr, s2 = nullptr, sc = nullptr;     // compute, look-ahead panel, copies
  double *L = nullptr;            // local columns of a resident factorisation (n x nloc, ld)
  double *pbuf[2] = {nullptr, nullptr};
  int *ipiv = nullptr;            // full pivot vector (every device applies every panel's interchanges)
  int *info = nullptr;
  size_t l_bytes = 0, p_bytes = 0;
};

struct Dist {
  int pr = 1, pc = 1, P = 1, nb = 512;
  std::vector<DevCtx> d;
  // resident factorisation
  int kind = 0;                   // 0 none, 1 LU, 2 Cholesky (lower storage)
  int n = 0, ld = 0;
  const void *host_a = nullptr;
  double ms[3] = {0, 0, 0};       // last factorisation: host -> devices, factorisation, devices -> host (jb_dist_last_ms)
};

#define DCHK(call)                                                                                          \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess) {                                                                               \
      set_error(JB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);  \
      cudaGetLastError();                                                                                   \
      return JB_ERR_CUDA;                                                                                   \
    }                                                                                                       \
  } while (0)

static int sync_all(Dist *h) {
  for (auto &c : h->d) {
    DevGuard g(c.dev);
    DCHK(cudaStreamSynchronize(c.s)); DCHK(cudaStreamSynchronize(c.s2)); DCHK(cudaStreamSynchronize(c.sc));
  }
  return 0;
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int nloc_blocks(int nblk, int P, int d) { return nblk > d ? (nblk - d + P - 1) / P : 0; }

static void free_resident(Dist *h) {
  for (auto &c : h->d) {
    DevGuard g(c.dev);
    cudaStreamSynchronize(c.s); cudaStreamSynchronize(c.s2); cudaStreamSynchronize(c.sc);
    if (c.L) cudaFree(c.L);
    for (auto &p : c.pbuf) { if (p) cudaFree(p); p = nullptr; }
    if (c.ipiv) cudaFree(c.ipiv);
    c.L = nullptr; c.ipiv = nullptr; c.l_bytes = c.p_bytes = 0;
  }
  h->kind = 0; h->host_a = nullptr; h->n = 0;
}

// (re)allocates the block-cyclic storage for an n x n matrix: local columns, two panel buffers, pivots
static int alloc_resident(Dist *h, int n) {
  if (h->n == n && !h->d.empty() && h->d[0].L) { h->kind = 0; h->host_a = nullptr; return 0; }     // same shape as last time: keep the buffers
  free_resident(h);
  const int nb = h->nb, nblk = (n + nb - 1) / nb;
  h->n = n; h->ld = n + (n & 1);
  for (int q = 0; q < h->P; q++) {
    DevCtx &c = h->d[q];
    DevGuard g(c.dev);
    const size_t cols = (size_t)std::max(nloc_blocks(nblk, h->P, q), 1) * nb;
    c.l_bytes = sizeof(double) * (size_t)h->ld * cols;
    c.p_bytes = sizeof(double) * (size_t)h->ld * nb;
    DCHK(cudaMalloc(&c.L, c.l_bytes));
    for (auto &p : c.pbuf) DCHK(cudaMalloc(&p, c.p_bytes));
    DCHK(cudaMalloc(&c.ipiv, sizeof(int) * (size_t)std::max(n, 1)));
  }
  return 0;
}

__global__ void add_offset_kernel(int *p, int n, int off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += off;
}
__global__ void first_info_kernel(const int *panel_info, int off, int *info) {     // keeps the first non-zero info (global index)
  if (*panel_info != 0 && *info == 0) *info = *panel_info + off;
}

// host (lda) <-> block-cyclic local columns; `upper_t`: the host holds the UPPER triangle and the devices its transpose
static int scatter_cols(Dist *h, const double *A, int lda, int n) {
  const int nb = h->nb, nblk = (n + nb - 1) / nb;
  for (int J = 0; J < nblk; J++) {
    DevCtx &c = h->d[J % h->P];
    DevGuard g(c.dev);
    const int jb = std::min(nb, n - J * nb);
    DCHK(cudaMemcpy2DAsync(c.L + (size_t)(J / h->P) * nb * h->ld, sizeof(double) * h->ld, A + (size_t)J * nb * lda, sizeof(double) * lda,
                           sizeof(double) * n, jb, cudaMemcpyDefault, c.s));
  }
  return 0;
}
static int gather_cols(Dist *h, double *A, int lda, int n, int) {
  const int nb = h->nb, nblk = (n + nb - 1) / nb;
  for (int J = 0; J < nblk; J++) {
    DevCtx &c = h->d[J % h->P];
    DevGuard g(c.dev);
    const int jb = std::min(nb, n - J * nb);
    DCHK(cudaMemcpy2DAsync(A + (size_t)J * nb * lda, sizeof(double) * lda, c.L + (size_t)(J / h->P) * nb * h->ld, sizeof(double) * h->ld,
                           sizeof(double) * n, jb, cudaMemcpyDefault, c.s));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ LU
static int dist_getrf(Dist *h, int n, double *A, int lda, int *ipiv, int *info_out) {
  if (int rc = alloc_resident(h, n)) return rc;
  const int nb = h->nb, P = h->P, ld = h->ld, nblk = (n + nb - 1) / nb;
  double t0 = now_ms();
  if (int rc = scatter_cols(h, A, lda, n)) return rc;
  if (int rc = sync_all(h)) return rc;
  h->ms[0] = now_ms() - t0; t0 = now_ms();
  std::vector<cudaEvent_t> ev_panel(nblk), ev_first(nblk);
  std::vector<std::vector<cudaEvent_t>> ev_recv(P, std::vector<cudaEvent_t>(nblk)), ev_upd(P, std::vector<cudaEvent_t>(nblk));
  int *d_panel_info = nullptr;
  for (int J = 0; J < nblk; J++) {
    { DevGuard g(h->d[J % P].dev); DCHK(cudaEventCreateWithFlags(&ev_panel[J], cudaEventDisableTiming)); }
    { DevGuard g(h->d[(J + 1) % P].dev); DCHK(cudaEventCreateWithFlags(&ev_first[J], cudaEventDisableTiming)); }   // recorded by the NEXT panel's owner
  }
  for (int q = 0; q < P; q++) {
    DevGuard g(h->d[q].dev);
    for (int J = 0; J < nblk; J++) {
      DCHK(cudaEventCreateWithFlags(&ev_recv[q][J], cudaEventDisableTiming));
      DCHK(cudaEventCreateWithFlags(&ev_upd[q][J], cudaEventDisableTiming));
    }
    DCHK(cudaMalloc(&h->d[q].info, sizeof(int) * 2));
    DCHK(cudaMemsetAsync(h->d[q].info, 0, sizeof(int) * 2, h->d[q].s));
  }
  (void)d_panel_info;
  int rc = 0;
  // panel J on its owner (stream `st`): factorise rows J0.. of its nb columns, make the pivots global
  auto factor_panel = [&](int J, cudaStream_t st) -> int {
    DevCtx &c = h->d[J % P];
    DevGuard g(c.dev);
    const int J0 = J * nb, jb = std::min(nb, n - J0);
    double *pan = c.L + (size_t)(J / P) * nb * ld + J0;
    DCHK(cudaMemsetAsync(c.info + 1, 0, sizeof(int), st));
    if (int r = getrf_device<double>(n - J0, jb, pan, ld, c.ipiv + J0, c.info + 1, st)) return r;
    add_offset_kernel<<<(jb + 255) / 256, 256, 0, st>>>(c.ipiv + J0, jb, J0);
    first_info_kernel<<<1, 1, 0, st>>>(c.info + 1, J0, c.info);
    count_launch(2);
    DCHK(cudaEventRecord(ev_panel[J], st));
    return 0;
  };
  // interchanges + block-row solve + rank-jb update of local columns [c0, c1) of device q with panel J
  auto update_cols = [&](int q, int J, const double *pan, int ldp, int c0, int c1, cudaStream_t st) -> int {
    if (c1 <= c0) return 0;
    DevCtx &c = h->d[q];
    const int J0 = J * nb, jb = std::min(nb, n - J0);
    double *cols = c.L + (size_t)c0 * ld;
    if (int r = laswp_device<double>(cols, ld, c1 - c0, c.ipiv, J0, J0 + jb, st)) return r;
    if (int r = trsm_device<double>(141, 122, 111, 132, jb, c1 - c0, 1.0, pan + J0, ldp, cols + J0, ld, st)) return r;
    const int mr = n - J0 - jb;
    if (mr > 0)
      if (int r = gemm_device<double>(111, 111, mr, c1 - c0, jb, -1.0, pan + J0 + jb, ldp, cols + J0, ld, 1.0, cols + J0 + jb, ld, st, 0)) return r;
    return 0;
  };
  rc = factor_panel(0, h->d[0].s);
  for (int J = 0; J < nblk && !rc; J++) {
    const int o = J % P, J0 = J * nb, jb = std::min(nb, n - J0);
    // broadcast panel J (rows J0.., with its pivots) to every other device
    for (int q = 0; q < P && !rc; q++) {
      if (q == o) continue;
      DevCtx &c = h->d[q];
      DevGuard g(c.dev);
      DCHK(cudaStreamWaitEvent(c.sc, ev_panel[J], 0));
      if (J >= 2) DCHK(cudaStreamWaitEvent(c.sc, ev_upd[q][J - 2], 0));            // the buffer's previous panel has been used
      const double *src = h->d[o].L + (size_t)(J / P) * nb * ld + J0;
      DCHK(cudaMemcpy2DAsync(c.pbuf[J & 1] + J0, sizeof(double) * ld, src, sizeof(double) * ld, sizeof(double) * (n - J0), jb, cudaMemcpyDefault, c.sc));
      DCHK(cudaMemcpyAsync(c.ipiv + J0, h->d[o].ipiv + J0, sizeof(int) * jb, cudaMemcpyDefault, c.sc));
      DCHK(cudaEventRecord(ev_recv[q][J], c.sc));
    }
    // every device updates its columns; the owner of panel J+1 does that block first and factorises it one step ahead
    for (int q = 0; q < P && !rc; q++) {
      DevCtx &c = h->d[q];
      DevGuard g(c.dev);
      const int nl = nloc_blocks(nblk, P, q);
      const int ncols_loc = nl == 0 ? 0 : (nl - 1) * nb + std::min(nb, n - ((nl - 1) * P + q) * nb);
      const int b0 = (J >= q) ? (J - q) / P + 1 : 0;                  // first local block with global index > J
      const int left_cols = (J >= q) ? ((J - q) / P) * nb + ((J % P == q) ? 0 : nb) : 0;   // local columns left of (and excluding) panel J
      const double *pan; int ldp = ld;
      if (q == o) { pan = c.L + (size_t)(J / P) * nb * ld; DCHK(cudaStreamWaitEvent(c.s, ev_panel[J], 0)); }
      else { pan = c.pbuf[J & 1]; DCHK(cudaStreamWaitEvent(c.s, ev_recv[q][J], 0)); }
      // columns already factorised take the interchanges only (LAPACK swaps whole rows)
      if (left_cols > 0) rc = laswp_device<double>(c.L, ld, left_cols, c.ipiv, J0, J0 + jb, c.s);
      int c0 = b0 * nb;
      if (!rc && J + 1 < nblk && (J + 1) % P == q) {
        const int jn = std::min(nb, n - (J + 1) * nb);
        rc = update_cols(q, J, pan, ldp, c0, c0 + jn, c.s);
        if (!rc) {
          DCHK(cudaEventRecord(ev_first[J], c.s));
          DCHK(cudaStreamWaitEvent(c.s2, ev_first[J], 0));
          rc = factor_panel(J + 1, c.s2);
        }
        c0 += jn;
        tls().gemm_cta_cap = std::max(num_sms() - 40, 8);      // leave SMs to the panel running one step ahead
      }
      if (!rc) rc = update_cols(q, J, pan, ldp, c0, ncols_loc, c.s);
      tls().gemm_cta_cap = 0;
      if (!rc) DCHK(cudaEventRecord(ev_upd[q][J], c.s));
    }
  }
  if (!rc) rc = sync_all(h);
  h->ms[1] = now_ms() - t0; t0 = now_ms();
  int info = 0;
  if (!rc) {
    // the first zero pivot over all panels: each owner kept the first of ITS panels
    for (int q = 0; q < P; q++) {
      int hi[2] = {0, 0};
      DevGuard g(h->d[q].dev);
      DCHK(cudaMemcpy(hi, h->d[q].info, sizeof(hi), cudaMemcpyDeviceToHost));
      if (hi[0] != 0 && (info == 0 || hi[0] < info)) info = hi[0];
    }
    rc = gather_cols(h, A, lda, n, 0);
    if (!rc) { DevGuard g(h->d[(nblk - 1) % P].dev); DCHK(cudaMemcpyAsync(ipiv, h->d[(nblk - 1) % P].ipiv, sizeof(int) * n, cudaMemcpyDefault, h->d[(nblk - 1) % P].s)); }
    // the last panel's owner holds every pivot only if it received all broadcasts: it did (every device does), except its own
    // earlier panels' pivots, which it produced itself
    if (!rc) rc = sync_all(h);
  }
  for (int J = 0; J < nblk; J++) { cudaEventDestroy(ev_panel[J]); cudaEventDestroy(ev_first[J]); }
  for (int q = 0; q < P; q++) {
    for (int J = 0; J < nblk; J++) { cudaEventDestroy(ev_recv[q][J]); cudaEventDestroy(ev_upd[q][J]); }
    DevGuard g(h->d[q].dev);
    cudaFree(h->d[q].info); h->d[q].info = nullptr;
  }
  h->ms[2] = now_ms() - t0;
  if (rc) { free_resident(h); h->n = 0; return rc; }
  h->kind = 1; h->host_a = A;
  *info_out = info;
  return 0;
}

// Chained substitution on the resident factors: the right-hand sides travel from owner to owner (nrhs is small in
// Jalla's calls, so this phase is latency / bandwidth bound, not compute bound).
//   mode 0: P^T, L (unit) forward, U backward  (getrs 'N');  mode 2: L forward, L^T backward  (potrs, lower storage)
static int dist_solve(Dist *h, int mode, int n, int nrhs, double *B, int ldb) {
  const int nb = h->nb, P = h->P, ld = h->ld, nblk = (n + nb - 1) / nb;
  const int ldx = n + (n & 1);
  std::vector<double *> X(P, nullptr);
  std::vector<cudaEvent_t> ev(2 * nblk + 2, nullptr);        // each created on the device that records it
  int rc = 0;
  for (int q = 0; q < P; q++) { DevGuard g(h->d[q].dev); DCHK(cudaMalloc(&X[q], sizeof(double) * (size_t)ldx * std::max(nrhs, 1))); }
  int cur = 0;      // device holding the current right-hand sides
  {
    DevCtx &c = h->d[0];
    DevGuard g(c.dev);
    DCHK(cudaMemcpy2DAsync(X[0], sizeof(double) * ldx, B, sizeof(double) * ldb, sizeof(double) * n, nrhs, cudaMemcpyDefault, c.s));
    if (mode == 0) {    // B <- P B: the interchanges in order, as ?getrs does
      for (int k0 = 0; k0 < n && !rc; k0 += 8192) rc = laswp_device<double>(X[0], ldx, nrhs, c.ipiv, k0, std::min(n, k0 + 8192), c.s);
    }
    DCHK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    DCHK(cudaEventRecord(ev[0], c.s));
  }
  int evi = 0;
  auto hop = [&](int to) -> int {       // move the right-hand sides to device `to` (ordered after ev[evi])
    if (to == cur) return 0;
    DevCtx &c = h->d[to];
    DevGuard g(c.dev);
    DCHK(cudaStreamWaitEvent(c.s, ev[evi], 0));
    DCHK(cudaMemcpy2DAsync(X[to], sizeof(double) * ldx, X[cur], sizeof(double) * ldx, sizeof(double) * n, nrhs, cudaMemcpyDefault, c.s));
    cur = to;
    return 0;
  };
  for (int pass = 0; pass < 2 && !rc; pass++) {
    for (int t = 0; t < nblk && !rc; t++) {
      const int J = pass == 0 ? t : nblk - 1 - t;
      const int o = J % P, J0 = J * nb, jb = std::min(nb, n - J0);
      rc = hop(o);
      if (rc) break;
      DevCtx &c = h->d[o];
      DevGuard g(c.dev);
      const double *col = c.L + (size_t)(J / P) * nb * ld;          // block column J, all rows
      double *x = X[o];
      if (pass == 0) {           // forward with L: x_J <- L_JJ^{-1} x_J ; x_below -= L_below,J x_J
        rc = trsm_device<double>(141, 122, 111, mode == 0 ? 132 : 131, jb, nrhs, 1.0, col + J0, ld, x + J0, ldx, c.s);
        const int mr = n - J0 - jb;
        if (!rc && mr > 0) rc = gemm_device<double>(111, 111, mr, nrhs, jb, -1.0, col + J0 + jb, ld, x + J0, ldx, 1.0, x + J0 + jb, ldx, c.s, 0);
      } else if (mode == 0) {    // backward with U: x_J <- U_JJ^{-1} x_J ; x_above -= U_above,J x_J
        rc = trsm_device<double>(141, 121, 111, 131, jb, nrhs, 1.0, col + J0, ld, x + J0, ldx, c.s);
        if (!rc && J0 > 0) rc = gemm_device<double>(111, 111, J0, nrhs, jb, -1.0, col, ld, x + J0, ldx, 1.0, x, ldx, c.s, 0);
      } else {                   // backward with L^T: x_J <- L_JJ^{-T} (x_J - L_below,J^T x_below)
        const int mr = n - J0 - jb;
        if (mr > 0) rc = gemm_device<double>(112, 111, jb, nrhs, mr, -1.0, col + J0 + jb, ld, x + J0 + jb, ldx, 1.0, x + J0, ldx, c.s, 0);
        if (!rc) rc = trsm_device<double>(141, 122, 112, 131, jb, nrhs, 1.0, col + J0, ld, x + J0, ldx, c.s);
      }
      evi++;
      if (!rc) { DCHK(cudaEventCreateWithFlags(&ev[evi], cudaEventDisableTiming)); DCHK(cudaEventRecord(ev[evi], c.s)); }
    }
  }
  if (!rc) {
    DevCtx &c = h->d[cur];
    DevGuard g(c.dev);
    DCHK(cudaMemcpy2DAsync(B, sizeof(double) * ldb, X[cur], sizeof(double) * ldx, sizeof(double) * n, nrhs, cudaMemcpyDefault, c.s));
  }
  int r2 = sync_all(h);
  for (int q = 0; q < P; q++) { DevGuard g(h->d[q].dev); cudaFree(X[q]); }
  for (auto &e : ev) if (e) cudaEventDestroy(e);
  return rc ? rc : r2;
}

// ------------------------------------------------------------------------------------------------ Cholesky (lower storage)
static int dist_potrf(Dist *h, int n, int *info_out) {
  const int nb = h->nb, P = h->P, ld = h->ld, nblk = (n + nb - 1) / nb;
  std::vector<cudaEvent_t> ev_panel(nblk), ev_first(nblk);
  std::vector<std::vector<cudaEvent_t>> ev_recv(P, std::vector<cudaEvent_t>(nblk)), ev_upd(P, std::vector<cudaEvent_t>(nblk));
  for (int J = 0; J < nblk; J++) {
    { DevGuard g(h->d[J % P].dev); DCHK(cudaEventCreateWithFlags(&ev_panel[J], cudaEventDisableTiming)); }
    { DevGuard g(h->d[(J + 1) % P].dev); DCHK(cudaEventCreateWithFlags(&ev_first[J], cudaEventDisableTiming)); }   // recorded by the NEXT panel's owner
  }
  for (int q = 0; q < P; q++) {
    DevGuard g(h->d[q].dev);
    for (int J = 0; J < nblk; J++) {
      DCHK(cudaEventCreateWithFlags(&ev_recv[q][J], cudaEventDisableTiming));
      DCHK(cudaEventCreateWithFlags(&ev_upd[q][J], cudaEventDisableTiming));
    }
    DCHK(cudaMalloc(&h->d[q].info, sizeof(int) * 2));
    DCHK(cudaMemsetAsync(h->d[q].info, 0, sizeof(int) * 2, h->d[q].s));
  }
  int rc = 0;
  auto factor_panel = [&](int J, cudaStream_t st) -> int {       // diagonal block, then the rows below it
    DevCtx &c = h->d[J % P];
    DevGuard g(c.dev);
    const int J0 = J * nb, jb = std::min(nb, n - J0);
    double *blk = c.L + (size_t)(J / P) * nb * ld + J0;
    DCHK(cudaMemsetAsync(c.info + 1, 0, sizeof(int), st));
    if (int r = potrf_device<double>('L', jb, blk, ld, c.info + 1, st)) return r;
    first_info_kernel<<<1, 1, 0, st>>>(c.info + 1, J0, c.info);
    count_launch();
    const int mr = n - J0 - jb;
    if (mr > 0)
      if (int r = trsm_device<double>(142, 122, 112, 131, mr, jb, 1.0, blk, ld, blk + jb, ld, st)) return r;
    DCHK(cudaEventRecord(ev_panel[J], st));
    return 0;
  };
  // trailing update of local block column `b` (global block Jp > J) of device q: A[Jp0.., Jp] -= L[Jp0.., J] L[Jp, J]^T
  auto update_block = [&](int q, int J, const double *pan, int b, cudaStream_t st) -> int {
    DevCtx &c = h->d[q];
    const int Jp = b * P + q, Jp0 = Jp * nb, jbp = std::min(nb, n - Jp0), jb = std::min(nb, n - J * nb);
    double *dst = c.L + (size_t)b * nb * ld + Jp0;
    if (int r = gemm_device<double>(111, 112, jbp, jbp, jb, -1.0, pan + Jp0, ld, pan + Jp0, ld, 1.0, dst, ld, st, 'L')) return r;
    const int mr = n - Jp0 - jbp;
    if (mr > 0)
      if (int r = gemm_device<double>(111, 112, mr, jbp, jb, -1.0, pan + Jp0 + jbp, ld, pan + Jp0, ld, 1.0, dst + jbp, ld, st, 0)) return r;
    return 0;
  };
  rc = factor_panel(0, h->d[0].s);
  for (int J = 0; J < nblk && !rc; J++) {
    const int o = J % P, J0 = J * nb, jb = std::min(nb, n - J0);
    for (int q = 0; q < P && !rc; q++) {
      if (q == o) continue;
      DevCtx &c = h->d[q];
      DevGuard g(c.dev);
      DCHK(cudaStreamWaitEvent(c.sc, ev_panel[J], 0));
      if (J >= 2) DCHK(cudaStreamWaitEvent(c.sc, ev_upd[q][J - 2], 0));
      const double *src = h->d[o].L + (size_t)(J / P) * nb * ld + J0;
      DCHK(cudaMemcpy2DAsync(c.pbuf[J & 1] + J0, sizeof(double) * ld, src, sizeof(double) * ld, sizeof(double) * (n - J0), jb, cudaMemcpyDefault, c.sc));
      DCHK(cudaEventRecord(ev_recv[q][J], c.sc));
    }
    for (int q = 0; q < P && !rc; q++) {
      DevCtx &c = h->d[q];
      DevGuard g(c.dev);
      const int nl = nloc_blocks(nblk, P, q);
      const double *pan;
      if (q == o) { pan = c.L + (size_t)(J / P) * nb * ld; DCHK(cudaStreamWaitEvent(c.s, ev_panel[J], 0)); }
      else { pan = c.pbuf[J & 1]; DCHK(cudaStreamWaitEvent(c.s, ev_recv[q][J], 0)); }
      int b = (J >= q) ? (J - q) / P + 1 : 0;
      if (J + 1 < nblk && (J + 1) % P == q && b < nl) {
        rc = update_block(q, J, pan, b, c.s);
        if (!rc) {
          DCHK(cudaEventRecord(ev_first[J], c.s));
          DCHK(cudaStreamWaitEvent(c.s2, ev_first[J], 0));
          rc = factor_panel(J + 1, c.s2);
        }
        b++;
        tls().gemm_cta_cap = std::max(num_sms() - 24, 8);
      }
      for (; b < nl && !rc; b++) rc = update_block(q, J, pan, b, c.s);
      tls().gemm_cta_cap = 0;
      if (!rc) DCHK(cudaEventRecord(ev_upd[q][J], c.s));
    }
  }
  if (!rc) rc = sync_all(h);
  int info = 0;
  if (!rc)
    for (int q = 0; q < P; q++) {
      int hi[2] = {0, 0};
      DevGuard g(h->d[q].dev);
      DCHK(cudaMemcpy(hi, h->d[q].info, sizeof(hi), cudaMemcpyDeviceToHost));
      if (hi[0] != 0 && (info == 0 || hi[0] < info)) info = hi[0];
    }
  for (int J = 0; J < nblk; J++) { cudaEventDestroy(ev_panel[J]); cudaEventDestroy(ev_first[J]); }
  for (int q = 0; q < P; q++) {
    for (int J = 0; J < nblk; J++) { cudaEventDestroy(ev_recv[q][J]); cudaEventDestroy(ev_upd[q][J]); }
    DevGuard g(h->d[q].dev);
    cudaFree(h->d[q].info); h->d[q].info = nullptr;
  }
  *info_out = info;
  return rc;
}

// ------------------------------------------------------------------------------------------------ SUMMA dgemm
static int dist_dgemm(Dist *h, int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
                      double beta, double *C, int ldc) {
  const int pr = h->pr, pc = h->pc, P = h->P;
  // block edges: rows of C over pr, columns over pc, K over lcm-free granules g = K / (pr*pc) is not needed: K is cut
  // into pc blocks for A's ownership and pr blocks for B's; panels are the cells of the common refinement
  auto edge = [](int len, int parts, int i) { return (int)(((long long)len * i) / parts / 2 * 2); };      // even edges (TMA alignment)
  auto cut = [&](int len, int parts, int i) { return i >= parts ? len : edge(len, parts, i); };
  struct Blk { double *A = nullptr, *B = nullptr, *C = nullptr, *T = nullptr; int mb = 0, nbk = 0, lda = 0, ldb = 0, ldc = 0; };
  std::vector<Blk> blk(P);
  std::vector<int> kcut;                         // refinement of K: every edge of either partition
  for (int i = 0; i <= pc; i++) kcut.push_back(cut(k, pc, i));
  for (int i = 0; i <= pr; i++) kcut.push_back(cut(k, pr, i));
  std::sort(kcut.begin(), kcut.end());
  kcut.erase(std::unique(kcut.begin(), kcut.end()), kcut.end());
  const int npan = (int)kcut.size() - 1;
  std::vector<std::vector<cudaEvent_t>> ev_own(P, std::vector<cudaEvent_t>(2)), ev_pan(P, std::vector<cudaEvent_t>(std::max(npan, 1)));
  int rc = 0;
  const bool need_c = beta != 0.0;
  // ---- scatter: device (i,j) gets op(A)[rows i, K block j] and op(B)[K block i, cols j] (stored sub-blocks, transposed on
  //      the device when the caller's operand is), C block when beta != 0
  for (int q = 0; q < P && !rc; q++) {
    const int i = q / pc, j = q % pc;
    DevCtx &c = h->d[q];
    DevGuard g(c.dev);
    Blk &b = blk[q];
    b.mb = cut(m, pr, i + 1) - cut(m, pr, i); b.nbk = cut(n, pc, j + 1) - cut(n, pc, j);
    b.lda = b.mb + (b.mb & 1); b.ldb = k + (k & 1); b.ldc = b.lda;
    for (auto &e : ev_own[q]) DCHK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : ev_pan[q]) DCHK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (b.mb == 0 || b.nbk == 0) continue;
    DCHK(cudaMalloc(&b.A, sizeof(double) * (size_t)b.lda * std::max(k, 1)));
    DCHK(cudaMalloc(&b.B, sizeof(double) * (size_t)b.ldb * b.nbk));
    DCHK(cudaMalloc(&b.C, sizeof(double) * (size_t)b.ldc * b.nbk));
    const int r0 = cut(m, pr, i), c0 = cut(n, pc, j);
    const int ka0 = cut(k, pc, j), ka1 = cut(k, pc, j + 1), kb0 = cut(k, pr, i), kb1 = cut(k, pr, i + 1);
    const int tmax = std::max(std::max(b.mb, ka1 - ka0), std::max(kb1 - kb0, b.nbk));
    if (ta != 111 || tb != 111) DCHK(cudaMalloc(&b.T, sizeof(double) * (size_t)(tmax + 1) * (tmax + 1)));
    if (ka1 > ka0) {
      if (ta == 111) DCHK(cudaMemcpy2DAsync(b.A + (size_t)ka0 * b.lda, sizeof(double) * b.lda, A + r0 + (size_t)ka0 * lda, sizeof(double) * lda, sizeof(double) * b.mb, ka1 - ka0, cudaMemcpyDefault, c.s));
      else {       // stored (ka1-ka0) x mb
        const int ldt = (ka1 - ka0) + ((ka1 - ka0) & 1);
        DCHK(cudaMemcpy2DAsync(b.T, sizeof(double) * ldt, A + ka0 + (size_t)r0 * lda, sizeof(double) * lda, sizeof(double) * (ka1 - ka0), b.mb, cudaMemcpyDefault, c.s));
        rc = lacpy_device<double>(112, ka1 - ka0, b.mb, b.T, ldt, b.A + (size_t)ka0 * b.lda, b.lda, c.s);
      }
    }
    DCHK(cudaEventRecord(ev_own[q][0], c.s));
    if (!rc && kb1 > kb0) {
      if (tb == 111) DCHK(cudaMemcpy2DAsync(b.B + kb0, sizeof(double) * b.ldb, B + kb0 + (size_t)c0 * ldb, sizeof(double) * ldb, sizeof(double) * (kb1 - kb0), b.nbk, cudaMemcpyDefault, c.s));
      else {       // stored nbk x (kb1-kb0)
        const int ldt = b.nbk + (b.nbk & 1);
        DCHK(cudaMemcpy2DAsync(b.T, sizeof(double) * ldt, B + c0 + (size_t)kb0 * ldb, sizeof(double) * ldb, sizeof(double) * b.nbk, kb1 - kb0, cudaMemcpyDefault, c.s));
        rc = lacpy_device<double>(112, b.nbk, kb1 - kb0, b.T, ldt, b.B + kb0, b.ldb, c.s);
      }
    }
    if (!rc && need_c) DCHK(cudaMemcpy2DAsync(b.C, sizeof(double) * b.ldc, C + r0 + (size_t)c0 * ldc, sizeof(double) * ldc, sizeof(double) * b.mb, b.nbk, cudaMemcpyDefault, c.s));
    DCHK(cudaEventRecord(ev_own[q][1], c.s));
  }
  // ---- pulls: K panel t of A comes from the row peer that owns it, of B from the column peer (copy engines, NVLink)
  for (int q = 0; q < P && !rc; q++) {
    const int i = q / pc, j = q % pc;
    DevCtx &c = h->d[q];
    Blk &b = blk[q];
    if (b.mb == 0 || b.nbk == 0) continue;
    DevGuard g(c.dev);
    for (int t = 0; t < npan; t++) {
      const int k0 = kcut[t], k1 = kcut[t + 1], w = k1 - k0;
      if (w <= 0) continue;
      int ja = 0; while (ja + 1 < pc && cut(k, pc, ja + 1) <= k0) ja++;
      int ib = 0; while (ib + 1 < pr && cut(k, pr, ib + 1) <= k0) ib++;
      const int qa = i * pc + ja, qb = ib * pc + j;
      if (qa != q) {
        DCHK(cudaStreamWaitEvent(c.sc, ev_own[qa][0], 0));
        DCHK(cudaMemcpy2DAsync(b.A + (size_t)k0 * b.lda, sizeof(double) * b.lda, blk[qa].A + (size_t)k0 * blk[qa].lda, sizeof(double) * blk[qa].lda,
                               sizeof(double) * b.mb, w, cudaMemcpyDefault, c.sc));
      }
      if (qb != q) {
        DCHK(cudaStreamWaitEvent(c.sc, ev_own[qb][1], 0));
        DCHK(cudaMemcpy2DAsync(b.B + k0, sizeof(double) * b.ldb, blk[qb].B + k0, sizeof(double) * blk[qb].ldb, sizeof(double) * w, b.nbk, cudaMemcpyDefault, c.sc));
      }
      DCHK(cudaEventRecord(ev_pan[q][t], c.sc));
    }
  }
  // ---- products, panel by panel (owned panels first would save nothing here: the host copies dominate this entry point)
  for (int q = 0; q < P && !rc; q++) {
    DevCtx &c = h->d[q];
    Blk &b = blk[q];
    if (b.mb == 0 || b.nbk == 0) continue;
    DevGuard g(c.dev);
    const int i = q / pc, j = q % pc;
    DCHK(cudaStreamWaitEvent(c.s, ev_own[q][1], 0));
    bool first = true;
    for (int t = 0; t < npan && !rc; t++) {
      const int k0 = kcut[t], w = kcut[t + 1] - k0;
      if (w <= 0) continue;
      DCHK(cudaStreamWaitEvent(c.s, ev_pan[q][t], 0));
      rc = gemm_device<double>(111, 111, b.mb, b.nbk, w, alpha, b.A + (size_t)k0 * b.lda, b.lda, b.B + k0, b.ldb, first ? beta : 1.0, b.C, b.ldc, c.s, 0);
      first = false;
    }
    if (!rc && k == 0) rc = gemm_device<double>(111, 111, b.mb, b.nbk, 0, alpha, b.A, b.lda, b.B, b.ldb, beta, b.C, b.ldc, c.s, 0);
    if (!rc) DCHK(cudaMemcpy2DAsync(C + cut(m, pr, i) + (size_t)cut(n, pc, j) * ldc, sizeof(double) * ldc, b.C, sizeof(double) * b.ldc, sizeof(double) * b.mb, b.nbk, cudaMemcpyDefault, c.s));
  }
  int r2 = sync_all(h);
  for (int q = 0; q < P; q++) {
    DevGuard g(h->d[q].dev);
    for (auto &e : ev_own[q]) if (e) cudaEventDestroy(e);
    for (auto &e : ev_pan[q]) if (e) cudaEventDestroy(e);
    cudaFree(blk[q].A); cudaFree(blk[q].B); cudaFree(blk[q].C); cudaFree(blk[q].T);
  }
  return rc ? rc : r2;
}

}  // namespace jb

using namespace jb;
extern "C" {

int jb_dist_create(void **handle, int pr, int pc, int nb) {
  if (!handle || pr < 1 || pc < 1 || nb < 32 || (nb & 31)) { set_error(JB_ERR_ARG, "jb_dist_create: need pr, pc >= 1 and nb a multiple of 32"); return JB_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error(JB_ERR_CUDA, "no usable CUDA device; jalla_b200 has no CPU path"); return JB_ERR_CUDA; }
  if (pr * pc > ndev) { set_error(JB_ERR_ARG, "jb_dist_create: %d x %d grid needs %d devices, %d visible", pr, pc, pr * pc, ndev); return JB_ERR_ARG; }
  if (int rc = ensure_device()) return rc;
  Dist *h = new Dist;
  h->pr = pr; h->pc = pc; h->P = pr * pc; h->nb = nb;
  h->d.resize(h->P);
  int prev = 0;
  cudaGetDevice(&prev);
  for (int q = 0; q < h->P; q++) {
    DevCtx &c = h->d[q];
    c.dev = q;
    cudaSetDevice(q);
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithPriority(&c.s2, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c.sc, cudaStreamNonBlocking) != cudaSuccess) {
      cudaSetDevice(prev); set_error(JB_ERR_CUDA, "jb_dist_create: stream creation failed on device %d", q); delete h; return JB_ERR_CUDA;
    }
    for (int p = 0; p < h->P; p++) {
      if (p == q) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, q, p);
      if (!can) { cudaSetDevice(prev); set_error(JB_ERR_CUDA, "jb_dist_create: device %d cannot access device %d", q, p); delete h; return JB_ERR_CUDA; }
      cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaSetDevice(prev); set_error(JB_ERR_CUDA, "peer access %d -> %d: %s", q, p, cudaGetErrorString(e)); delete h; return JB_ERR_CUDA; }
      cudaGetLastError();
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, q) == cudaSuccess) { uint64_t thr = UINT64_MAX; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
  }
  cudaSetDevice(prev);
  *handle = h;
  return 0;
}
int jb_dist_destroy(void *handle) {
  Dist *h = (Dist *)handle;
  if (!h) return 0;
  free_resident(h);
  for (auto &c : h->d) { DevGuard g(c.dev); cudaStreamDestroy(c.s); cudaStreamDestroy(c.s2); cudaStreamDestroy(c.sc); }
  delete h;
  return 0;
}
int jb_dist_ngpus(void *handle) { return handle ? ((Dist *)handle)->P : 0; }
// wall-clock phases of the last jb_dist_dgetrf / dpotrf: 0 host -> devices, 1 factorisation (all devices), 2 devices -> host
double jb_dist_last_ms(void *handle, int phase) { return (handle && phase >= 0 && phase < 3) ? ((Dist *)handle)->ms[phase] : -1.0; }

int jb_dist_dgemm(void *handle, int transa, int transb, int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb,
                  double beta, double *c, int ldc) {
  Dist *h = (Dist *)handle;
  if (!h || transa < 111 || transa > 113 || transb < 111 || transb > 113 || m < 0 || n < 0 || k < 0 || lda < std::max(1, transa == 111 ? m : k) ||
      ldb < std::max(1, transb == 111 ? k : n) || ldc < std::max(1, m)) { set_error(JB_ERR_ARG, "jb_dist_dgemm: bad argument"); return JB_ERR_ARG; }
  if (m == 0 || n == 0) return 0;
  return dist_dgemm(h, transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}

// LAPACKE_dgetrf semantics (column-major, square): A <- P L U, ipiv 1-based, returns info
int jb_dist_dgetrf(void *handle, int n, double *a, int lda, int *ipiv) {
  Dist *h = (Dist *)handle;
  if (!h) return -1;
  if (n < 0) return -2;
  if (lda < std::max(1, n)) return -4;
  if (n == 0) return 0;
  int info = 0;
  if (int rc = dist_getrf(h, n, a, lda, ipiv, &info)) return rc;
  return info;
}
int jb_dist_dgetrs(void *handle, char trans, int n, int nrhs, const double *a, int lda, const int *ipiv, double *b, int ldb) {
  Dist *h = (Dist *)handle;
  if (!h) return -1;
  if (!(trans == 'N' || trans == 'n')) { set_error(JB_ERR_ARG, "jb_dist_dgetrs: only trans = 'N' is distributed"); return -2; }
  if (n < 0) return -3;
  if (nrhs < 0) return -4;
  if (lda < std::max(1, n)) return -6;
  if (ldb < std::max(1, n)) return -9;
  if (n == 0 || nrhs == 0) return 0;
  if (!(h->kind == 1 && h->host_a == a && h->n == n)) {      // factors not resident: bring them (and the pivots) in
    if (int rc = alloc_resident(h, n)) return rc;
    if (int rc = scatter_cols(h, a, lda, n)) return rc;
    for (auto &c : h->d) { DevGuard g(c.dev); DCHK(cudaMemcpyAsync(c.ipiv, ipiv, sizeof(int) * n, cudaMemcpyDefault, c.s)); }
    if (int rc = sync_all(h)) return rc;
    h->kind = 1; h->host_a = a;
  }
  return dist_solve(h, 0, n, nrhs, b, ldb);
}
// LAPACKE_dpotrf / dpotrs semantics, column-major, uplo = 'L' (the upper form is the transposed problem: pass the
// row-major view, as LAPACKE itself does)
int jb_dist_dpotrf(void *handle, char uplo, int n, double *a, int lda) {
  Dist *h = (Dist *)handle;
  if (!h) return -1;
  if (!(uplo == 'L' || uplo == 'l')) { set_error(JB_ERR_ARG, "jb_dist_dpotrf: only uplo = 'L' is distributed"); return -2; }
  if (n < 0) return -3;
  if (lda < std::max(1, n)) return -5;
  if (n == 0) return 0;
  if (int rc = alloc_resident(h, n)) return rc;
  double t0 = now_ms();
  if (int rc = scatter_cols(h, a, lda, n)) return rc;
  if (int rc = sync_all(h)) return rc;
  h->ms[0] = now_ms() - t0; t0 = now_ms();
  int info = 0;
  if (int rc = dist_potrf(h, n, &info)) { free_resident(h); return rc; }
  h->ms[1] = now_ms() - t0; t0 = now_ms();
  // the device columns still hold the caller's strictly upper triangle (never written), so whole columns go back
  if (int rc = gather_cols(h, a, lda, n, 0)) return rc;
  if (int rc = sync_all(h)) return rc;
  h->ms[2] = now_ms() - t0;
  h->kind = 2; h->host_a = a;
  return info;
}
int jb_dist_dpotrs(void *handle, char uplo, int n, int nrhs, const double *a, int lda, double *b, int ldb) {
  Dist *h = (Dist *)handle;
  if (!h) return -1;
  if (!(uplo == 'L' || uplo == 'l')) { set_error(JB_ERR_ARG, "jb_dist_dpotrs: only uplo = 'L' is distributed"); return -2; }
  if (n < 0) return -3;
  if (nrhs < 0) return -4;
  if (lda < std::max(1, n)) return -6;
  if (ldb < std::max(1, n)) return -8;
  if (n == 0 || nrhs == 0) return 0;
  if (!(h->kind == 2 && h->host_a == a && h->n == n)) {
    if (int rc = alloc_resident(h, n)) return rc;
    if (int rc = scatter_cols(h, a, lda, n)) return rc;
    if (int rc = sync_all(h)) return rc;
    h->kind = 2; h->host_a = a;
  }
  return dist_solve(h, 2, n, nrhs, b, ldb);
}

// batch slices per device, no communication; host or UVA pointers, LAPACK-style per-system info
int jb_dist_dgesv_batched(void *handle, int n, int nrhs, double *a, int lda, long long stride_a, int *ipiv, double *b, int ldb, long long stride_b,
                          int *info, int batch) {
  Dist *h = (Dist *)handle;
  if (!h || n < 1 || n > 32 || nrhs < 0 || lda < n || ldb < n || batch < 0 || stride_a < (long long)lda * (n - 1) + n) { set_error(JB_ERR_ARG, "jb_dist_dgesv_batched: bad argument"); return JB_ERR_ARG; }
  if (batch == 0) return 0;
  const int P = h->P;
  struct Buf { double *A = nullptr, *B = nullptr; int *piv = nullptr, *info = nullptr; int lo = 0, cnt = 0; };
  std::vector<Buf> buf(P);
  int rc = 0;
  int prev = 0;
  cudaGetDevice(&prev);
  const cudaStream_t saved = tls().stream;
  const long long sb_elems = (long long)ldb * std::max(nrhs, 1);
  for (int q = 0; q < P && !rc; q++) {
    Buf &u = buf[q];
    u.lo = (int)((long long)batch * q / P); u.cnt = (int)((long long)batch * (q + 1) / P) - u.lo;
    if (u.cnt == 0) continue;
    DevCtx &c = h->d[q];
    cudaSetDevice(c.dev);
    const size_t abytes = sizeof(double) * (size_t)stride_a * u.cnt, bbytes = sizeof(double) * (size_t)stride_b * u.cnt;
    if (cudaMalloc(&u.A, abytes) != cudaSuccess || cudaMalloc(&u.B, std::max(bbytes, (size_t)16)) != cudaSuccess ||
        cudaMalloc(&u.piv, sizeof(int) * (size_t)n * u.cnt) != cudaSuccess || cudaMalloc(&u.info, sizeof(int) * (size_t)u.cnt) != cudaSuccess) { rc = JB_ERR_ALLOC; break; }
    cudaMemcpyAsync(u.A, a + (size_t)u.lo * stride_a, abytes, cudaMemcpyDefault, c.s);
    if (nrhs > 0) cudaMemcpyAsync(u.B, b + (size_t)u.lo * stride_b, bbytes, cudaMemcpyDefault, c.s);
    tls().stream = c.s; tls().device = c.dev;      // the batched entry points launch on the calling thread's stream / device
    if (nrhs > 0) rc = jb_dgesv_batched(n, nrhs, u.A, lda, stride_a, u.piv, u.B, ldb, stride_b, u.info, u.cnt);
    else rc = jb_dgetrf_batched(n, u.A, lda, stride_a, u.piv, u.info, u.cnt);
    cudaSetDevice(c.dev);    // (the entry point re-binds the library's default device)
    cudaMemcpyAsync(a + (size_t)u.lo * stride_a, u.A, abytes, cudaMemcpyDefault, c.s);
    if (nrhs > 0) cudaMemcpyAsync(b + (size_t)u.lo * stride_b, u.B, bbytes, cudaMemcpyDefault, c.s);
    cudaMemcpyAsync(ipiv + (size_t)u.lo * n, u.piv, sizeof(int) * (size_t)n * u.cnt, cudaMemcpyDefault, c.s);
    cudaMemcpyAsync(info + u.lo, u.info, sizeof(int) * (size_t)u.cnt, cudaMemcpyDefault, c.s);
  }
  (void)sb_elems;
  tls().stream = saved; tls().device = -1;
  for (int q = 0; q < P; q++) {
    DevCtx &c = h->d[q];
    cudaSetDevice(c.dev);
    if (cudaStreamSynchronize(c.s) != cudaSuccess && !rc) { rc = JB_ERR_CUDA; set_error(JB_ERR_CUDA, "jb_dist_dgesv_batched: device %d failed", q); }
    cudaFree(buf[q].A); cudaFree(buf[q].B); cudaFree(buf[q].piv); cudaFree(buf[q].info);
  }
  cudaSetDevice(prev);
  return rc;
}
}
