// diffiqult_b200 — host-callable launchers of the kernels in dq_kernels.cu (all asynchronous on `st`).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <map>
#include <utility>

#include "dq_layout.h"

namespace dq {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: remember what has been granted per
// (kernel ADDRESS, device) -- template instantiations of one kernel share a function type, so the type cannot be the key --
// so that one process driving several GPUs (one host thread each) opts in on every one of them.
template <class K>
inline void ensure_dynamic_smem(K kernel, size_t bytes, bool prefer_shared_carveout = false) {
  if (bytes <= 48 * 1024 && !prefer_shared_carveout) return;
  int dev = 0;
  cudaGetDevice(&dev);
  static thread_local std::map<std::pair<const void*, int>, size_t> granted;
  size_t& g = granted[{(const void*)kernel, dev}];
  if (bytes > g) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (prefer_shared_carveout) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    g = bytes;
  }
}

int upload_boys_constants();
int* boys_fail_flag();          // device int of the current device: set when a Boys loop ran out of terms
long long launch_count();
void reset_launch_count();

void launch_normalize(cudaStream_t st, int N, const int* f_off, const int* f_type, const double* alpha, const double* craw, double* cn);
void launch_normalize_bwd(cudaStream_t st, int N, const int* f_off, const int* f_type, const double* alpha, const double* craw,
                          const double* cbar, double* abar, double* crawbar);
void launch_pairs(cudaStream_t st, const SystemDev& s, double* pd, double* bp_kmax);
void launch_one_electron(cudaStream_t st, const SystemDev& s, double* S, double* T, double* V);
void launch_one_electron_grad(cudaStream_t st, const SystemDev& s, int kmax, const double* Sbar, const double* Hbar, double* g_alpha,
                              double* g_coef, double* g_xyz);
// sum_ij Hbar_ij dV_ij / d(nuclear coordinates) -> g_atoms[natoms*3] (accumulators: limb pairs when s.det)
void launch_nuclear_grad(cudaStream_t st, const SystemDev& s, int kmax, const double* Hbar, double* g_atoms);
void launch_eri_tiles(cudaStream_t st, int cls, const Tile* tiles, int nt, const SystemDev& s, const double* pd, double* store, int kkmax);
// the same tiles with the surviving primitive pairs compacted first (systems with many underflowed Gaussian products)
void launch_eri_tiles_compact(cudaStream_t st, int cls, const Tile* tiles, int nt, const SystemDev& s, const double* pd, double* store,
                              int kkmax);
void launch_eri_grad_chunks(cudaStream_t st, int cls, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s,
                            const double* pd, int nterms, const double* At, const double* Bt, double* padj, double* pasum, int kkmax,
                            unsigned* dmask = nullptr, unsigned char* dbin = nullptr, int tile0 = 0);
void launch_eri_grad_deferred(cudaStream_t st, int cls, const Tile* tiles, int tile0, int nt, const SystemDev& s, const double* pd,
                              int nterms, const double* At, const double* Bt, double* padj, double* pasum, const unsigned* dmask,
                              const unsigned char* dbin, int kkmax);
// sparse-weight variant (chunks of <= 32 tiles) and the nonzero count the host chooses by
void launch_eri_grad_sparse(cudaStream_t st, int cls, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s,
                            const double* pd, int nterms, const double* At, const double* Bt, double* padj, double* pasum, int kkmax);
void launch_count_nonzero(cudaStream_t st, int n2, const double* A, int* out);
void launch_weight_mask(cudaStream_t st, int n2, int nterms, const double* Bt, unsigned char* mask);
size_t eri_grad_smem_bytes(int kkmax);
void launch_scatter_packed(cudaStream_t st, const Tile* tiles, int nt, const SystemDev& s, const int* f_orig, const double* store,
                           double* packed);
void launch_jk_chunks(cudaStream_t st, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* store,
                      const double* D, double* Jacc, double* Kacc);
// fixed-point accumulators (pairs of 64-bit limbs) -> doubles: out[i] = beta * out[i] + value
void launch_det_finalize(cudaStream_t st, long long n, const double* in, double* out, double beta);
// integral-direct J/K: the tiles of each chunk (same bra block pair) are evaluated and contracted with D inside one kernel
size_t jk_direct_smem_bytes(int kkmax, int N, int det);
void launch_jk_direct(cudaStream_t st, int cls, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* pd,
                      const double* D, double* Jacc, double* Kacc, int kkmax);
void launch_pair_bwd(cudaStream_t st, const SystemDev& s, const double* padj, const double* pasum, double* g_alpha, double* g_coef,
                     double* g_xyz);
void launch_gemm(cudaStream_t st, bool ta, bool tb, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                 int ldc, double alpha, double beta);
size_t eigh_workspace_bytes(int n, int batch);
// `work`: eigh_workspace_bytes(n, batch) bytes of device memory (NULL: slow global-memory fallback)
void launch_eigh(cudaStream_t st, int n, int batch, double* A, double* Vwork, double* Vout, double* w, void* work);
// cluster one-sided Jacobi (48 <= n <= 236): A symmetric n x n (not modified), V0 = NULL or an orthogonal warm start
bool eigh_cluster_applicable(int n);
void launch_eigh_cluster(cudaStream_t st, int n, int batch, const double* A, const double* V0, double* Vout, double* w, void* work);
void launch_scale_cols_rsqrt(cudaStream_t st, int n, const double* A, const double* w, double* B);
void launch_fock_combine(cudaStream_t st, int n, const int* f_block, const double* H, const double* Jacc, const double* Kacc, double* F,
                         double* G);
void launch_dot3(cudaStream_t st, int n, const double* A, const double* B, const double* C, double* out);
void launch_make_y(cudaStream_t st, int n, int ne, const double* Z, const double* eps, double* Y);
void launch_phi_rsqrt(cudaStream_t st, int n, const double* w, double* M);
void launch_axpby(cudaStream_t st, int n, double a, const double* X, double b, double* Y);
void launch_symmetrize(cudaStream_t st, int n, double* A);
void launch_add_transpose(cudaStream_t st, int n, const double* A, double* B);
void launch_exp(cudaStream_t st, int n, const double* x, double* y);
void launch_mul(cudaStream_t st, int n, const double* a, const double* b, double* y);
void launch_boys(cudaStream_t st, int n, const double* t, double* F, double* dF);
void launch_fp64_peak(cudaStream_t st, int blocks, int iters, double* out);

// batched small-molecule path (dq_batch.cu)
int upload_boys_constants_batch(int* fail_flag);
long long batch_launch_count();
void reset_batch_launch_count();
void launch_batch_prepare(cudaStream_t st, const BatchDev& b, const double* alpha_in, int log_alpha);
void launch_batch_one_electron(cudaStream_t st, const BatchDev& b, double* S, double* H);
void launch_batch_eri(cudaStream_t st, int cls, const BatchDev& b, const int2* qlist, int nq, double* eri);
size_t batch_scf_smem_bytes(int n, int neri, bool eri_in_smem);
int launch_batch_scf(cudaStream_t st, const BatchDev& b, const double* S, const double* H, const double* eri, int max_scf, double tol,
                     int want_grad, double* hist, double* Eelec, int* niter, int* conv, double* esteps, double* C, double* eps,
                     double* terms, int* nterms, double* Hbar, double* Sbar);
void launch_batch_eri_grad(cudaStream_t st, int cls, const BatchDev& b, const int2* qlist, int nq, const double* terms, const int* nterms,
                           int max_scf, double* padj, double* pasum);
void launch_batch_grad_accumulate(cudaStream_t st, const BatchDev& b, const double* padj, const double* pasum, const double* Sbar,
                                  const double* Hbar, double* ga, double* gc, double* gx);
void launch_batch_grad_finish(cudaStream_t st, const BatchDev& b, int log_alpha, double* ga, const double* gc, double* graw);

}  // namespace dq
