// diffiqult_b200 — batched path for MANY SMALL molecules of one structure (BASELINE.json config 5: 1024 perturbed
// CH4 / STO-3G; SURVEY.md 8e "pure replicas").  Same mathematics as the single-molecule path (dq_math.cuh, the
// reference's formulas and its truncated SCF), different mapping: for N ~ 10 basis functions a 256-quartet tile is
// almost empty and one molecule cannot fill a GPU, so here
//   * every (unique contracted quartet, molecule) is ONE THREAD, neighbouring lanes = neighbouring molecules, so a warp
//     never diverges on the integral class and all per-molecule arrays consumed by these kernels are molecule-fastest,
//   * the whole SCF of a molecule (S^-1/2, Roothaan iterations with the reference's stopping rule, Energy.py:138-195)
//     AND its exact reverse sweep run inside ONE CTA with every matrix in shared memory: one launch
//     for the whole batch instead of ~40 launches per molecule and iteration.
// Kernels: kb_pairs, kb_one_electron, kb_eri<class>, kb_scf, kb_eri_grad<class>, kb_pair_bwd, kb_one_electron_grad,
// kb_finish.  Molecules are independent: a rank of a multi-GPU job simply gets a slice of the batch.
#include <cuda_runtime.h>
#include <stdio.h>

#include "dq_layout.h"
#include "dq_launch.h"
#include "dq_math.cuh"

namespace dq {

__constant__ BoysConst c_boys_b;

int upload_boys_constants_batch(int* fail_flag) {
  BoysConst h;
  boys_const_init(&h);
  h.fail = fail_flag;
  return (int)cudaMemcpyToSymbol(c_boys_b, &h, sizeof(h));
}

static thread_local long long g_blaunches = 0;
long long batch_launch_count() { return g_blaunches; }
void reset_batch_launch_count() { g_blaunches = 0; }
#define DQB_LAUNCHED() (++g_blaunches)

__device__ __forceinline__ int tri_idx(int x, int y, int m) { return x * m - ((x * (x + 1)) >> 1) + y; }   // x <= y (Tools.py:17-22)

// ---------------------------------------------------------------------------------------------
// exponents (exp of ln alpha), normalisation (Integrals.py:477-501): one thread per (function, molecule)
// ---------------------------------------------------------------------------------------------
__global__ void kb_prepare(BatchDev b, const double* __restrict__ alpha_in, int log_alpha) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= b.N * b.nmol) return;
  const int f = id / b.nmol, mol = id - f * b.nmol;
  const int L = b.f_type[f] > 0 ? 1 : 0;
  const double pw = 0.5 * L + 0.75, e = L + 1.5;
  const double fac = pow(2.0 / kPi, 0.75) * (L ? 2.0 : 1.0);
  const double div = pow(kPi, 1.5) / (L ? 2.0 : 1.0);
  const int p0 = b.f_off[f], p1 = b.f_off[f + 1];
  double* al = b.alpha + (size_t)mol * b.nprim;
  const double* cr = b.craw + (size_t)mol * b.nprim;
  double* cn = b.cn + (size_t)mol * b.nprim;
  for (int p = p0; p < p1; ++p) { const double a = alpha_in[(size_t)mol * b.nprim + p]; al[p] = log_alpha ? exp(a) : a; }
  double tm = 0.0;
  for (int p = p0; p < p1; ++p) {
    const double up = fac * pow(al[p], pw) * cr[p];
    for (int q = p0; q < p1; ++q) tm += fac * pow(al[q], pw) * cr[q] * up / pow(al[p] + al[q], e);
  }
  const double nrm = sqrt(tm * div);
  for (int p = p0; p < p1; ++p) cn[p] = fac * pow(al[p], pw) * cr[p] / nrm;
}

// primitive-pair records, molecule-fastest: prec[(r*8 + field)*nmol + mol]; one thread per (record, molecule)
__global__ void kb_pairs(BatchDev b) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)b.nrec * b.nmol) return;
  const int r = (int)(id / b.nmol), mol = (int)(id - (long long)r * b.nmol);
  const int x = b.rec_pair[r], pp = r - b.pair_off[x];
  const int i = b.pairs[x].x, j = b.pairs[x].y;
  const int kj = b.f_off[j + 1] - b.f_off[j];
  const int p = b.f_off[i] + pp / kj, q = b.f_off[j] + pp % kj;
  const double* al = b.alpha + (size_t)mol * b.nprim;
  const double* cn = b.cn + (size_t)mol * b.nprim;
  const double* xyz = b.xyz + (size_t)mol * b.N * 3;
  const PairPrim pr = make_pair(al[p], cn[p], xyz + 3 * i, b.f_type[i], al[q], cn[q], xyz + 3 * j, b.f_type[j]);
  double* o = b.prec + (size_t)r * kPairFields * b.nmol + mol;
  const size_t st = b.nmol;
  o[0] = pr.P[0]; o[st] = pr.P[1]; o[2 * st] = pr.P[2]; o[3 * st] = pr.g; o[4 * st] = pr.ginv; o[5 * st] = pr.K; o[6 * st] = pr.pa;
  o[7 * st] = pr.pb;
}

__device__ __forceinline__ PairPrim load_pair_b(const double* __restrict__ rec, size_t st) {
  PairPrim r;
  r.P[0] = rec[0]; r.P[1] = rec[st]; r.P[2] = rec[2 * st]; r.g = rec[3 * st]; r.ginv = rec[4 * st]; r.K = rec[5 * st];
  r.pa = rec[6 * st]; r.pb = rec[7 * st];
  return r;
}

// S, T, V -> S and H = T + V per molecule ([mol][N*N], symmetric fill): one thread per (function pair, molecule)
__global__ void kb_one_electron(BatchDev b, double* __restrict__ S, double* __restrict__ H) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= b.M * b.nmol) return;
  const int x = id / b.nmol, mol = id - x * b.nmol;
  const int i = b.pairs[x].x, j = b.pairs[x].y, n = b.N;
  const double* al = b.alpha + (size_t)mol * b.nprim;
  const double* cn = b.cn + (size_t)mol * b.nprim;
  const double* xyz = b.xyz + (size_t)mol * n * 3;
  const double* C = b.atoms + (size_t)mol * b.natoms * 3;
  double sv = 0.0, tv = 0.0, vv = 0.0;
  for (int p = b.f_off[i]; p < b.f_off[i + 1]; ++p)
    for (int q = b.f_off[j]; q < b.f_off[j + 1]; ++q) {
      double s1, t1, v1;
      one_electron_primitive<double>(al[p], cn[p], xyz + 3 * i, b.f_type[i], al[q], cn[q], xyz + 3 * j, b.f_type[j], b.natoms,
                                     b.Z, C, c_boys_b, &s1, &t1, &v1);
      sv += s1; tv += t1; vv += v1;
    }
  double* Sm = S + (size_t)mol * n * n;
  double* Hm = H + (size_t)mol * n * n;
  Sm[i * n + j] = Sm[j * n + i] = sv;
  Hm[i * n + j] = Hm[j * n + i] = tv + vv;
}

// ---------------------------------------------------------------------------------------------
// contracted ERIs (Integrals.py:190-401), packed in the reference's order per molecule: eri[mol][eri_index]
// ---------------------------------------------------------------------------------------------
template <bool LA, bool LB, bool LC, bool LD>
__global__ void __launch_bounds__(128) kb_eri(BatchDev b, const int2* __restrict__ qlist, int nq, double* __restrict__ eri) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)nq * b.nmol) return;
  const int qi = (int)(id / b.nmol), mol = (int)(id - (long long)qi * b.nmol);
  const int2 q = qlist[qi];
  const int i = b.pairs[q.x].x, j = b.pairs[q.x].y, k = b.pairs[q.y].x, l = b.pairs[q.y].y;
  const int ia = b.f_type[i] - 1, ib = b.f_type[j] - 1, ic = b.f_type[k] - 1, id_ = b.f_type[l] - 1;
  const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id_);
  const size_t st = b.nmol;
  const int kkb = b.pair_off[q.x + 1] - b.pair_off[q.x], kkk = b.pair_off[q.y + 1] - b.pair_off[q.y];
  const double* brec0 = b.prec + (size_t)b.pair_off[q.x] * kPairFields * st + mol;
  const double* krec0 = b.prec + (size_t)b.pair_off[q.y] * kPairFields * st + mol;
  double v = 0.0;
  for (int pb = 0; pb < kkb; ++pb) {
    const double* brec = brec0 + (size_t)pb * kPairFields * st;
    const PairPrim bra = load_pair_b(brec, st);
    double bs[4] = {0.0, 0.0, 0.0, 0.0};
    if (LA) bs[0] = brec[ia * st];
    if (LB) bs[1] = brec[ib * st];
    if (LC) bs[2] = brec[ic * st];
    if (LD) bs[3] = brec[id_ * st];
    for (int pk = 0; pk < kkk; ++pk) {
      const double* krec = krec0 + (size_t)pk * kPairFields * st;
      const PairPrim ket = load_pair_b(krec, st);
      double ds[4] = {0.0, 0.0, 0.0, 0.0};
      if (LA) ds[0] = bs[0] - krec[ia * st];
      if (LB) ds[1] = bs[1] - krec[ib * st];
      if (LC) ds[2] = bs[2] - krec[ic * st];
      if (LD) ds[3] = bs[3] - krec[id_ * st];
      v += prim_quartet<LA, LB, LC, LD, false>(bra, ket, ds, dl, c_boys_b, 0.0, nullptr);
    }
  }
  eri[(size_t)mol * b.neri + tri_idx(q.x, q.y, b.M)] = v;
}

// ---------------------------------------------------------------------------------------------
// the SCF of one molecule and its reverse sweep, one CTA, matrices in shared memory
// ---------------------------------------------------------------------------------------------
struct Cta {   // cooperative helpers over n x n row-major matrices in shared memory; every call ends with a barrier
  int n, tid, nt;
  __device__ void sync() const { __syncthreads(); }
  // C = op(A) op(B) with inner dimension kk (<= n): used with kk = ne for the occupied projector
  __device__ void mm(const double* A, bool ta, const double* B, bool tb, double* C, int kk) const {
    for (int x = tid; x < n * n; x += nt) {
      const int r = x / n, c = x - r * n;
      double acc = 0.0;
      for (int k = 0; k < kk; ++k) acc += (ta ? A[k * n + r] : A[r * n + k]) * (tb ? B[c * n + k] : B[k * n + c]);
      C[x] = acc;
    }
    sync();
  }
  __device__ void copy(const double* A, double* B, int cnt) const { for (int x = tid; x < cnt; x += nt) B[x] = A[x]; sync(); }
  __device__ void axpy(double a, const double* X, double* Y) const { for (int x = tid; x < n * n; x += nt) Y[x] += a * X[x]; sync(); }
  __device__ void add_transpose(const double* A, double* B) const {   // B += A + A^T
    for (int x = tid; x < n * n; x += nt) { const int r = x / n, c = x - r * n; B[x] += A[x] + A[c * n + r]; }
    sync();
  }
  __device__ void symmetrize(double* A) const {
    for (int x = tid; x < n * n; x += nt) { const int r = x / n, c = x - r * n; if (r < c) { const double v = 0.5 * (A[x] + A[c * n + r]); A[x] = v; A[c * n + r] = v; } }
    sync();
  }
};

// parallel cyclic Jacobi (round-robin pairing) on a symmetric n x n matrix in shared memory.  A is destroyed; on return
// w[] holds the eigenvalues ascending and V their eigenvectors in columns (Eigen.py:69-74: numpy eigh ordering).
// work: n*n doubles + 2*m doubles + 2*m ints
__device__ void jacobi_smem(const Cta& c, double* A, double* V, double* w, double* work) {
  const int n = c.n, m = (n + 1) / 2, np = 2 * m, tid = c.tid, nt = c.nt;
  double* Vt = work;
  double* cs = work + n * n;
  double* sn = cs + m;
  int* pp = (int*)(sn + m);
  int* qq = pp + m;
  __shared__ double s_off, s_tot;
  for (int x = tid; x < n * n; x += nt) Vt[x] = (x / n == x % n) ? 1.0 : 0.0;
  c.sync();
  for (int sweep = 0; sweep < 40; ++sweep) {
    if (tid == 0) {   // n <= 32: a serial norm is negligible next to the rotations and keeps the test deterministic
      double off = 0.0, dg = 0.0;
      for (int x = 0; x < n * n; ++x) { const double a = A[x]; if (x / n == x % n) dg += a * a; else off += a * a; }
      s_off = off; s_tot = off + dg;
    }
    c.sync();
    if (s_off <= 1e-30 * s_tot) break;
    for (int round = 0; round < np - 1; ++round) {
      if (tid < m) {
        const int a = (tid == 0) ? np - 1 : (round + tid) % (np - 1);
        const int bq = (tid == 0) ? round : (round + np - 1 - tid) % (np - 1);
        const int p = a < bq ? a : bq;
        int q = a < bq ? bq : a;
        double cc = 1.0, ss = 0.0;
        if (q < n) {
          const double apq = A[p * n + q], app = A[p * n + p], aqq = A[q * n + q];
          if (fabs(apq) > 1e-18 * (fabs(app) + fabs(aqq)) && fabs(apq) > 1e-300) {
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            cc = 1.0 / sqrt(t * t + 1.0); ss = t * cc;
          }
        } else { q = -1; }
        cs[tid] = cc; sn[tid] = ss; pp[tid] = p; qq[tid] = q;
      }
      c.sync();
      for (int x = tid; x < m * n; x += nt) {        // columns: A <- A J, V <- V J
        const int r = x / m, k = x - r * m, q = qq[k];
        if (q < 0 || sn[k] == 0.0) continue;
        const int p = pp[k]; const double cc = cs[k], ss = sn[k];
        const double arp = A[r * n + p], arq = A[r * n + q];
        A[r * n + p] = cc * arp - ss * arq; A[r * n + q] = ss * arp + cc * arq;
        const double vrp = Vt[r * n + p], vrq = Vt[r * n + q];
        Vt[r * n + p] = cc * vrp - ss * vrq; Vt[r * n + q] = ss * vrp + cc * vrq;
      }
      c.sync();
      for (int x = tid; x < m * n; x += nt) {        // rows: A <- J^T A
        const int k = x / n, col = x - k * n, q = qq[k];
        if (q < 0 || sn[k] == 0.0) continue;
        const int p = pp[k]; const double cc = cs[k], ss = sn[k];
        const double apc = A[p * n + col], aqc = A[q * n + col];
        A[p * n + col] = cc * apc - ss * aqc; A[q * n + col] = ss * apc + cc * aqc;
      }
      c.sync();
    }
  }
  for (int i = tid; i < n; i += nt) {
    const double wi = A[i * n + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) { const double wj = A[j * n + j]; rank += (wj < wi) || (wj == wi && j < i); }
    w[rank] = wi;
    for (int r = 0; r < n; ++r) V[r * n + rank] = Vt[r * n + i];
  }
  c.sync();
}

// G[D]_ij = sum_kl D_kl [2 (ij|kl) - (ik|jl)]  (Energy.py:29-38, gather form: deterministic); F = H + G if H != null
__device__ void fock_smem(const Cta& c, const int* pidx, int M, const double* eri, const double* D, const double* H, double* F) {
  const int n = c.n;
  for (int x = c.tid; x < n * n; x += c.nt) {
    const int i = x / n, j = x - i * n;
    if (i > j) continue;
    const int xij = pidx[x];
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
      const int xik = pidx[i * n + k];
      for (int l = 0; l < n; ++l) {
        const int xkl = pidx[k * n + l], xjl = pidx[j * n + l];
        const double vj = eri[xij <= xkl ? tri_idx(xij, xkl, M) : tri_idx(xkl, xij, M)];
        const double vk = eri[xik <= xjl ? tri_idx(xik, xjl, M) : tri_idx(xjl, xik, M)];
        acc += D[k * n + l] * (2.0 * vj - vk);
      }
    }
    const double f = (H ? H[x] : 0.0) + acc;
    F[x] = f; F[j * n + i] = f;
  }
  c.sync();
}

// fixed-order sum_x A (B + C) by one thread (n*n <= 1024 terms): deterministic
__device__ double dot3_smem(const Cta& c, const double* A, const double* B, const double* Cm, double* slot) {
  if (c.tid == 0) { double acc = 0.0; for (int x = 0; x < c.n * c.n; ++x) acc += A[x] * (B[x] + Cm[x]); *slot = acc; }
  c.sync();
  return *slot;
}

// hist layout per molecule and iteration t (0-based): [Fprev | U | D | F] (4 n^2) + eps (n)
__global__ void __launch_bounds__(128) kb_scf(BatchDev b, const double* __restrict__ Sall, const double* __restrict__ Hall,
                                              const double* __restrict__ eriall, int eri_in_smem, int max_scf, double tol, int want_grad,
                                              double* __restrict__ hist, double* __restrict__ Eelec, int* __restrict__ niter,
                                              int* __restrict__ conv, double* __restrict__ esteps, double* __restrict__ Cout,
                                              double* __restrict__ epsout, double* __restrict__ terms, int* __restrict__ nterms,
                                              double* __restrict__ Hbar_out, double* __restrict__ Sbar_out) {
  extern __shared__ double sm[];
  const int mol = blockIdx.x, n = b.N, nn = n * n, ne = b.ne, M = b.M;
  Cta c; c.n = n; c.tid = threadIdx.x; c.nt = blockDim.x;
  double* p = sm;
  double *S = p; p += nn; double *H = p; p += nn; double *X = p; p += nn; double *L = p; p += nn; double *F = p; p += nn;
  double *U = p; p += nn; double *D = p; p += nn; double *t1 = p; p += nn; double *t2 = p; p += nn; double *t3 = p; p += nn;
  double *t4 = p; p += nn; double *wS = p; p += n; double *eps = p; p += n; double *slot = p; p += 2;
  double* jw = p; p += nn + 2 * ((n + 1) / 2) + ((n + 1) / 2) + 2;     // Jacobi work (rotation arrays + int pairs)
  int* pidx = (int*)p; p += (nn + 1) / 2 + 1;
  double* eri_s = p;
  const double* eri_g = eriall + (size_t)mol * b.neri;
  for (int x = c.tid; x < nn; x += c.nt) {
    S[x] = Sall[(size_t)mol * nn + x]; H[x] = Hall[(size_t)mol * nn + x];
    const int i = x / n, j = x - i * n, lo = i < j ? i : j, hi = i < j ? j : i;
    pidx[x] = lo * n - ((lo * (lo + 1)) >> 1) + hi;
  }
  if (eri_in_smem) for (int x = c.tid; x < b.neri; x += c.nt) eri_s[x] = eri_g[x];
  c.sync();
  const double* eri = eri_in_smem ? eri_s : eri_g;

  // S^-1/2 = L diag(w^-1/2) L^T                                                  Energy.py:138-144
  c.copy(S, t1, nn);
  jacobi_smem(c, t1, L, wS, jw);
  for (int x = c.tid; x < nn; x += c.nt) t2[x] = L[x] * rsqrt(wS[x % n]);
  c.sync();
  c.mm(t2, false, L, true, X, n);
  c.copy(H, F, nn);                                                               // Energy.py:159 core guess
  double oldE = 1e8, e_elec = 0.0;
  int it = 0, converged = 0;
  double* hm = hist + (size_t)mol * max_scf * (4 * nn + n);
  for (it = 0; it < max_scf; ++it) {
    double* h = hm + (size_t)it * (4 * nn + n);
    if (want_grad) for (int x = c.tid; x < nn; x += c.nt) h[x] = F[x];            // F_{t-1}
    c.mm(X, true, F, false, t1, n);                                                // Energy.py:166
    c.mm(t1, false, X, false, t2, n);
    c.symmetrize(t2);
    jacobi_smem(c, t2, U, eps, jw);                                                // Energy.py:167-168
    c.mm(X, false, U, false, t1, n);                                               // C = X U (Energy.py:169), kept in t1
    c.mm(t1, false, t1, true, D, ne);                                              // Energy.py:174-180
    fock_smem(c, pidx, M, eri, D, H, F);                                           // Energy.py:188
    e_elec = dot3_smem(c, D, H, F, slot);                                          // Energy.py:189
    if (c.tid == 0) esteps[(size_t)mol * max_scf + it] = e_elec;
    if (want_grad) {
      for (int x = c.tid; x < nn; x += c.nt) { h[nn + x] = U[x]; h[2 * nn + x] = D[x]; h[3 * nn + x] = F[x]; }
      for (int x = c.tid; x < n; x += c.nt) h[4 * nn + x] = eps[x];
    }
    if (fabs(e_elec - oldE) < tol) { converged = 1; ++it; break; }                 // Energy.py:192-194
    oldE = e_elec;
  }
  const int k = it;        // iterations performed
  if (c.tid == 0) { Eelec[mol] = e_elec; niter[mol] = k; conv[mol] = converged; }
  for (int x = c.tid; x < nn; x += c.nt) Cout[(size_t)mol * nn + x] = t1[x];
  for (int x = c.tid; x < n; x += c.nt) epsout[(size_t)mol * n + x] = eps[x];
  if (!want_grad) return;
  c.sync();

  // ---- exact reverse sweep of the k iterations performed (same sequence as run_rhf in dq_api.cu)
  double* Dbar = S;          // S itself is no longer needed (L, wS carry it)
  double* Hbar = t4;
  double* Xbar = U;          // U_t is re-read from the history
  double* tm = terms + (size_t)mol * (max_scf + 1) * 2 * nn;
  const double* hk = hm + (size_t)(k - 1) * (4 * nn + n);
  for (int x = c.tid; x < nn; x += c.nt) {
    const double dk = hk[2 * nn + x];
    Dbar[x] = 2.0 * hk[3 * nn + x]; Hbar[x] = 2.0 * dk; Xbar[x] = 0.0;
    tm[x] = dk; tm[nn + x] = dk;                                                   // term 0: (D_k, D_k)
  }
  c.sync();
  int nt_ = 1;
  double* Ut = D;            // D_t not needed as a matrix in the sweep (taken from the history where required)
  double* Fbar = F;
  for (int t = k; t >= 1; --t) {
    const double* ht = hm + (size_t)(t - 1) * (4 * nn + n);
    for (int x = c.tid; x < nn; x += c.nt) Ut[x] = ht[nn + x];
    for (int x = c.tid; x < n; x += c.nt) eps[x] = ht[4 * nn + x];
    c.sync();
    c.mm(Ut, false, Ut, true, t3, ne);                  // P_t = U_occ U_occ^T
    c.mm(Dbar, false, X, false, t1, n);                 // t1 = Dbar X
    c.mm(t1, false, t3, false, t2, n);                  // t2 = Dbar X P
    c.add_transpose(t2, Xbar);                          // Xbar += t2 + t2^T
    c.mm(X, false, t1, false, t2, n);                   // Pbar = X Dbar X
    c.mm(Ut, true, t2, false, t1, n);                   // U^T Pbar
    c.mm(t1, false, Ut, false, t2, n);                  // Z = U^T Pbar U
    for (int x = c.tid; x < nn; x += c.nt) {            // Y_ov = Y_vo = Z_ov / (eps_o - eps_v)
      const int i = x / n, j = x - i * n;
      const bool oi = i < ne, oj = j < ne;
      double y = 0.0;
      if (oi != oj) { const int o = oi ? i : j, v = oi ? j : i; y = 0.5 * (t2[o * n + v] + t2[v * n + o]) / (eps[o] - eps[v]); }
      t3[x] = y;
    }
    c.sync();
    c.mm(Ut, false, t3, false, t1, n);                  // U Y
    c.mm(t1, false, Ut, true, t2, n);                   // Abar = U Y U^T
    c.mm(t2, false, X, false, t1, n);                   // t1 = Abar X
    for (int x = c.tid; x < nn; x += c.nt) t2[x] = ht[x];   // F_{t-1}
    c.sync();
    c.mm(t1, false, t2, false, t3, n);                  // Abar X F_{t-1}
    c.add_transpose(t3, Xbar);
    c.mm(X, false, t1, false, Fbar, n);                 // Fbar = X Abar X
    c.symmetrize(Fbar);
    c.axpy(1.0, Fbar, Hbar);
    if (t > 1) {
      const double* hp = hm + (size_t)(t - 2) * (4 * nn + n);
      double* tt = tm + (size_t)nt_ * 2 * nn;
      for (int x = c.tid; x < nn; x += c.nt) { tt[x] = Fbar[x]; tt[nn + x] = hp[2 * nn + x]; }   // (Fbar_{t-1}, D_{t-1})
      ++nt_;
      fock_smem(c, pidx, M, eri, Fbar, nullptr, Dbar);  // Dbar_{t-1} = G[Fbar]
    }
  }
  // Sbar = L [ (L^T Xbar L) o Phi ] L^T,  Phi = divided difference of lambda^-1/2
  c.mm(L, true, Xbar, false, t1, n);
  c.mm(t1, false, L, false, t2, n);
  for (int x = c.tid; x < nn; x += c.nt) {
    const double a = wS[x / n], bb = wS[x % n];
    t2[x] *= -rsqrt(a) * rsqrt(bb) / (sqrt(a) + sqrt(bb));
  }
  c.sync();
  c.mm(L, false, t2, false, t1, n);
  c.mm(t1, false, L, true, t2, n);
  for (int x = c.tid; x < nn; x += c.nt) { Sbar_out[(size_t)mol * nn + x] = t2[x]; Hbar_out[(size_t)mol * nn + x] = Hbar[x]; }
  if (c.tid == 0) nterms[mol] = nt_;
}

// ---------------------------------------------------------------------------------------------
// ERI gradient: one thread per (unique quartet, molecule); ket adjoints in per-thread shared-memory slots
// ---------------------------------------------------------------------------------------------
template <bool LA, bool LB, bool LC, bool LD>
__global__ void __launch_bounds__(128) kb_eri_grad(BatchDev b, const int2* __restrict__ qlist, int nq, const double* __restrict__ terms,
                                                   const int* __restrict__ nterms, int max_scf, double* __restrict__ padj,
                                                   double* __restrict__ pasum) {
  extern __shared__ double sacc[];       // [kkmax*5][128]
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)nq * b.nmol) return;
  const int tid = threadIdx.x;
  const int qi = (int)(id / b.nmol), mol = (int)(id - (long long)qi * b.nmol);
  const int2 q = qlist[qi];
  const int i = b.pairs[q.x].x, j = b.pairs[q.x].y, k = b.pairs[q.y].x, l = b.pairs[q.y].y;
  const int n = b.N, nn = n * n;
  double W = 0.0;
  {
    const double* tm = terms + (size_t)mol * (max_scf + 1) * 2 * nn;
    const int nt = nterms[mol];
    for (int t = 0; t < nt; ++t) {
      const double* A = tm + (size_t)t * 2 * nn;
      const double* B = A + nn;
      W += 8.0 * (A[i * n + j] * B[k * n + l] + A[k * n + l] * B[i * n + j]) -
           2.0 * (A[i * n + k] * B[j * n + l] + A[j * n + k] * B[i * n + l] + A[i * n + l] * B[j * n + k] + A[j * n + l] * B[i * n + k]);
    }
    W *= (i == j ? 0.5 : 1.0) * (k == l ? 0.5 : 1.0) * (q.x == q.y ? 0.5 : 1.0);
  }
  const int ia = b.f_type[i] - 1, ib = b.f_type[j] - 1, ic = b.f_type[k] - 1, id_ = b.f_type[l] - 1;
  const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id_);
  const size_t st = b.nmol;
  const int kkb = b.pair_off[q.x + 1] - b.pair_off[q.x], kkk = b.pair_off[q.y + 1] - b.pair_off[q.y];
  const double* brec0 = b.prec + (size_t)b.pair_off[q.x] * kPairFields * st + mol;
  const double* krec0 = b.prec + (size_t)b.pair_off[q.y] * kPairFields * st + mol;
  const long long badj0 = (long long)b.pair_off[q.x] * kAdjFields * st + mol;      // element offsets into padj
  const long long kadj0 = (long long)b.pair_off[q.y] * kAdjFields * st + mol;
  for (int x = 0; x < kkk * kAdjFields; ++x) sacc[x * 128 + tid] = 0.0;
  double pas = 0.0, pbs = 0.0, qcs = 0.0, qds = 0.0;
  if (W != 0.0) {
    for (int pb = 0; pb < kkb; ++pb) {
      const double* brec = brec0 + (size_t)pb * kPairFields * st;
      const PairPrim bra = load_pair_b(brec, st);
      PairAdj ba; ba.P[0] = ba.P[1] = ba.P[2] = ba.g = ba.K = 0.0;
      double bsl[4] = {0.0, 0.0, 0.0, 0.0}, bs[4] = {0.0, 0.0, 0.0, 0.0};
      if (LA) bs[0] = brec[ia * st];
      if (LB) bs[1] = brec[ib * st];
      if (LC) bs[2] = brec[ic * st];
      if (LD) bs[3] = brec[id_ * st];
      for (int pk = 0; pk < kkk; ++pk) {
        const double* krec = krec0 + (size_t)pk * kPairFields * st;
        const PairPrim ket = load_pair_b(krec, st);
        double ds[4] = {0.0, 0.0, 0.0, 0.0};
        if (LA) ds[0] = bs[0] - krec[ia * st];
        if (LB) ds[1] = bs[1] - krec[ib * st];
        if (LC) ds[2] = bs[2] - krec[ic * st];
        if (LD) ds[3] = bs[3] - krec[id_ * st];
        QuartetAdj qa;
        prim_quartet<LA, LB, LC, LD, true>(bra, ket, ds, dl, c_boys_b, W, &qa);
        ba.P[0] += qa.dP[0]; ba.P[1] += qa.dP[1]; ba.P[2] += qa.dP[2]; ba.g += qa.gb; ba.K += qa.Kb;
        if (LA) { bsl[0] += qa.ds[0] + qa.e[0]; pas += qa.e[0]; }
        if (LB) { bsl[1] += qa.ds[1] + qa.e[1]; pbs += qa.e[1]; }
        if (LC) { bsl[2] += qa.ds[2]; qcs += qa.e[2]; }
        if (LD) { bsl[3] += qa.ds[3]; qds += qa.e[3]; }
        double* a = sacc + (pk * kAdjFields) * 128 + tid;
        a[0] -= qa.dP[0]; a[128] -= qa.dP[1]; a[256] -= qa.dP[2]; a[384] += qa.gk; a[512] += qa.Kk;
        if (LA) a[ia * 128] -= qa.ds[0];
        if (LB) a[ib * 128] -= qa.ds[1];
        if (LC) a[ic * 128] += qa.e[2] - qa.ds[2];
        if (LD) a[id_ * 128] += qa.e[3] - qa.ds[3];
      }
      if (LA) add3(ba.P, ia, bsl[0]);
      if (LB) add3(ba.P, ib, bsl[1]);
      if (LC) add3(ba.P, ic, bsl[2]);
      if (LD) add3(ba.P, id_, bsl[3]);
      const long long o = badj0 + (long long)pb * kAdjFields * st;     // molecule-fastest: the 32 lanes of a warp hit one line
      acc_add(padj, o, ba.P[0], b.det); acc_add(padj, o + st, ba.P[1], b.det); acc_add(padj, o + 2 * st, ba.P[2], b.det);
      acc_add(padj, o + 3 * st, ba.g, b.det); acc_add(padj, o + 4 * st, ba.K, b.det);
    }
    for (int x = 0; x < kkk * kAdjFields; ++x) acc_add(padj, kadj0 + (long long)x * st, sacc[x * 128 + tid], b.det);
    if (LA) acc_add(pasum, ((long long)q.x * 2 + 0) * st + mol, pas, b.det);
    if (LB) acc_add(pasum, ((long long)q.x * 2 + 1) * st + mol, pbs, b.det);
    if (LC) acc_add(pasum, ((long long)q.y * 2 + 0) * st + mol, qcs, b.det);
    if (LD) acc_add(pasum, ((long long)q.y * 2 + 1) * st + mol, qds, b.det);
  }
}

// pair-field adjoints -> exponents, NORMALISED coefficients, centres: one thread per (function pair, molecule);
// outputs per molecule: ga[mol][nprim], gc[mol][nprim], gx[mol][N*3]
__global__ void kb_pair_bwd(BatchDev b, const double* __restrict__ padj, const double* __restrict__ pasum, double* __restrict__ ga,
                            double* __restrict__ gc, double* __restrict__ gx) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= b.M * b.nmol) return;
  const int x = id / b.nmol, mol = id - x * b.nmol;
  const int i = b.pairs[x].x, j = b.pairs[x].y;
  const double* al = b.alpha + (size_t)mol * b.nprim;
  const double* cn = b.cn + (size_t)mol * b.nprim;
  const double* xyz = b.xyz + (size_t)mol * b.N * 3;
  const size_t st = b.nmol;
  const int kj = b.f_off[j + 1] - b.f_off[j], kk = b.pair_off[x + 1] - b.pair_off[x];
  double Ab[3] = {0, 0, 0}, Bb[3] = {0, 0, 0};
  for (int pp = 0; pp < kk; ++pp) {
    const int p = b.f_off[i] + pp / kj, q = b.f_off[j] + pp % kj;
    const long long a = (long long)(b.pair_off[x] + pp) * kAdjFields * st + mol;
    PairAdj adj;
    if (b.det) { adj.P[0] = det_value(padj, a); adj.P[1] = det_value(padj, a + st); adj.P[2] = det_value(padj, a + 2 * st); adj.g = det_value(padj, a + 3 * st); adj.K = det_value(padj, a + 4 * st); }
    else { adj.P[0] = padj[a]; adj.P[1] = padj[a + st]; adj.P[2] = padj[a + 2 * st]; adj.g = padj[a + 3 * st]; adj.K = padj[a + 4 * st]; }
    double ab = 0, bb = 0, cab = 0, cbb = 0;
    pair_backward(al[p], cn[p], xyz + 3 * i, b.f_type[i], al[q], cn[q], xyz + 3 * j, b.f_type[j], adj, 0.0, 0.0, &ab, &bb, &cab, &cbb,
                  Ab, Bb);
    acc_add(ga, (long long)mol * b.nprim + p, ab, b.det); acc_add(ga, (long long)mol * b.nprim + q, bb, b.det);
    acc_add(gc, (long long)mol * b.nprim + p, cab, b.det); acc_add(gc, (long long)mol * b.nprim + q, cbb, b.det);
  }
  const long long pa = ((long long)x * 2 + 0) * st + mol, pb = ((long long)x * 2 + 1) * st + mol;
  if (b.f_type[i] > 0) Ab[b.f_type[i] - 1] -= b.det ? det_value(pasum, pa) : pasum[pa];
  if (b.f_type[j] > 0) Bb[b.f_type[j] - 1] -= b.det ? det_value(pasum, pb) : pasum[pb];
  const long long g = (long long)mol * b.N * 3;
#pragma unroll
  for (int cdir = 0; cdir < 3; ++cdir) { acc_add(gx, g + 3 * i + cdir, Ab[cdir], b.det); acc_add(gx, g + 3 * j + cdir, Bb[cdir], b.det); }
}

// Sbar : dS + Hbar : d(T+V), one thread per (primitive-pair record, molecule), 10-direction duals
__global__ void kb_one_electron_grad(BatchDev b, const double* __restrict__ Sbar, const double* __restrict__ Hbar,
                                     double* __restrict__ ga, double* __restrict__ gc, double* __restrict__ gx) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)b.nrec * b.nmol) return;
  const int r = (int)(id / b.nmol), mol = (int)(id - (long long)r * b.nmol);
  const int x = b.rec_pair[r], pp = r - b.pair_off[x];
  const int i = b.pairs[x].x, j = b.pairs[x].y, n = b.N;
  const int kj = b.f_off[j + 1] - b.f_off[j];
  const int p = b.f_off[i] + pp / kj, q = b.f_off[j] + pp % kj;
  const double* al = b.alpha + (size_t)mol * b.nprim;
  const double* cn = b.cn + (size_t)mol * b.nprim;
  const double* xyz = b.xyz + (size_t)mol * n * 3;
  typedef Dual<10> Du;
  Du A[3], B[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { A[c] = seed<10>(xyz[3 * i + c], 4 + c); B[c] = seed<10>(xyz[3 * j + c], 7 + c); }
  Du Sv, Tv, Vv;
  one_electron_primitive<Du>(seed<10>(al[p], 0), seed<10>(cn[p], 2), A, b.f_type[i], seed<10>(al[q], 1), seed<10>(cn[q], 3), B,
                             b.f_type[j], b.natoms, b.Z, b.atoms + (size_t)mol * b.natoms * 3, c_boys_b, &Sv, &Tv, &Vv);
  const double* Sb = Sbar + (size_t)mol * n * n;
  const double* Hb = Hbar + (size_t)mol * n * n;
  const double ws = (i == j) ? Sb[i * n + i] : Sb[i * n + j] + Sb[j * n + i];
  const double wh = (i == j) ? Hb[i * n + i] : Hb[i * n + j] + Hb[j * n + i];
  double g[10];
#pragma unroll
  for (int c = 0; c < 10; ++c) g[c] = ws * Sv.d[c] + wh * (Tv.d[c] + Vv.d[c]);
  const long long gam = (long long)mol * b.nprim, gxm = (long long)mol * n * 3;
  acc_add(ga, gam + p, g[0], b.det); acc_add(ga, gam + q, g[1], b.det); acc_add(gc, gam + p, g[2], b.det); acc_add(gc, gam + q, g[3], b.det);
#pragma unroll
  for (int c = 0; c < 3; ++c) { acc_add(gx, gxm + 3 * i + c, g[4 + c], b.det); acc_add(gx, gxm + 3 * j + c, g[7 + c], b.det); }
}

// normalisation adjoint (gc: normalised -> raw, ga +=) and d/d ln(alpha): one thread per (function, molecule)
__global__ void kb_finish(BatchDev b, int log_alpha, double* __restrict__ ga, const double* __restrict__ gc, double* __restrict__ graw) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= b.N * b.nmol) return;
  const int f = id / b.nmol, mol = id - f * b.nmol;
  const int L = b.f_type[f] > 0 ? 1 : 0;
  const double pw = 0.5 * L + 0.75, e = L + 1.5;
  const double fac = pow(2.0 / kPi, 0.75) * (L ? 2.0 : 1.0);
  const double div = pow(kPi, 1.5) / (L ? 2.0 : 1.0);
  const int p0 = b.f_off[f], p1 = b.f_off[f + 1];
  const double* alpha = b.alpha + (size_t)mol * b.nprim;
  const double* craw = b.craw + (size_t)mol * b.nprim;
  const double* cbar = gc + (size_t)mol * b.nprim;
  double* abar = ga + (size_t)mol * b.nprim;
  double* crawbar = graw + (size_t)mol * b.nprim;
  double tm = 0.0, dot = 0.0;
  for (int p = p0; p < p1; ++p) {
    const double up = fac * pow(alpha[p], pw) * craw[p];
    dot += cbar[p] * up;
    for (int q = p0; q < p1; ++q) tm += fac * pow(alpha[q], pw) * craw[q] * up / pow(alpha[p] + alpha[q], e);
  }
  const double nrm = sqrt(tm * div);
  const double nrmb = -dot / (nrm * nrm);
  const double tmb = nrmb * div / (2.0 * nrm);
  for (int p = p0; p < p1; ++p) {
    const double np_ = fac * pow(alpha[p], pw);
    const double up = np_ * craw[p];
    double sg = 0.0, sdg = 0.0;
    for (int q = p0; q < p1; ++q) {
      const double uq = fac * pow(alpha[q], pw) * craw[q];
      const double gpq = pow(alpha[p] + alpha[q], -e);
      sg += uq * gpq;
      sdg += uq * gpq * (-e) / (alpha[p] + alpha[q]);
    }
    const double ub = cbar[p] / nrm + tmb * 2.0 * sg;
    crawbar[p] = ub * np_;
    double a = abar[p] + ub * up * pw / alpha[p] + tmb * 2.0 * up * sdg;
    if (log_alpha) a *= alpha[p];
    abar[p] = a;
  }
}

// =============================================================================================
// launchers
// =============================================================================================
static inline int cdivb(long long a, int b) { return (int)((a + b - 1) / b); }

void launch_batch_prepare(cudaStream_t st, const BatchDev& b, const double* alpha_in, int log_alpha) {
  kb_prepare<<<cdivb((long long)b.N * b.nmol, 128), 128, 0, st>>>(b, alpha_in, log_alpha); DQB_LAUNCHED();
  kb_pairs<<<cdivb((long long)b.nrec * b.nmol, 128), 128, 0, st>>>(b); DQB_LAUNCHED();
}
void launch_batch_one_electron(cudaStream_t st, const BatchDev& b, double* S, double* H) {
  kb_one_electron<<<cdivb((long long)b.M * b.nmol, 64), 64, 0, st>>>(b, S, H); DQB_LAUNCHED();
}

#define DQB_CLASS_SWITCH(cls, CALL)                                                              \
  switch (cls) {                                                                                \
    case 0: CALL(false, false, false, false); break; case 1: CALL(false, false, false, true); break;   \
    case 2: CALL(false, false, true, false); break;  case 3: CALL(false, false, true, true); break;    \
    case 4: CALL(false, true, false, false); break;  case 5: CALL(false, true, false, true); break;    \
    case 6: CALL(false, true, true, false); break;   case 7: CALL(false, true, true, true); break;     \
    case 8: CALL(true, false, false, false); break;  case 9: CALL(true, false, false, true); break;    \
    case 10: CALL(true, false, true, false); break;  case 11: CALL(true, false, true, true); break;    \
    case 12: CALL(true, true, false, false); break;  case 13: CALL(true, true, false, true); break;    \
    case 14: CALL(true, true, true, false); break;   case 15: CALL(true, true, true, true); break;     \
  }

void launch_batch_eri(cudaStream_t st, int cls, const BatchDev& b, const int2* qlist, int nq, double* eri) {
  if (nq <= 0) return;
  const int blocks = cdivb((long long)nq * b.nmol, 128);
#define CALL(a, bb, c, d) kb_eri<a, bb, c, d><<<blocks, 128, 0, st>>>(b, qlist, nq, eri)
  DQB_CLASS_SWITCH(cls, CALL)
#undef CALL
  DQB_LAUNCHED();
}

template <bool A, bool B, bool C, bool D>
static void batch_grad_launch(cudaStream_t st, int blocks, size_t smem, const BatchDev& b, const int2* qlist, int nq, const double* terms,
                              const int* nterms, int max_scf, double* padj, double* pasum) {
  ensure_dynamic_smem(kb_eri_grad<A, B, C, D>, smem);
  kb_eri_grad<A, B, C, D><<<blocks, 128, smem, st>>>(b, qlist, nq, terms, nterms, max_scf, padj, pasum);
}
void launch_batch_eri_grad(cudaStream_t st, int cls, const BatchDev& b, const int2* qlist, int nq, const double* terms, const int* nterms,
                           int max_scf, double* padj, double* pasum) {
  if (nq <= 0) return;
  const int blocks = cdivb((long long)nq * b.nmol, 128);
  const size_t smem = (size_t)b.kkmax * kAdjFields * 128 * sizeof(double);
#define CALL(a, bb, c, d) batch_grad_launch<a, bb, c, d>(st, blocks, smem, b, qlist, nq, terms, nterms, max_scf, padj, pasum)
  DQB_CLASS_SWITCH(cls, CALL)
#undef CALL
  DQB_LAUNCHED();
}

size_t batch_scf_smem_bytes(int n, int neri, bool eri_in_smem) {
  const int nn = n * n, m = (n + 1) / 2;
  size_t d = (size_t)11 * nn + 2 * n + 2 + (nn + 3 * m + 2) + ((nn + 1) / 2 + 1);
  if (eri_in_smem) d += neri;
  return d * sizeof(double);
}
int launch_batch_scf(cudaStream_t st, const BatchDev& b, const double* S, const double* H, const double* eri, int max_scf, double tol,
                     int want_grad, double* hist, double* Eelec, int* niter, int* conv, double* esteps, double* C, double* eps,
                     double* terms, int* nterms, double* Hbar, double* Sbar) {
  bool in_smem = batch_scf_smem_bytes(b.N, b.neri, true) <= 100 * 1024;
  const size_t smem = batch_scf_smem_bytes(b.N, b.neri, in_smem);
  if (smem > 220 * 1024) return -1;
  ensure_dynamic_smem(kb_scf, smem);
  kb_scf<<<b.nmol, 128, smem, st>>>(b, S, H, eri, in_smem ? 1 : 0, max_scf, tol, want_grad, hist, Eelec, niter, conv, esteps, C, eps, terms,
                                    nterms, Hbar, Sbar);
  DQB_LAUNCHED();
  return 0;
}
// ga, gc, gx: accumulators (limb pairs when b.det)
void launch_batch_grad_accumulate(cudaStream_t st, const BatchDev& b, const double* padj, const double* pasum, const double* Sbar,
                                  const double* Hbar, double* ga, double* gc, double* gx) {
  kb_pair_bwd<<<cdivb((long long)b.M * b.nmol, 64), 64, 0, st>>>(b, padj, pasum, ga, gc, gx); DQB_LAUNCHED();
  kb_one_electron_grad<<<cdivb((long long)b.nrec * b.nmol, 64), 64, 0, st>>>(b, Sbar, Hbar, ga, gc, gx); DQB_LAUNCHED();
}
// ga, gc: doubles
void launch_batch_grad_finish(cudaStream_t st, const BatchDev& b, int log_alpha, double* ga, const double* gc, double* graw) {
  kb_finish<<<cdivb((long long)b.N * b.nmol, 128), 128, 0, st>>>(b, log_alpha, ga, gc, graw); DQB_LAUNCHED();
}

}  // namespace dq
