// diffiqult_b200 — CUDA kernels of the RHF hot path for sm_100a (B200).  FP64 throughout.
//
//   k_normalize / k_normalize_bwd   contraction-coefficient normalisation + adjoint   (Integrals.py:463-501)
//   k_pairs                         primitive-pair records                             (Integrals.py:237-249)
//   k_one_electron(+_grad)          S, T, V and their parameter gradient                (Integrals.py:10-188)
//   k_eri_tiles<class>              contracted ERIs, one 256-quartet tile per CTA       (Integrals.py:190-401)
//   k_scatter_packed                tile store -> the reference's packed Eri_vec order  (Tools.py:17-51)
//   k_jk_cols                       J/K contraction of the stored tiles with D          (Energy.py:29-38)
//   k_eri_grad_chunks<class>        sum Gamma_ijkl d(ij|kl)/d(pair fields), reverse mode, chunks of tiles sharing the ket
//   k_pair_bwd                      pair-field adjoints -> exponents, coefficients, centres
//   k_gemm, k_eigh_jacobi, small elementwise kernels for the SCF and its reverse sweep (Energy.py:138-195, Eigen.py)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <vector>

#include "dq_layout.h"
#include "dq_launch.h"
#include "dq_math.cuh"

namespace dq {

__constant__ BoysConst c_boys;
static thread_local long long g_launches = 0;   // kernels launched by this host thread (reported by dq_stats)

long long launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }

// one "Boys loop ran out of terms" flag per device, shared by every host thread driving that device
int* boys_fail_flag() {
  static std::mutex mu;
  static std::map<int, int*> flags;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  int*& f = flags[dev];
  if (!f) {
    if (cudaMalloc(&f, sizeof(int)) != cudaSuccess) { f = nullptr; return nullptr; }
    cudaMemset(f, 0, sizeof(int));
  }
  return f;
}

int upload_boys_constants() {
  BoysConst h;
  boys_const_init(&h);
  h.fail = boys_fail_flag();
  if (!h.fail) return 1;
  return (int)cudaMemcpyToSymbol(c_boys, &h, sizeof(h));
}

#define DQ_LAUNCHED() (++g_launches)

// =============================================================================================
// normalisation
// =============================================================================================
__global__ void k_normalize(int N, const int* __restrict__ f_off, const int* __restrict__ f_type,
                            const double* __restrict__ alpha, const double* __restrict__ craw, double* __restrict__ cn) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  const int L = f_type[f] > 0 ? 1 : 0;
  const double pw = 0.5 * L + 0.75, e = L + 1.5;
  const double fac = pow(2.0 / kPi, 0.75) * (L ? 2.0 : 1.0);         // (2/pi)^(3/4) * div * 2^(L/2)
  const double div = pow(kPi, 1.5) / (L ? 2.0 : 1.0);
  const int p0 = f_off[f], p1 = f_off[f + 1];
  double tm = 0.0;
  for (int p = p0; p < p1; ++p) {
    const double up = fac * pow(alpha[p], pw) * craw[p];
    for (int q = p0; q < p1; ++q) {
      const double uq = fac * pow(alpha[q], pw) * craw[q];
      tm += uq * up / pow(alpha[p] + alpha[q], e);
    }
  }
  const double nrm = sqrt(tm * div);
  for (int p = p0; p < p1; ++p) cn[p] = fac * pow(alpha[p], pw) * craw[p] / nrm;
}

// adjoint: cbar (wrt normalised coefficients) -> crawbar (=) and abar (+=)
__global__ void k_normalize_bwd(int N, const int* __restrict__ f_off, const int* __restrict__ f_type,
                                const double* __restrict__ alpha, const double* __restrict__ craw,
                                const double* __restrict__ cbar, double* __restrict__ abar, double* __restrict__ crawbar) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  const int L = f_type[f] > 0 ? 1 : 0;
  const double pw = 0.5 * L + 0.75, e = L + 1.5;
  const double fac = pow(2.0 / kPi, 0.75) * (L ? 2.0 : 1.0);
  const double div = pow(kPi, 1.5) / (L ? 2.0 : 1.0);
  const int p0 = f_off[f], p1 = f_off[f + 1];
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
    abar[p] += ub * up * pw / alpha[p] + tmb * 2.0 * up * sdg;
  }
}

// =============================================================================================
// primitive-pair records
// =============================================================================================
__global__ void __launch_bounds__(128) k_pairs(SystemDev s, double* __restrict__ pd, double* __restrict__ bp_kmax) {
  const BlockPair b = s.bp[blockIdx.x];
  double kmax = 0.0;
  const int lane = threadIdx.x & 15, il = lane >> 2, jl = lane & 3;
  const bool valid = il < b.nI && jl < b.nJ;
  const int i = b.sI + il, j = b.sJ + jl;
  for (int pp = threadIdx.x >> 4; pp < b.KK; pp += blockDim.x >> 4) {
    PairPrim r;
    if (valid) {
      const int p = s.f_off[i] + pp / b.KJ, q = s.f_off[j] + pp % b.KJ;
      r = make_pair(s.p_alpha[p], s.p_coef[p], s.f_xyz + 3 * i, s.f_type[i], s.p_alpha[q], s.p_coef[q], s.f_xyz + 3 * j,
                    s.f_type[j]);
    } else {
      r.P[0] = r.P[1] = r.P[2] = 0.0; r.g = 1.0; r.ginv = 1.0; r.K = 0.0; r.pa = r.pb = 0.0;
    }
    double* o = pd + ((b.off + (long long)pp * kPairFields) * 16 + lane);
    o[0] = r.P[0]; o[16] = r.P[1]; o[32] = r.P[2]; o[48] = r.g; o[64] = r.ginv; o[80] = r.K; o[96] = r.pa; o[112] = r.pb;
    kmax = fmax(kmax, fabs(r.K));
    if (valid && s.underflowed && pair_underflowed(r.K)) atomicAdd(s.underflowed, 1ULL);
  }
  // non-negative doubles order like their bit patterns (bp_kmax is zeroed by the launcher)
  if (kmax > 0.0) atomicMax((unsigned long long*)(bp_kmax + blockIdx.x), (unsigned long long)__double_as_longlong(kmax));
}

__device__ __forceinline__ PairPrim load_pair(const double* __restrict__ rec, int lane) {
  PairPrim r;
  r.P[0] = rec[lane]; r.P[1] = rec[16 + lane]; r.P[2] = rec[32 + lane];
  r.g = rec[48 + lane]; r.ginv = rec[64 + lane]; r.K = rec[80 + lane]; r.pa = rec[96 + lane]; r.pb = rec[112 + lane];
  return r;
}

// =============================================================================================
// one-electron integrals
// =============================================================================================
__device__ __forceinline__ void pair_from_linear(int x, int n, int* i, int* j) {  // inverse of i*n - i(i-1)/2 + (j-i)
  int ii = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * x)) * 0.5);
  while (ii > 0 && (long long)ii * n - (long long)ii * (ii - 1) / 2 > x) --ii;
  while ((long long)(ii + 1) * n - (long long)(ii + 1) * ii / 2 <= x) ++ii;
  *i = ii; *j = ii + (x - (ii * n - ii * (ii - 1) / 2));
}

__global__ void k_one_electron(SystemDev s, double* __restrict__ S, double* __restrict__ T, double* __restrict__ V) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = s.N;
  if (x >= n * (n + 1) / 2) return;
  int i, j; pair_from_linear(x, n, &i, &j);
  double sv = 0.0, tv = 0.0, vv = 0.0;
  for (int p = s.f_off[i]; p < s.f_off[i + 1]; ++p)
    for (int q = s.f_off[j]; q < s.f_off[j + 1]; ++q) {
      double a, b, c;
      one_electron_primitive<double>(s.p_alpha[p], s.p_coef[p], s.f_xyz + 3 * i, s.f_type[i], s.p_alpha[q], s.p_coef[q],
                                     s.f_xyz + 3 * j, s.f_type[j], s.natoms, s.Z, s.C, c_boys, &a, &b, &c);
      sv += a; tv += b; vv += c;
    }
  S[i * n + j] = S[j * n + i] = sv; T[i * n + j] = T[j * n + i] = tv; V[i * n + j] = V[j * n + i] = vv;
}

// one thread per (function pair, primitive pair): contributions of Sbar:dS + Hbar:(dT+dV)
__global__ void k_one_electron_grad(SystemDev s, int kmax, const double* __restrict__ Sbar, const double* __restrict__ Hbar,
                                    double* __restrict__ g_alpha, double* __restrict__ g_coef, double* __restrict__ g_xyz) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = s.N;
  const int npair = n * (n + 1) / 2;
  const int x = (int)(id / (kmax * kmax));
  if (x >= npair) return;
  const int pp = (int)(id % (kmax * kmax));
  int i, j; pair_from_linear(x, n, &i, &j);
  const int ki = s.f_off[i + 1] - s.f_off[i], kj = s.f_off[j + 1] - s.f_off[j];
  const int pi = pp / kmax, pj = pp % kmax;
  if (pi >= ki || pj >= kj) return;
  const int p = s.f_off[i] + pi, q = s.f_off[j] + pj;
  typedef Dual<10> D;
  D A[3], B[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { A[c] = seed<10>(s.f_xyz[3 * i + c], 4 + c); B[c] = seed<10>(s.f_xyz[3 * j + c], 7 + c); }
  D Sv, Tv, Vv;
  one_electron_primitive<D>(seed<10>(s.p_alpha[p], 0), seed<10>(s.p_coef[p], 2), A, s.f_type[i], seed<10>(s.p_alpha[q], 1),
                            seed<10>(s.p_coef[q], 3), B, s.f_type[j], s.natoms, s.Z, s.C, c_boys, &Sv, &Tv, &Vv);
  const double ws = (i == j) ? Sbar[i * n + i] : Sbar[i * n + j] + Sbar[j * n + i];
  const double wh = (i == j) ? Hbar[i * n + i] : Hbar[i * n + j] + Hbar[j * n + i];
  double g[10];
#pragma unroll
  for (int c = 0; c < 10; ++c) g[c] = ws * Sv.d[c] + wh * (Tv.d[c] + Vv.d[c]);
  // g_alpha, g_coef, g_xyz are three windows of ONE accumulator buffer (limb pairs in fixed-point mode): index from g_alpha
  const long long oc = g_coef - g_alpha, ox = g_xyz - g_alpha;
  acc_add(g_alpha, p, g[0], s.det); acc_add(g_alpha, q, g[1], s.det);
  acc_add(g_alpha, oc + p, g[2], s.det); acc_add(g_alpha, oc + q, g[3], s.det);
#pragma unroll
  for (int c = 0; c < 3; ++c) { acc_add(g_alpha, ox + 3 * i + c, g[4 + c], s.det); acc_add(g_alpha, ox + 3 * j + c, g[7 + c], s.det); }
}

// Nuclear-coordinate gradient (the reference's branch Energy.py:125-130: xyz_atom carries the tangents, the Gaussians stay
// where they are, so only V and E_nuc depend on the nuclei): g_atoms[k][c] += sum_ij Hbar_ij dV_ij/dC_kc.  One thread per
// (function pair, primitive pair) walks the nuclei with 3-direction duals seeded on C_k; the lanes of a warp walk the
// same nucleus, so each (k, c) is reduced over the warp before ONE accumulation.
__global__ void __launch_bounds__(128) k_nuclear_grad(SystemDev s, int kmax, const double* __restrict__ Hbar, double* __restrict__ g_atoms) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = s.N;
  const int npair = n * (n + 1) / 2;
  const int x = (int)(id / (kmax * kmax));
  bool on = x < npair;
  const int pp = (int)(id % (kmax * kmax));
  int i = 0, j = 0;
  if (on) pair_from_linear(x, n, &i, &j);
  const int ki = s.f_off[i + 1] - s.f_off[i], kj = s.f_off[j + 1] - s.f_off[j];
  const int pi = pp / kmax, pj = pp % kmax;
  on = on && pi < ki && pj < kj;
  const int p = s.f_off[i] + (on ? pi : 0), q = s.f_off[j] + (on ? pj : 0);
  typedef Dual<3> D3;
  D3 A[3], B[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { A[c] = D3(s.f_xyz[3 * i + c]); B[c] = D3(s.f_xyz[3 * j + c]); }
  const double wh = !on ? 0.0 : ((i == j) ? Hbar[i * n + i] : Hbar[i * n + j] + Hbar[j * n + i]);
  for (int k = 0; k < s.natoms; ++k) {
    D3 Ck[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) Ck[c] = seed<3>(s.C[3 * k + c], c);
    D3 Sv, Tv, Vv;
    one_electron_primitive<D3, D3>(D3(s.p_alpha[p]), D3(s.p_coef[p]), A, s.f_type[i], D3(s.p_alpha[q]), D3(s.p_coef[q]), B, s.f_type[j],
                                   1, s.Z + k, Ck, c_boys, &Sv, &Tv, &Vv);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double g = wh * Vv.d[c];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if ((threadIdx.x & 31) == 0 && g != 0.0) acc_add(g_atoms, 3 * k + c, g, s.det);
    }
  }
}

// =============================================================================================
// ERI tiles
// =============================================================================================
struct TileCtx {
  BlockPair b1, b2;
  int il, jl, kl, ll, bl, kt;
  bool valid;
};

__device__ __forceinline__ TileCtx tile_ctx(const Tile tl, const SystemDev& s) {
  TileCtx c;
  c.b1 = s.bp[tl.bp1]; c.b2 = s.bp[tl.bp2];
  const int tid = threadIdx.x;
  c.il = tid >> 6; c.jl = (tid >> 4) & 3; c.kl = (tid >> 2) & 3; c.ll = tid & 3;
  c.bl = tid >> 4; c.kt = tid & 15;
  c.valid = c.il < c.b1.nI && c.jl < c.b1.nJ && c.kl < c.b2.nI && c.ll < c.b2.nJ;
  return c;
}

__device__ __forceinline__ void stage_pairs(const TileCtx& c, const double* __restrict__ pd, double* sbra, double* sket) {
  const double* gb = pd + c.b1.off * 16;
  const double* gk = pd + c.b2.off * 16;
  const int nb = c.b1.KK * kPairFields * 16, nk = c.b2.KK * kPairFields * 16;
  for (int x = threadIdx.x; x < nb; x += blockDim.x) sbra[x] = gb[x];
  for (int x = threadIdx.x; x < nk; x += blockDim.x) sket[x] = gk[x];
}

// opt-in screening (dq_options.screen_threshold): 2 pi^(5/2) max|K_ab| max|K_cd| bounds the prefactor of every primitive
// quartet of the tile; CTA-uniform
__device__ __forceinline__ bool tile_screened(const SystemDev& s, const Tile tl) {
  return s.screen > 0.0 && kTwoPi52 * s.bp_kmax[tl.bp1] * s.bp_kmax[tl.bp2] < s.screen;
}

// ZS: leave primitive pairs with an exactly zero prefactor out of the loops.  A template parameter, not a run-time test:
// the variant without it is the kernel as it was before the skip existed (a dead branch in these loops costs ~10 %).
template <bool LA, bool LB, bool LC, bool LD, bool ZS>
__global__ void __launch_bounds__(256) k_eri_tiles(const Tile* __restrict__ tiles, SystemDev s, const double* __restrict__ pd,
                                                   double* __restrict__ store) {
  extern __shared__ double sm[];
  if (tile_screened(s, tiles[blockIdx.x])) {
    store[(long long)blockIdx.x * kTile + threadIdx.x] = 0.0;
    if (threadIdx.x == 0) atomicAdd(s.skipped, 1ULL);
    return;
  }
  const TileCtx c = tile_ctx(tiles[blockIdx.x], s);
  double* sbra = sm;
  double* sket = sm + c.b1.KK * kPairFields * 16;
  stage_pairs(c, pd, sbra, sket);
  __syncthreads();
  double v = 0.0;
  if (c.valid) {
    const int ia = c.b1.axI, ib = c.b1.axJ, ic = c.b2.axI, id = c.b2.axJ;
    const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
    for (int pb = 0; pb < c.b1.KK; ++pb) {
      const double* brec = sbra + pb * kPairFields * 16;
      const PairPrim bra = load_pair(brec, c.bl);
      if (ZS && bra.K == 0.0) continue;      // exactly zero prefactor: the quartets below are exact zeros
      // the components of P along the axes of the p functions come straight from the staged records at a CTA-uniform
      // offset: no per-primitive select
      double bs[4] = {0.0, 0.0, 0.0, 0.0};
      if (LA) bs[0] = brec[ia * 16 + c.bl];
      if (LB) bs[1] = brec[ib * 16 + c.bl];
      if (LC) bs[2] = brec[ic * 16 + c.bl];
      if (LD) bs[3] = brec[id * 16 + c.bl];
      for (int pk = 0; pk < c.b2.KK; ++pk) {
        const double* krec = sket + pk * kPairFields * 16;
        const PairPrim ket = load_pair(krec, c.kt);
        if (ZS && ket.K == 0.0) continue;
        double ds[4] = {0.0, 0.0, 0.0, 0.0};
        if (LA) ds[0] = bs[0] - krec[ia * 16 + c.kt];
        if (LB) ds[1] = bs[1] - krec[ib * 16 + c.kt];
        if (LC) ds[2] = bs[2] - krec[ic * 16 + c.kt];
        if (LD) ds[3] = bs[3] - krec[id * 16 + c.kt];
        v += prim_quartet<LA, LB, LC, LD, false>(bra, ket, ds, dl, c_boys, 0.0, nullptr);
      }
    }
  }
  // streaming store (evict-first): the 2.6 GB tile stream must not push the 15 MB of pair records, which every CTA
  // re-reads, out of L2 - the pass is bound by the latency of those reads
  __stcs(store + (long long)blockIdx.x * kTile + threadIdx.x, v);
}

// ---------------------------------------------------------------------------------------------
// ERI tiles with the LOOP-REGIME quartets deferred and compacted per tile (the default for contractions of <= 96 primitive
// pairs per tile side product, i.e. STO-3G-like bases).  Timing experiment (DQ_EXPERIMENT_NO_LOOPS build): on the dense
// (H2O)_32 cluster the 6 % of the primitive quartets whose Boys argument lies below the mid range cost 42 % of the pass
// (229 ms against 134 ms when they are sent through the loop-free branch): they drag whole warps through the series /
// continued-fraction loops with 8 of 32 lanes on.  Compaction inside a warp or a row does not pay (k_eri_tiles_b);
// what does is taking them out of the warp's instruction stream altogether:
//   phase A  every thread walks its 81 primitive quartets as k_eri_tiles does, evaluates the loop-free ones (closed form /
//            mid range: no loop code is compiled into this phase) and only MARKS the others - one bit per step in a
//            3-word mask - counting them per bin of the Boys argument (32 bins below the mid range);
//   list     after a scan of the bin counts the marked quartets are listed bin by bin (shared-memory cursor per bin), so
//            that the lanes of a warp of phase B run series / fraction loops of like length;
//   phase B  the 256 threads take list entries round robin - whoever owns them - evaluate the quartet with the loop
//            code, and add the value to the owner's accumulator in shared memory: fixed point (two 64-bit limbs), so the
//            sum does not depend on who finishes first.  The tile value is phase A's register sum plus that accumulator:
//            bit-reproducible, and equal to k_eri_tiles' value up to the order of the additions.
// ---------------------------------------------------------------------------------------------
constexpr int kDefMaxSteps = 96;      // primitive quartets per contracted quartet the 3-word mask can mark

__device__ __forceinline__ void det_add_shared(unsigned long long* p, double x) {      // dq::det_add on a shared-memory limb pair
  const double h = rint(x * 16777216.0);
  const double r = fma(-h, 1.0 / 16777216.0, x);
  const long long hi = (long long)h, lo = (long long)rint(r * 147573952589676412928.0);
  if (hi != 0) atomicAdd(p, (unsigned long long)hi);
  if (lo != 0) atomicAdd(p + 1, (unsigned long long)lo);
}

template <bool LA, bool LB, bool LC, bool LD, bool ZS>
__global__ void __launch_bounds__(256) k_eri_tiles_def(const Tile* __restrict__ tiles, SystemDev s, const double* __restrict__ pd,
                                                       double* __restrict__ store) {
  constexpr int L = (int)LA + (int)LB + (int)LC + (int)LD;
  extern __shared__ double sm[];
  __shared__ int s_cnt[32], s_cur[32], s_total;
  if (tile_screened(s, tiles[blockIdx.x])) {
    store[(long long)blockIdx.x * kTile + threadIdx.x] = 0.0;
    if (threadIdx.x == 0) atomicAdd(s.skipped, 1ULL);
    return;
  }
  const TileCtx c = tile_ctx(tiles[blockIdx.x], s);
  const int KKb = c.b1.KK, KKk = c.b2.KK;            // KKb * KKk <= kDefMaxSteps (host)
  const int tid = threadIdx.x;
  double* sbra = sm;
  double* sket = sbra + KKb * kPairFields * 16;
  unsigned long long* accfx = (unsigned long long*)(sket + KKk * kPairFields * 16);      // [256][2]
  unsigned short* items = (unsigned short*)(accfx + 512);                                 // [256 * steps] (owner << 7 | step)
  stage_pairs(c, pd, sbra, sket);
  accfx[2 * tid] = 0; accfx[2 * tid + 1] = 0;
  if (tid < 32) s_cnt[tid] = 0;
  __syncthreads();
  const int ia = c.b1.axI, ib = c.b1.axJ, ic = c.b2.axI, id = c.b2.axJ;
  const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
  const double tmid = boys_midrange_from<L>();
  const double binscale = 31.99 / tmid;      // 32 bins of the Boys argument below the mid range
  auto bin_of = [&](double t) { return t > 0.0 ? min(31, (int)(t * binscale)) : 0; };
  // ---- phase A: loop-free quartets evaluated, loop-regime quartets marked and counted per argument bin
  double v = 0.0;
  unsigned m0 = 0u, m1 = 0u, m2 = 0u;      // (three scalars: a dynamically indexed array would live in local memory)
  // the argument bins of the first 36 marked quartets, 5 bits each, in marking order: listing them needs no second
  // evaluation of t (ncu: recomputing it for every listed quartet made the listing loop 12 % of the kernel's instructions
  // at 6 lanes per instruction)
  unsigned long long p0 = 0ull, p1 = 0ull, p2 = 0ull;
  int nmark = 0;
  for (int pb = 0; pb < KKb; ++pb) {
    const double* brec = sbra + pb * kPairFields * 16;
    const PairPrim bra = load_pair(brec, c.bl);
    if (!c.valid || (ZS && bra.K == 0.0)) continue;
    double bs[4] = {0.0, 0.0, 0.0, 0.0};
    if (LA) bs[0] = brec[ia * 16 + c.bl];
    if (LB) bs[1] = brec[ib * 16 + c.bl];
    if (LC) bs[2] = brec[ic * 16 + c.bl];
    if (LD) bs[3] = brec[id * 16 + c.bl];
    for (int pk = 0; pk < KKk; ++pk) {
      const int st = pb * KKk + pk;
      const double* krec = sket + pk * kPairFields * 16;
      const PairPrim ket = load_pair(krec, c.kt);
      if (ZS && ket.K == 0.0) continue;
      const double t = prim_quartet_t(bra, ket);
      if (!(t >= tmid)) {                       // (a NaN argument goes to the loops, which flag it)
        const unsigned bit = 1u << (st & 31);
        if (st < 32) m0 |= bit; else if (st < 64) m1 |= bit; else m2 |= bit;
        const int bin = bin_of(t);
        atomicAdd(&s_cnt[bin], 1);
        if (nmark < 12) p0 |= (unsigned long long)bin << (5 * nmark);
        else if (nmark < 24) p1 |= (unsigned long long)bin << (5 * (nmark - 12));
        else if (nmark < 36) p2 |= (unsigned long long)bin << (5 * (nmark - 24));
        ++nmark;
        continue;
      }
      double ds[4] = {0.0, 0.0, 0.0, 0.0};
      if (LA) ds[0] = bs[0] - krec[ia * 16 + c.kt];
      if (LB) ds[1] = bs[1] - krec[ib * 16 + c.kt];
      if (LC) ds[2] = bs[2] - krec[ic * 16 + c.kt];
      if (LD) ds[3] = bs[3] - krec[id * 16 + c.kt];
      v += prim_quartet<LA, LB, LC, LD, false, kBoysNoLoops>(bra, ket, ds, dl, c_boys, 0.0, nullptr);
    }
  }
  __syncthreads();
  if (tid < 32) {                               // exclusive scan of the 32 bin counts
    const int cnt = s_cnt[tid];
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
    s_cur[tid] = incl - cnt;
    if (tid == 31) s_total = incl;
  }
  __syncthreads();
  const int total = s_total;
  if (total > 0) {                              // CTA-uniform
    // the marked quartets listed by argument bin: the lanes of a warp of phase B run loops of like length
    int j = 0;
#pragma unroll
    for (int wd = 0; wd < 3; ++wd) {
      for (unsigned m = wd == 0 ? m0 : (wd == 1 ? m1 : m2); m; m &= m - 1, ++j) {
        const int st = wd * 32 + __ffs(m) - 1;
        int bin;
        if (j < 36) {
          const unsigned long long pw = j < 12 ? p0 : (j < 24 ? p1 : p2);
          bin = (int)((pw >> (5 * (j % 12))) & 31ull);
        } else {                                 // more than 36 marked quartets in one thread: recompute
          const int pb = st / KKk, pk = st - pb * KKk;
          const PairPrim bra = load_pair(sbra + pb * kPairFields * 16, c.bl);
          const PairPrim ket = load_pair(sket + pk * kPairFields * 16, c.kt);
          bin = bin_of(prim_quartet_t(bra, ket));
        }
        items[atomicAdd(&s_cur[bin], 1)] = (unsigned short)((tid << 7) | st);
      }
    }
    __syncthreads();
    // ---- phase B: list entries round robin, whoever owns them
    for (int x = tid; x < total; x += 256) {
      const int item = items[x];
      const int owner = item >> 7, st = item & 127;
      const int pb = st / KKk, pk = st - pb * KKk;
      const int bl = owner >> 4, kt = owner & 15;
      const double* brec = sbra + pb * kPairFields * 16;
      const double* krec = sket + pk * kPairFields * 16;
      const PairPrim bra = load_pair(brec, bl);
      const PairPrim ket = load_pair(krec, kt);
      double ds[4] = {0.0, 0.0, 0.0, 0.0};
      if (LA) ds[0] = brec[ia * 16 + bl] - krec[ia * 16 + kt];
      if (LB) ds[1] = brec[ib * 16 + bl] - krec[ib * 16 + kt];
      if (LC) ds[2] = brec[ic * 16 + bl] - krec[ic * 16 + kt];
      if (LD) ds[3] = brec[id * 16 + bl] - krec[id * 16 + kt];
      const double r = prim_quartet<LA, LB, LC, LD, false, kBoysLoopsOnly>(bra, ket, ds, dl, c_boys, 0.0, nullptr);
      det_add_shared(accfx + 2 * owner, r);
    }
    __syncthreads();
    v += (double)(long long)accfx[2 * tid] * (1.0 / 16777216.0) + (double)(long long)accfx[2 * tid + 1] * (1.0 / 147573952589676412928.0);
  }
  __stcs(store + (long long)blockIdx.x * kTile + tid, v);      // streaming store, see k_eri_tiles
}

// ---------------------------------------------------------------------------------------------
// ERI tiles with WARP-BATCHED Boys evaluation (opt-in: DQ_ERI_KERNEL=batched; measured SLOWER than the plain kernel, see
// eri_launch - kept for A/B runs).  ncu on the dense (H2O)_32 cluster: 79 % of the primitive
// quartets have t >= 50 (closed form), but 46 % of a warp's primitive-quartet steps contain at least one lane below it, and
// then the whole warp walks exp(-t) and the series / continued-fraction loops of every order with 14.7 of its 32 lanes
// active (2.9 in the series loops): 45 % of all warp instructions at < 50 % lane use.  The slow evaluations are therefore
// GATHERED: for one bra primitive a lane computes t for up to 9 ket primitives, the lanes with t < 50 append (ordered
// ballot) their t to a per-warp task list in shared memory, all 32 lanes then evaluate the listed tasks - whoever they
// belong to - with boys_ref and leave F_0..F_L in shared memory, and finally every lane runs the recurrences of its own
// quartets with the closed form or the values another lane computed for it.  Same arithmetic on the same t: the results are
// the unbatched kernel's; only the lane that executes the Boys loops changes.
// ---------------------------------------------------------------------------------------------
constexpr int kBoysCap = 128;      // slow Boys tasks a warp gathers at most before evaluating them (ids are 7 bits)
constexpr int kBoysBatch = 9;      // ket primitives per gathering round (their t stay in registers)

template <bool LA, bool LB, bool LC, bool LD, bool ZS>
__global__ void __launch_bounds__(256) k_eri_tiles_b(const Tile* __restrict__ tiles, SystemDev s, const double* __restrict__ pd,
                                                     double* __restrict__ store) {
  constexpr int L = (int)LA + (int)LB + (int)LC + (int)LD;
  extern __shared__ double sm[];
  if (tile_screened(s, tiles[blockIdx.x])) {
    store[(long long)blockIdx.x * kTile + threadIdx.x] = 0.0;
    if (threadIdx.x == 0) atomicAdd(s.skipped, 1ULL);
    return;
  }
  const TileCtx c = tile_ctx(tiles[blockIdx.x], s);
  const int KKb = c.b1.KK, KKk = c.b2.KK;
  double* sbra = sm;
  double* sket = sm + KKb * kPairFields * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tin = sket + KKk * kPairFields * 16 + warp * (kBoysCap * (L + 2));      // [kBoysCap] arguments of the warp's tasks
  double* res = tin + kBoysCap;                                                    // [kBoysCap][L + 1] their Boys values
  stage_pairs(c, pd, sbra, sket);
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  const int ia = c.b1.axI, ib = c.b1.axJ, ic = c.b2.axI, id = c.b2.axJ;
  const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
  double v = 0.0;
  for (int pb = 0; pb < KKb; ++pb) {
    const double* brec = sbra + pb * kPairFields * 16;
    const PairPrim bra = load_pair(brec, c.bl);
    const bool bon = c.valid && !(ZS && bra.K == 0.0);      // idle lanes stay in the loops: the warp gathers its tasks together
    double bs[4] = {0.0, 0.0, 0.0, 0.0};
    if (LA) bs[0] = brec[ia * 16 + c.bl];
    if (LB) bs[1] = brec[ib * 16 + c.bl];
    if (LC) bs[2] = brec[ic * 16 + c.bl];
    if (LD) bs[3] = brec[id * 16 + c.bl];
    for (int pk0 = 0; pk0 < KKk;) {
      unsigned long long ids = 0;         // 7-bit task number of each slow ket primitive of this lane
      unsigned onmask = 0, slowmask = 0;
      int ntask = 0, nk = 0;
      // 1. Boys arguments of the next ket primitives; the slow ones go to the warp's task list.  (Plain loops, one copy of
      //    the recurrence code below: unrolling the rounds made the kernel 25 % slower - instruction cache.)
      for (; nk < kBoysBatch && pk0 + nk < KKk && ntask <= kBoysCap - 32; ++nk) {              // warp-uniform
        const double* krec = sket + (pk0 + nk) * kPairFields * 16;
        const double d[3] = {bra.P[0] - krec[c.kt], bra.P[1] - krec[16 + c.kt], bra.P[2] - krec[32 + c.kt]};
        const bool on = bon && !(ZS && krec[80 + c.kt] == 0.0);
        const double t = quartet_t(bra.g, krec[48 + c.kt], d);
        const bool slow = on && t < boys_asymptotic_from<L>();
        const unsigned bal = __ballot_sync(0xffffffffu, slow);
        if (slow) {
          const int tk = ntask + __popc(bal & lt);
          tin[tk] = t;
          ids |= (unsigned long long)tk << (7 * nk);
          slowmask |= 1u << nk;
        }
        if (on) onmask |= 1u << nk;
        ntask += __popc(bal);
      }
      __syncwarp();
      // 2. all lanes evaluate the gathered tasks
      for (int q = lane; q < ntask; q += 32) {
        double F[L + 1], dF[L + 1];
        boys_ref<L, false>(tin[q], c_boys, F, dF);
#pragma unroll
        for (int m = 0; m <= L; ++m) res[q * (L + 1) + m] = F[m];
      }
      __syncwarp();
      // 3. the quartets themselves
      for (int j = 0; j < nk; ++j) {
        if ((onmask >> j) & 1u) {
          const double* krec = sket + (pk0 + j) * kPairFields * 16;
          const PairPrim ket = load_pair(krec, c.kt);
          double ds[4] = {0.0, 0.0, 0.0, 0.0};
          if (LA) ds[0] = bs[0] - krec[ia * 16 + c.kt];
          if (LB) ds[1] = bs[1] - krec[ib * 16 + c.kt];
          if (LC) ds[2] = bs[2] - krec[ic * 16 + c.kt];
          if (LD) ds[3] = bs[3] - krec[id * 16 + c.kt];
          double F[L + 1];
          if ((slowmask >> j) & 1u) {
            const double* r = res + (int)((ids >> (7 * j)) & 127u) * (L + 1);
#pragma unroll
            for (int m = 0; m <= L; ++m) F[m] = r[m];
          } else {
            boys_asymptotic<L, false>(prim_quartet_t(bra, ket), c_boys, F, nullptr);
          }
          v += prim_quartet_F<LA, LB, LC, LD, false>(bra, ket, ds, dl, F, nullptr, 0.0, nullptr);
        }
      }
      __syncwarp();          // the task list is rewritten by the next round
      pk0 += nk;
    }
  }
  __stcs(store + (long long)blockIdx.x * kTile + threadIdx.x, v);      // streaming store, see k_eri_tiles
}

// ---------------------------------------------------------------------------------------------
// ERI tiles for systems in which many Gaussian products underflow (selected by the host from the count k_pairs makes).
// With the per-lane skip of k_eri_tiles, lanes whose primitive pair vanished idle beside lanes whose pair survived (a
// block holds O 1s next to H 1s).  Here the CTA first COMPACTS: the surviving (lane, primitive) entries of the bra and of
// the ket are listed (ordered ballots: deterministic), and the primitive quartets of the tile are the Cartesian product
// of the two lists.  A thread owns ONE ket entry (record in registers) and walks a contiguous share of the bra list - the
// threads of a warp walk the same share, so the bra records are broadcast loads - adding each primitive integral to a
// private shared-memory cell per bra lane; a last pass sums, for every quartet, the cells of the ket entries of its lane.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int compact_entries(const double* __restrict__ srec, int KK, int nI, int nJ, unsigned short* list,
                                               int* lane_start, int* s_wc, int tid) {
  // entries e = lane*KK + pp, lane-major; survivors (valid lane, K != 0) are written in order as (lane << 8 | pp);
  // lane_start[l] = number of
  // survivors with lane < l (lane_start[16] = total).  Returns the total (valid after the final barrier).
  const int ne = 16 * KK;
  int base = 0;
  for (int e0 = 0; e0 < ne; e0 += 256) {
    const int e = e0 + tid;
    const int lane = e < ne ? e / KK : 0, pp = e - lane * KK;
    const bool alive = e < ne && (lane >> 2) < nI && (lane & 3) < nJ && srec[pp * kPairFields * 16 + 80 + lane] != 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, alive);
    if ((tid & 31) == 0) s_wc[tid >> 5] = __popc(bal);
    __syncthreads();
    int pre = base;
    for (int w = 0; w < (tid >> 5); ++w) pre += s_wc[w];
    const int pos = pre + __popc(bal & ((1u << (tid & 31)) - 1u));
    if (alive) list[pos] = (unsigned short)((lane << 8) | pp);      // (pp < 256: the staging limit keeps KK <= 97)
    if (e < ne && pp == 0) lane_start[lane] = pos;
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += s_wc[w];
    base += tot;
    __syncthreads();
  }
  if (tid == 0) lane_start[16] = base;
  return base;
}

template <bool LA, bool LB, bool LC, bool LD>
__global__ void __launch_bounds__(256, 3) k_eri_tiles_compact(const Tile* __restrict__ tiles, SystemDev s,
                                                           const double* __restrict__ pd, double* __restrict__ store) {
  extern __shared__ double sm[];
  __shared__ int s_wc[8];
  __shared__ int s_bstart[17], s_kstart[17];
  if (tile_screened(s, tiles[blockIdx.x])) {
    store[(long long)blockIdx.x * kTile + threadIdx.x] = 0.0;
    if (threadIdx.x == 0) atomicAdd(s.skipped, 1ULL);
    return;
  }
  const TileCtx c = tile_ctx(tiles[blockIdx.x], s);
  const int tid = threadIdx.x;
  const int KKb = c.b1.KK, KKk = c.b2.KK;
  double* sbra = sm;
  double* sket = sbra + KKb * kPairFields * 16;
  double* acc = sket + KKk * kPairFields * 16;                  // [16 bra lanes][256 threads]
  unsigned short* lb = (unsigned short*)(acc + 16 * 256);       // [16 * KKb]
  unsigned short* lk = lb + 16 * KKb;                           // [16 * KKk]
  stage_pairs(c, pd, sbra, sket);
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r * 256 + tid] = 0.0;
  __syncthreads();
  const int NB = compact_entries(sbra, KKb, c.b1.nI, c.b1.nJ, lb, s_bstart, s_wc, tid);
  const int NK = compact_entries(sket, KKk, c.b2.nI, c.b2.nJ, lk, s_kstart, s_wc, tid);
  __syncthreads();
  double val = 0.0;                     // quartet (bra lane c.bl, ket lane c.kt) of this thread
  if (NB > 0 && NK > 0) {
    const int ia = c.b1.axI, ib = c.b1.axJ, ic = c.b2.axI, id = c.b2.axJ;
    const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
    for (int kc0 = 0; kc0 < NK; kc0 += 256) {
      const int nkc = NK - kc0 < 256 ? NK - kc0 : 256;
      const int parts = 256 / nkc;
      const int part = tid / nkc, ikl = tid - part * nkc;
      if (part < parts) {
        const int ek = lk[kc0 + ikl];
        const int ktl = ek >> 8, pk = ek & 255;
        const double* krec = sket + pk * kPairFields * 16;
        const PairPrim ket = load_pair(krec, ktl);
        double ks[4] = {0.0, 0.0, 0.0, 0.0};
        if (LA) ks[0] = krec[ia * 16 + ktl];
        if (LB) ks[1] = krec[ib * 16 + ktl];
        if (LC) ks[2] = krec[ic * 16 + ktl];
        if (LD) ks[3] = krec[id * 16 + ktl];
        const int share = (NB + parts - 1) / parts;
        const int b0 = part * share, b1e = b0 + share < NB ? b0 + share : NB;
        int curbl = -1;
        double cur = 0.0;
        for (int ibx = b0; ibx < b1e; ++ibx) {
          const int eb = lb[ibx];
          const int bll = eb >> 8, pb = eb & 255;
          const double* brec = sbra + pb * kPairFields * 16;
          const PairPrim bra = load_pair(brec, bll);
          double ds[4] = {0.0, 0.0, 0.0, 0.0};
          if (LA) ds[0] = brec[ia * 16 + bll] - ks[0];
          if (LB) ds[1] = brec[ib * 16 + bll] - ks[1];
          if (LC) ds[2] = brec[ic * 16 + bll] - ks[2];
          if (LD) ds[3] = brec[id * 16 + bll] - ks[3];
          const double v = prim_quartet<LA, LB, LC, LD, false>(bra, ket, ds, dl, c_boys, 0.0, nullptr);
          if (bll != curbl) {
            if (curbl >= 0) acc[curbl * 256 + tid] += cur;
            curbl = bll; cur = 0.0;
          }
          cur += v;
        }
        if (curbl >= 0) acc[curbl * 256 + tid] += cur;
      }
      __syncthreads();
      {     // this ket chunk's contribution to the quartet of this thread
        const int lo = (s_kstart[c.kt] > kc0 ? s_kstart[c.kt] : kc0) - kc0;
        const int hi = (s_kstart[c.kt + 1] < kc0 + nkc ? s_kstart[c.kt + 1] : kc0 + nkc) - kc0;
        for (int p = 0; p < parts; ++p)
          for (int x = lo; x < hi; ++x) val += acc[c.bl * 256 + p * nkc + x];
      }
      __syncthreads();
      if (kc0 + 256 < NK) {
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r * 256 + tid] = 0.0;
        __syncthreads();
      }
    }
  }
  __stcs(store + (long long)blockIdx.x * kTile + tid, val);      // streaming store, see k_eri_tiles
}

// tile store -> reference order: Eri_vec[eri_index(i,j,k,l)] with ORIGINAL function indices (Tools.py:25-51)
__global__ void __launch_bounds__(256) k_scatter_packed(const Tile* __restrict__ tiles, SystemDev s, const int* __restrict__ f_orig,
                                                        const double* __restrict__ store, double* __restrict__ packed) {
  const TileCtx c = tile_ctx(tiles[blockIdx.x], s);
  if (!c.valid) return;
  const int si = c.b1.sI + c.il, sj = c.b1.sJ + c.jl, sk = c.b2.sI + c.kl, sl = c.b2.sJ + c.ll;
  const long long n = s.N;
  // one canonical representative per unique quartet: only diagonal tiles hold a quartet more than once
  if (c.b1.bI == c.b1.bJ && si > sj) return;
  if (c.b2.bI == c.b2.bJ && sk > sl) return;
  if (tiles[blockIdx.x].bp1 == tiles[blockIdx.x].bp2) {
    const int a0 = min(si, sj), a1 = max(si, sj), b0 = min(sk, sl), b1 = max(sk, sl);
    if (a0 > b0 || (a0 == b0 && a1 > b1)) return;
  }
  int i = f_orig[si], j = f_orig[sj], k = f_orig[sk], l = f_orig[sl];
  if (i > j) { int t = i; i = j; j = t; }
  if (k > l) { int t = k; k = l; l = t; }
  long long x = i * n - (long long)i * (i + 1) / 2 + j, y = k * n - (long long)k * (k + 1) / 2 + l;
  if (x > y) { long long t = x; x = y; y = t; }
  const long long m = n * (n + 1) / 2;
  packed[x * m - x * (x + 1) / 2 + y] = store[(long long)blockIdx.x * kTile + threadIdx.x];
}

// J/K contraction of the stored tiles with a symmetric matrix D (Energy.py:29-38 without the N^4 index arithmetic).
// One CTA walks a CHUNK of consecutive tiles that share the bra block pair (I,J):
//   J[i,j]  += fKL sum_kl D_kl v      accumulated in a register per thread for the whole chunk
//   J[k,l]  += fIJ sum_ij D_ij v      one global atomic per (tile, kl)
//   K[i,k], K[j,k], K[i,l], K[j,l]    accumulated in two shared-memory row strips K[I,:], K[J,:], flushed once
// with D[I,:], D[J,:] staged in shared memory.  Diagonal tiles carry weights 1/2 so that every tile is treated
// uniformly (final J: upper block triangle mirrored; final K = Kacc + Kacc^T; DESIGN.md "Fock build").
struct JkDesc { int sK, sL; short nK, nL; unsigned char same_bp, dKL, skip, pad_; };   // ket block pair of one tile, 16 bytes

// ---------------------------------------------------------------------------------------------
// J/K contraction, column-owner form: no per-tile CTA barrier, every WARP streams its own share of the chunk, two tiles
// per step, one LANE per ket pair (k,l) of a tile.  The lane reads the 16 bra-pair values of its column (coalesced: 16 lanes x 8 B per row of
// the tile) and owns them in registers:
//   J[k,l]   = sum_ij D_ij v            complete in the lane: one global atomic
//   J[i,j]  += D_kl v                   16 register accumulators kept for the whole chunk, reduced once at the end
//   K[i,k], K[i,l], K[j,k], K[j,l]      4 partial sums each, reduce-scattered over the 4 lanes that share k (or l) with
//                                       3 shuffles, so that every lane ends with ONE complete value and adds it to the
//                                       shared-memory row strips K[I,:], K[J,:]
// 96 FMAs per lane and tile-column: 1536 per tile as before, now spread over all lanes with the loads of 16 warps in
// flight per SM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double reduce_scatter4(double a0, double a1, double a2, double a3, int hi, int lo, int mhi, int mlo) {
  // 4 lanes differing in the lane bits (mhi, mlo) hold 4 partial sums each; returns, in the lane with (hi, lo), the
  // complete sum number 2*hi + lo
  const double s0 = hi ? a0 : a2, s1 = hi ? a1 : a3;
  double k0 = hi ? a2 : a0, k1 = hi ? a3 : a1;
  k0 += __shfl_xor_sync(0xffffffffu, s0, mhi);
  k1 += __shfl_xor_sync(0xffffffffu, s1, mhi);
  const double s = lo ? k0 : k1;
  double k = lo ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, s, mlo);
  return k;
}

__global__ void __launch_bounds__(256, 2) k_jk_cols(const Tile* __restrict__ tiles, const int2* __restrict__ chunks, SystemDev s,
                                                    const double* __restrict__ store, const double* __restrict__ D,
                                                    double* __restrict__ Jacc, double* __restrict__ Kacc) {
  extern __shared__ double sm[];
  const int n = s.N, tid = threadIdx.x;
  const int ns = n | 1;
  const int det = s.det;
  const int nk = (det ? 16 : 8) * ns;   // K strips: doubles, or pairs of integer limbs (fixed-point accumulation)
  double* dstrip = sm;                  // [8][ns]  rows 0-3: D[I rows, :], rows 4-7: D[J rows, :]
  double* kstrip = sm + 8 * ns;         // [8][ns] (x 2 limbs)
  double* sdij = kstrip + nk;           // [16]
  JkDesc* desc = (JkDesc*)(sdij + 16);  // [256]
  const int2 ch = chunks[blockIdx.x];
  const int bp1 = tiles[ch.x].bp1;
  const BlockPair b1 = s.bp[bp1];
  const bool dIJ = b1.bI == b1.bJ;
  if (tid < ch.y) {
    const int bp2 = tiles[ch.x + tid].bp2;
    const BlockPair b2 = s.bp[bp2];
    JkDesc d; d.sK = b2.sI; d.sL = b2.sJ; d.nK = (short)b2.nI; d.nL = (short)b2.nJ; d.same_bp = bp2 == bp1; d.dKL = b2.bI == b2.bJ;
    d.skip = tile_screened(s, tiles[ch.x + tid]); d.pad_ = 0;
    desc[tid] = d;
  }
  for (int x = tid; x < 8 * ns; x += 256) {
    const int row = x / ns, col = x - row * ns, r = row & 3;
    const bool isJ = row >= 4;
    const bool ok = r < (isJ ? b1.nJ : b1.nI) && col < n;
    dstrip[x] = ok ? D[((isJ ? b1.sJ : b1.sI) + r) * n + col] : 0.0;
  }
  for (int x = tid; x < nk; x += 256) kstrip[x] = 0.0;        // (all-zero bits are the integer 0 as well)
  if (tid < 16) sdij[tid] = ((tid >> 2) < b1.nI && (tid & 3) < b1.nJ) ? D[(b1.sI + (tid >> 2)) * n + b1.sJ + (tid & 3)] : 0.0;
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, h = lane >> 4, kl = lane & 15, k = kl >> 2, l = kl & 3;
  int per = (ch.y + 7) >> 3;
  per += per & 1;                       // whole two-tile steps per warp
  const int tb = warp * per, te = min(ch.y, tb + per);
  double jacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) jacc[q] = 0.0;
  for (int t0 = tb; t0 < te; t0 += 2) {
    const int t = t0 + h;
    JkDesc d = desc[t < te ? t : tb];
    const bool act = t < te && !d.skip;
    const bool inb = act && k < d.nK && l < d.nL;
    double v[16];
    const double* src = store + (long long)(ch.x + t) * kTile + kl;
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = act ? src[q * 16] : 0.0;
    const double dkl = inb ? D[(d.sK + k) * n + d.sL + l] : 0.0;
    const double wS = d.same_bp ? 0.5 : 1.0;
    const bool dKL = d.dKL != 0;
    const double jw = wS * (dKL ? 1.0 : 2.0) * dkl;
    double jkl = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { jacc[q] = fma(jw, v[q], jacc[q]); jkl = fma(sdij[q], v[q], jkl); }
    if (inb && jkl != 0.0) acc_add(Jacc, (long long)(d.sK + k) * n + d.sL + l, wS * (dIJ ? 1.0 : 2.0) * jkl, det);
    // exchange: D of the OTHER bra block at the lane's ket column / row
    const int ck = d.sK + k, cl = d.sL + l;
    // columns of ragged blocks (k >= nK, l >= nL) lie outside the strip rows: their tile values are zero, but 0 x (whatever
    // bits sit there - integer limbs in fixed-point mode) may be NaN, so the READS are redirected to a valid column
    const int ckr = k < d.nK ? ck : d.sK, clr = l < d.nL ? cl : d.sL;
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double dJl = dstrip[(4 + j) * ns + clr], dJk = dstrip[(4 + j) * ns + ckr];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = fma(dJl, v[i * 4 + j], a[i]); b[i] = fma(dJk, v[i * 4 + j], b[i]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double dIl = dstrip[i * ns + clr], dIk = dstrip[i * ns + ckr];
#pragma unroll
      for (int j = 0; j < 4; ++j) { c[j] = fma(dIl, v[i * 4 + j], c[j]); e[j] = fma(dIk, v[i * 4 + j], e[j]); }
    }
    const double w = wS * (dIJ ? 0.5 : 1.0) * (dKL ? 0.5 : 1.0);
    // K[i,k] (sum over j,l): lanes sharing k differ in l = lane bits 1,0 -> this lane ends with i = l
    const double kik = reduce_scatter4(a[0], a[1], a[2], a[3], l >> 1, l & 1, 2, 1);
    // K[i,l] (sum over j,k): lanes sharing l differ in k = lane bits 3,2 -> i = k
    const double kil = reduce_scatter4(b[0], b[1], b[2], b[3], k >> 1, k & 1, 8, 4);
    const double kjk = reduce_scatter4(c[0], c[1], c[2], c[3], l >> 1, l & 1, 2, 1);     // j = l
    const double kjl = reduce_scatter4(e[0], e[1], e[2], e[3], k >> 1, k & 1, 8, 4);     // j = k
    if (act) {
      if (k < d.nK) {
        if (kik != 0.0) acc_add(kstrip, l * ns + ck, w * kik, det);
        if (kjk != 0.0) acc_add(kstrip, (4 + l) * ns + ck, w * kjk, det);
      }
      if (l < d.nL) {
        if (kil != 0.0) acc_add(kstrip, k * ns + cl, w * kil, det);
        if (kjl != 0.0) acc_add(kstrip, (4 + k) * ns + cl, w * kjl, det);
      }
    }
  }
  // J[i,j]: sum the 16 accumulators over the 32 lanes of the warp
#pragma unroll
  for (int q = 0; q < 16; ++q) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) jacc[q] += __shfl_xor_sync(0xffffffffu, jacc[q], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if ((q >> 2) < b1.nI && (q & 3) < b1.nJ && jacc[q] != 0.0) acc_add(Jacc, (long long)(b1.sI + (q >> 2)) * n + b1.sJ + (q & 3), jacc[q], det);
  }
  __syncthreads();
  for (int x = tid; x < 8 * ns; x += 256) {
    const int row = x / ns, col = x - row * ns, r = row & 3;
    const long long g = (long long)((row >= 4 ? b1.sJ : b1.sI) + r) * n + col;
    if (det) {          // the strip's limbs go to the global limbs as they are: integer adds, nothing is re-rounded
      const unsigned long long* sp = (const unsigned long long*)kstrip + 2 * x;
      unsigned long long* gp = (unsigned long long*)Kacc + 2 * g;
      if (sp[0]) atomicAdd(gp, sp[0]);
      if (sp[1]) atomicAdd(gp + 1, sp[1]);
    } else {
      const double val = kstrip[x];
      if (val != 0.0) atomicAdd(Kacc + g, val);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// J/K contraction, bulk-async form (the default): the same chunks, strips, lane math and weights as k_jk_cols, but the tiles
// do not go through the register file on their way in.  Every WARP owns a ring of kJkSlots slots of two tiles (4 KB) in
// shared memory; lane 0 arms the slot's mbarrier with the byte count (mbarrier.arrive.expect_tx) and issues ONE 1-D TMA
// copy per slot (cp.async.bulk.shared::cluster.global, evict-first L2 policy: the 2.6 GB tile stream is read once per
// build and must not push D and the strips out of L2); the warp waits on the barrier's phase bit, pulls its 16 values per
// lane out of the slot, and lane 0 re-arms the slot with the copy of the stage kJkSlots ahead BEFORE the 96 FMAs of the
// current one.  No CTA barrier and no cross-warp hand-over inside a chunk: a slot is produced and consumed by one warp, so
// "slot free" is a __syncwarp.  16 warps per SM keep 16 x kJkSlots x 4 KB = 128 KB of copies in flight per SM, against the
// ~45 KB that 6.5 TB/s x ~1 us of loaded HBM latency / 148 SMs asks for.
// ---------------------------------------------------------------------------------------------
constexpr int kJkSlots = 2;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DQ_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DQ_DONE_%=;\n"
      "bra DQ_WAIT_%=;\n"
      "DQ_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// one 1-D bulk copy global -> shared of `bytes` (multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

// PRIV: every warp accumulates its exchange contributions in a PRIVATE pair of K row strips (plain shared-memory
// read-modify-writes in program order: no atomics, no fixed-point conversion in the inner loop, and - the order being
// fixed - bit-reproducible as it stands); the eight strips are summed in warp order at the end of the chunk and flushed
// once.  Needs 8 x 8 x N doubles of shared memory on top of the ring (N = 224: 198 KB, one CTA per SM).  Opt-in
// (DQ_JK_PRIVATE=1): with the shared strips the kernel is bound by instruction issue, not by HBM (356 warp instructions per
// 2 KB tile, a third of them the compare-and-swap loops of 8 shared-memory atomics per lane), but eight warps per SM do not
// hide the latency of a stage: measured slower (0.96 against 0.74 ms per build).
template <bool PRIV>
__global__ void __launch_bounds__(256, PRIV ? 1 : 2) k_jk_stream(const Tile* __restrict__ tiles, const int2* __restrict__ chunks, SystemDev s,
                                                      const double* __restrict__ store, const double* __restrict__ D,
                                                      double* __restrict__ Jacc, double* __restrict__ Kacc) {
  extern __shared__ __align__(128) double sm[];
  const int n = s.N, tid = threadIdx.x;
  const int ns = n | 1;
  const int det = s.det;
  const int nk = PRIV ? 64 * ns : (det ? 16 : 8) * ns;     // PRIV: [8 warps][8][ns] doubles
  double* ring = sm;                                        // [8 warps][kJkSlots][2 tiles][256]  (first: 128-byte aligned)
  unsigned long long* bars = (unsigned long long*)(ring + 8 * kJkSlots * 2 * kTile);      // [8][kJkSlots]
  JkDesc* desc = (JkDesc*)(bars + 8 * kJkSlots);            // [256]
  double* sdij = (double*)(desc + 256);                     // [16]
  double* dstrip = sdij + 16;                               // [8][ns]
  double* kstrip = dstrip + 8 * ns;                         // [8][ns] (x 2 limbs), PRIV: [8 warps][8][ns]
  const int2 ch = chunks[blockIdx.x];
  const int bp1 = tiles[ch.x].bp1;
  const BlockPair b1 = s.bp[bp1];
  const bool dIJ = b1.bI == b1.bJ;
  const int warp = tid >> 5, lane = tid & 31, h = lane >> 4, kl = lane & 15, k = kl >> 2, l = kl & 3;
  // stages of this warp: a CONTIGUOUS share j = warp * per + i of the chunk's stages (tiles 2j and 2j+1).  Consecutive tiles
  // mostly share the ket's first block K (the kets of a bra come in K-major order, 12-24 tiles per K), so a lane's K-role
  // sums K[I,K], K[J,K] stay in two registers while K does not change and reach the strip once per run instead of once
  // per tile: half of the strip atomics - and the collisions of the two half-warps of a stage on them - are gone.
  const int nstage = (ch.y + 1) >> 1;
  const int per = (nstage + 7) >> 3;
  const int ni = max(0, min(per, nstage - warp * per));
  double* myring = ring + (size_t)warp * kJkSlots * 2 * kTile;
  unsigned long long* mybar = bars + warp * kJkSlots;
  const unsigned long long pol = l2_evict_first_policy();
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < kJkSlots; ++r) mbar_init(mybar + r, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // make the initialised barriers visible to the async proxy
  }
  __syncwarp();
  auto issue = [&](int i) {          // lane 0: the copy of this warp's stage i into slot i % kJkSlots
    const int j = warp * per + i;
    const unsigned bytes = (unsigned)(min(2, ch.y - 2 * j) * kTile * sizeof(double));
    unsigned long long* bar = mybar + (i % kJkSlots);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(myring + (size_t)(i % kJkSlots) * 2 * kTile, store + (long long)(ch.x + 2 * j) * kTile, bytes, bar, pol);
  };
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < kJkSlots; ++r) if (r < ni) issue(r);
  }
  // while the first copies fly: ket descriptors, D strips, zeroed K strips
  if (tid < ch.y) {
    const int bp2 = tiles[ch.x + tid].bp2;
    const BlockPair b2 = s.bp[bp2];
    JkDesc d; d.sK = b2.sI; d.sL = b2.sJ; d.nK = (short)b2.nI; d.nL = (short)b2.nJ; d.same_bp = bp2 == bp1; d.dKL = b2.bI == b2.bJ;
    d.skip = tile_screened(s, tiles[ch.x + tid]); d.pad_ = 0;
    desc[tid] = d;
  }
  for (int x = tid; x < 8 * ns; x += 256) {
    const int row = x / ns, col = x - row * ns, r = row & 3;
    const bool isJ = row >= 4;
    const bool ok = r < (isJ ? b1.nJ : b1.nI) && col < n;
    dstrip[x] = ok ? D[((isJ ? b1.sJ : b1.sI) + r) * n + col] : 0.0;
  }
  for (int x = tid; x < nk; x += 256) kstrip[x] = 0.0;
  if (tid < 16) sdij[tid] = ((tid >> 2) < b1.nI && (tid & 3) < b1.nJ) ? D[(b1.sI + (tid >> 2)) * n + b1.sJ + (tid & 3)] : 0.0;
  __syncthreads();

  double jacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) jacc[q] = 0.0;
  double rik = 0.0, rjk = 0.0;      // K-role sums of the lane's current run: K[I_l, K_k], K[J_l, K_k]
  int runK = -1, runnK = 0;         // first function and size of that run's K block
  auto flush_run = [&]() {
    if (runK >= 0 && k < runnK) {
      if (rik != 0.0) acc_add(kstrip, l * ns + runK + k, rik, det);
      if (rjk != 0.0) acc_add(kstrip, (4 + l) * ns + runK + k, rjk, det);
    }
    rik = 0.0; rjk = 0.0;
  };
  for (int i = 0; i < ni; ++i) {
    const int t = 2 * (warp * per + i) + h;
    const JkDesc d = desc[t < ch.y ? t : ch.y - 1];
    const bool act = t < ch.y && !d.skip;
    const bool inb = act && k < d.nK && l < d.nL;
    const double dkl = inb ? D[(d.sK + k) * n + d.sL + l] : 0.0;
    const int slot = i % kJkSlots;
    mbar_wait(mybar + slot, (unsigned)((i / kJkSlots) & 1));
    double v[16];
    const double* src = myring + (size_t)slot * 2 * kTile + h * kTile + kl;
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = act ? src[q * 16] : 0.0;
    __syncwarp();                                   // every lane has its values: the slot is free
    if (lane == 0 && i + kJkSlots < ni) issue(i + kJkSlots);
    const double wS = d.same_bp ? 0.5 : 1.0;
    const bool dKL = d.dKL != 0;
    const double jw = wS * (dKL ? 1.0 : 2.0) * dkl;
    double jkl = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { jacc[q] = fma(jw, v[q], jacc[q]); jkl = fma(sdij[q], v[q], jkl); }
    if (inb && jkl != 0.0) acc_add(Jacc, (long long)(d.sK + k) * n + d.sL + l, wS * (dIJ ? 1.0 : 2.0) * jkl, det);
    const int ck = d.sK + k, cl = d.sL + l;
    // columns of ragged blocks (k >= nK, l >= nL) lie outside the strip rows: their tile values are zero, but 0 x (whatever
    // bits sit there - integer limbs in fixed-point mode) may be NaN, so the READS are redirected to a valid column
    const int ckr = k < d.nK ? ck : d.sK, clr = l < d.nL ? cl : d.sL;
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double dJl = dstrip[(4 + j) * ns + clr], dJk = dstrip[(4 + j) * ns + ckr];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) { a[ii] = fma(dJl, v[ii * 4 + j], a[ii]); b[ii] = fma(dJk, v[ii * 4 + j], b[ii]); }
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const double dIl = dstrip[ii * ns + clr], dIk = dstrip[ii * ns + ckr];
#pragma unroll
      for (int j = 0; j < 4; ++j) { c[j] = fma(dIl, v[ii * 4 + j], c[j]); e[j] = fma(dIk, v[ii * 4 + j], e[j]); }
    }
    const double w = wS * (dIJ ? 0.5 : 1.0) * (dKL ? 0.5 : 1.0);
    const double kik = reduce_scatter4(a[0], a[1], a[2], a[3], l >> 1, l & 1, 2, 1);
    const double kil = reduce_scatter4(b[0], b[1], b[2], b[3], k >> 1, k & 1, 8, 4);
    const double kjk = reduce_scatter4(c[0], c[1], c[2], c[3], l >> 1, l & 1, 2, 1);
    const double kjl = reduce_scatter4(e[0], e[1], e[2], e[3], k >> 1, k & 1, 8, 4);
    if (PRIV) {
      // the two tiles of the stage (half-warps) hit the same strip elements when they share the K (or the L) block:
      // then the upper half hands its weighted values to the lower half; otherwise the 32 targets are distinct
      double* kp = kstrip + (size_t)warp * 8 * ns;
      const int osK = __shfl_xor_sync(0xffffffffu, d.sK, 16), osL = __shfl_xor_sync(0xffffffffu, d.sL, 16);
      const bool oact = __shfl_xor_sync(0xffffffffu, (int)act, 16) != 0;
      const bool sameK = act && oact && osK == d.sK, sameL = act && oact && osL == d.sL;
      double x0 = (act && k < d.nK) ? w * kik : 0.0, x1 = (act && k < d.nK) ? w * kjk : 0.0;
      double y0 = (act && l < d.nL) ? w * kil : 0.0, y1 = (act && l < d.nL) ? w * kjl : 0.0;
      const double ox0 = __shfl_xor_sync(0xffffffffu, x0, 16), ox1 = __shfl_xor_sync(0xffffffffu, x1, 16);
      const double oy0 = __shfl_xor_sync(0xffffffffu, y0, 16), oy1 = __shfl_xor_sync(0xffffffffu, y1, 16);
      if (sameK) { x0 = h ? 0.0 : x0 + ox0; x1 = h ? 0.0 : x1 + ox1; }
      if (sameL) { y0 = h ? 0.0 : y0 + oy0; y1 = h ? 0.0 : y1 + oy1; }
      if (x0 != 0.0) kp[l * ns + ck] += x0;
      if (x1 != 0.0) kp[(4 + l) * ns + ck] += x1;
      __syncwarp();                     // a diagonal ket (K == L) makes the next two adds revisit these elements
      if (y0 != 0.0) kp[k * ns + cl] += y0;
      if (y1 != 0.0) kp[(4 + k) * ns + cl] += y1;
      __syncwarp();
    } else if (act) {
      if (d.sK != runK) { flush_run(); runK = d.sK; runnK = d.nK; }
      if (k < d.nK) { rik = fma(w, kik, rik); rjk = fma(w, kjk, rjk); }
      if (l < d.nL) {
        if (kil != 0.0) acc_add(kstrip, k * ns + cl, w * kil, det);
        if (kjl != 0.0) acc_add(kstrip, (4 + k) * ns + cl, w * kjl, det);
      }
    }
  }
  if (!PRIV) flush_run();
#pragma unroll
  for (int q = 0; q < 16; ++q) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) jacc[q] += __shfl_xor_sync(0xffffffffu, jacc[q], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if ((q >> 2) < b1.nI && (q & 3) < b1.nJ && jacc[q] != 0.0) acc_add(Jacc, (long long)(b1.sI + (q >> 2)) * n + b1.sJ + (q & 3), jacc[q], det);
  }
  __syncthreads();
  for (int x = tid; x < 8 * ns; x += 256) {
    const int row = x / ns, col = x - row * ns, r = row & 3;
    const long long g = (long long)((row >= 4 ? b1.sJ : b1.sI) + r) * n + col;
    if (PRIV) {
      double val = 0.0;
#pragma unroll
      for (int wp = 0; wp < 8; ++wp) val += kstrip[(size_t)wp * 8 * ns + x];      // fixed order
      if (val != 0.0 && col < n && r < (row >= 4 ? b1.nJ : b1.nI)) acc_add(Kacc, g, val, det);
    } else if (det) {
      const unsigned long long* sp = (const unsigned long long*)kstrip + 2 * x;
      unsigned long long* gp = (unsigned long long*)Kacc + 2 * g;
      if (sp[0]) atomicAdd(gp, sp[0]);
      if (sp[1]) atomicAdd(gp + 1, sp[1]);
    } else {
      const double val = kstrip[x];
      if (val != 0.0) atomicAdd(Kacc + g, val);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Integral-direct J/K (dq_options.jk_mode = 2; Energy.py:29-38 with no stored Eri vector): ONE kernel per angular class
// evaluates the tiles of a chunk (same bra block pair) exactly as k_eri_tiles does and contracts each tile's 256 ERIs with
// D right after its primitive loop - the tile never leaves the SM.  The thread of quartet (ij|kl) drops its value into a
// double-buffered 2 KB shared-memory tile; after the next CTA barrier (the one the ket staging of the following tile needs
// anyway) 16 lanes of warp 0 run the lane math of k_jk_cols on it - lane = ket pair for J[kl] and the four exchange pieces,
// lane = bra pair for J[ij] - while the other warps already stage the next ket: ~150 instructions of one half-warp per
// tile against 256 x 81 primitive quartets.  The bra records, D[I,:], D[J,:] and the K row strips are staged once per chunk.
// ---------------------------------------------------------------------------------------------
template <bool LA, bool LB, bool LC, bool LD, bool ZS>
__global__ void __launch_bounds__(256) k_jk_direct(const Tile* __restrict__ tiles, const int2* __restrict__ chunks, SystemDev s,
                                                   const double* __restrict__ pd, const double* __restrict__ D,
                                                   double* __restrict__ Jacc, double* __restrict__ Kacc, int kkmax) {
  extern __shared__ double sm[];
  const int n = s.N, tid = threadIdx.x;
  const int ns = n | 1;
  const int det = s.det;
  const int nk = (det ? 16 : 8) * ns;
  double* sbra = sm;                                   // [kkmax * 8 * 16]
  double* sket = sbra + kkmax * kPairFields * 16;      // [kkmax * 8 * 16]
  double* sv = sket + kkmax * kPairFields * 16;        // [2][16 * 17]  the tile being digested / the tile being filled
  double* sdij = sv + 2 * 16 * 17;                     // [16]
  double* dstrip = sdij + 16;                          // [8][ns]
  double* kstrip = dstrip + 8 * ns;                    // [8][ns] (x 2 limbs)
  const int2 ch = chunks[blockIdx.x];
  const int bp1 = tiles[ch.x].bp1;
  const BlockPair b1 = s.bp[bp1];
  const bool dIJ = b1.bI == b1.bJ;
  const int il = tid >> 6, jl = (tid >> 4) & 3, kl4 = (tid >> 2) & 3, ll = tid & 3, bl = tid >> 4, kt = tid & 15;
  {
    const double* gb = pd + b1.off * 16;
    for (int x = tid; x < b1.KK * kPairFields * 16; x += 256) sbra[x] = gb[x];
  }
  for (int x = tid; x < 8 * ns; x += 256) {
    const int row = x / ns, col = x - row * ns, r = row & 3;
    const bool isJ = row >= 4;
    const bool ok = r < (isJ ? b1.nJ : b1.nI) && col < n;
    dstrip[x] = ok ? D[((isJ ? b1.sJ : b1.sI) + r) * n + col] : 0.0;
  }
  for (int x = tid; x < nk; x += 256) kstrip[x] = 0.0;
  if (tid < 16) sdij[tid] = ((tid >> 2) < b1.nI && (tid & 3) < b1.nJ) ? D[(b1.sI + (tid >> 2)) * n + b1.sJ + (tid & 3)] : 0.0;
  const int ia = b1.axI, ib = b1.axJ;
  double jij = 0.0;                 // lanes 0-15 of warp 0: J[i,j] of bra pair `tid`, accumulated over the chunk

  // digest of one finished tile: the whole warp 0 calls it (the shuffles name all 32 lanes); lanes 16-31 mirror lanes 0-15 and
  // add nothing.  lane = ket pair kl for J[kl] and the exchange pieces, then lane = bra pair ij for J[ij].
  auto digest = [&](int t) {
    const Tile tl = tiles[ch.x + t];
    if (tile_screened(s, tl)) return;
    const BlockPair b2 = s.bp[tl.bp2];
    const double* svb = sv + (t & 1) * (16 * 17);
    const int c16 = tid & 15;
    const bool on = tid < 16;
    const int k = c16 >> 2, l = c16 & 3;
    const bool inb = k < b2.nI && l < b2.nJ;
    const bool dKL = b2.bI == b2.bJ;
    const double wS = tl.bp2 == bp1 ? 0.5 : 1.0;
    const double dkl = inb ? D[(b2.sI + k) * n + b2.sJ + l] : 0.0;
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = svb[q * 17 + c16];
    double jkl = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) jkl = fma(sdij[q], v[q], jkl);
    if (on && inb && jkl != 0.0) acc_add(Jacc, (long long)(b2.sI + k) * n + b2.sJ + l, wS * (dIJ ? 1.0 : 2.0) * jkl, det);
    // J[i,j] of bra pair c16: sum over the ket pairs, the weighted D_kl of lane x fetched by shuffle
    {
      const double jw = wS * (dKL ? 1.0 : 2.0) * dkl;
      double acc = 0.0;
#pragma unroll
      for (int x = 0; x < 16; ++x) acc = fma(__shfl_sync(0xffffffffu, jw, x), svb[c16 * 17 + x], acc);
      jij += acc;
    }
    const int ck = b2.sI + k, cl = b2.sJ + l;
    const int ckr = k < b2.nI ? ck : b2.sI, clr = l < b2.nJ ? cl : b2.sJ;      // ragged blocks: see k_jk_cols
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double dJl = dstrip[(4 + j) * ns + clr], dJk = dstrip[(4 + j) * ns + ckr];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = fma(dJl, v[i * 4 + j], a[i]); b[i] = fma(dJk, v[i * 4 + j], b[i]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double dIl = dstrip[i * ns + clr], dIk = dstrip[i * ns + ckr];
#pragma unroll
      for (int j = 0; j < 4; ++j) { c[j] = fma(dIl, v[i * 4 + j], c[j]); e[j] = fma(dIk, v[i * 4 + j], e[j]); }
    }
    const double w = wS * (dIJ ? 0.5 : 1.0) * (dKL ? 0.5 : 1.0);
    const double kik = reduce_scatter4(a[0], a[1], a[2], a[3], l >> 1, l & 1, 2, 1);
    const double kil = reduce_scatter4(b[0], b[1], b[2], b[3], k >> 1, k & 1, 8, 4);
    const double kjk = reduce_scatter4(c[0], c[1], c[2], c[3], l >> 1, l & 1, 2, 1);
    const double kjl = reduce_scatter4(e[0], e[1], e[2], e[3], k >> 1, k & 1, 8, 4);
    if (on && k < b2.nI) {
      if (kik != 0.0) acc_add(kstrip, l * ns + ck, w * kik, det);
      if (kjk != 0.0) acc_add(kstrip, (4 + l) * ns + ck, w * kjk, det);
    }
    if (on && l < b2.nJ) {
      if (kil != 0.0) acc_add(kstrip, k * ns + cl, w * kil, det);
      if (kjl != 0.0) acc_add(kstrip, (4 + k) * ns + cl, w * kjl, det);
    }
  };

  for (int t = 0; t < ch.y; ++t) {
    const Tile tl = tiles[ch.x + t];
    const BlockPair b2 = s.bp[tl.bp2];
    __syncthreads();                  // tile t-1: primitive loops done (sket free), its 256 values are in sv[(t-1) & 1]
    if (t > 0 && tid < 32) digest(t - 1);      // warp 0; the other warps go on to stage the next ket
    if (tile_screened(s, tl)) {       // CTA-uniform
      if (tid == 0) atomicAdd(s.skipped, 1ULL);
      continue;
    }
    {
      const double* gk = pd + b2.off * 16;
      for (int x = tid; x < b2.KK * kPairFields * 16; x += 256) sket[x] = gk[x];
    }
    __syncthreads();
    double v = 0.0;
    if (il < b1.nI && jl < b1.nJ && kl4 < b2.nI && ll < b2.nJ) {
      const int ic = b2.axI, id = b2.axJ;
      const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
      for (int pb = 0; pb < b1.KK; ++pb) {
        const double* brec = sbra + pb * kPairFields * 16;
        const PairPrim bra = load_pair(brec, bl);
        if (ZS && bra.K == 0.0) continue;
        double bs[4] = {0.0, 0.0, 0.0, 0.0};
        if (LA) bs[0] = brec[ia * 16 + bl];
        if (LB) bs[1] = brec[ib * 16 + bl];
        if (LC) bs[2] = brec[ic * 16 + bl];
        if (LD) bs[3] = brec[id * 16 + bl];
        for (int pk = 0; pk < b2.KK; ++pk) {
          const double* krec = sket + pk * kPairFields * 16;
          const PairPrim ket = load_pair(krec, kt);
          if (ZS && ket.K == 0.0) continue;
          double ds[4] = {0.0, 0.0, 0.0, 0.0};
          if (LA) ds[0] = bs[0] - krec[ia * 16 + kt];
          if (LB) ds[1] = bs[1] - krec[ib * 16 + kt];
          if (LC) ds[2] = bs[2] - krec[ic * 16 + kt];
          if (LD) ds[3] = bs[3] - krec[id * 16 + kt];
          v += prim_quartet<LA, LB, LC, LD, false>(bra, ket, ds, dl, c_boys, 0.0, nullptr);
        }
      }
    }
    sv[(t & 1) * (16 * 17) + bl * 17 + kt] = v;
  }
  __syncthreads();
  if (tid < 32) digest(ch.y - 1);
  if (tid < 16 && (tid >> 2) < b1.nI && (tid & 3) < b1.nJ && jij != 0.0)
    acc_add(Jacc, (long long)(b1.sI + (tid >> 2)) * n + b1.sJ + (tid & 3), jij, det);
  __syncthreads();
  for (int x = tid; x < 8 * ns; x += 256) {
    const int row = x / ns, col = x - row * ns, r = row & 3;
    const long long g = (long long)((row >= 4 ? b1.sJ : b1.sI) + r) * n + col;
    if (det) {
      const unsigned long long* sp = (const unsigned long long*)kstrip + 2 * x;
      unsigned long long* gp = (unsigned long long*)Kacc + 2 * g;
      if (sp[0]) atomicAdd(gp, sp[0]);
      if (sp[1]) atomicAdd(gp + 1, sp[1]);
    } else {
      const double val = kstrip[x];
      if (val != 0.0) atomicAdd(Kacc + g, val);
    }
  }
}

// limb pairs -> doubles: out[i] = beta * out[i] + value(in, i)
__global__ void k_det_finalize(long long n, const double* __restrict__ in, double* __restrict__ out, double beta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (beta != 0.0 ? beta * out[i] : 0.0) + det_value(in, i);
}

// ---------------------------------------------------------------------------------------------
constexpr int kDeferWords = 3;     // deferred gradient form: 3 mask words per thread and tile (+ one bin byte per marked step in dbin)
constexpr int kGradKetChunk = 9;   // ket primitive pairs whose adjoint slots are resident at once
constexpr int kSparseChunk = 20;   // tiles per chunk of the sparse-weight gradient kernel (its work list lives in shared memory)

// ERI gradient: sum over the quartets of Wsym(ijkl) * d(ij|kl)/d(primitive-pair fields)
//   Wsym = wtile * sum_terms [ 8 (A_ij B_kl + A_kl B_ij) - 2 (A_ik B_jl + A_jk B_il + A_il B_jk + A_jl B_ik) ]
// where sum_terms <A, dG[B]> is the two-electron part of dE from the SCF reverse sweep.
//
// One CTA walks a CHUNK of tiles that share the KET block pair (the gradient pass does not touch the tile store, so its
// tiles are re-ordered by ket).  The ket records are staged once and the per-thread ket-adjoint slots in shared memory
// accumulate over the whole chunk; they are summed over the 16 bra lanes and flushed once per chunk.  Inside the chunk
// there is NO CTA barrier: a warp = 2 bra lanes x 16 ket lanes reads its bra records straight from global memory (L1/L2
// resident, 16 lanes per address), reduces its bra adjoints with shuffles and moves on to the next tile on its own.
// (With one tile per CTA the 8 warps drifted apart over the 81 primitive quartets of a tile and ~10 % of the stall
// samples sat on the end-of-tile barrier.)
// ---------------------------------------------------------------------------------------------
// occupancy: two CTAs per SM for every class (128 registers, 2 x ~105 KB shared memory).  (pp|pp) wants 251 registers; with
// the rolled Boys loops of round 2 it is faster squeezed to 128 (240 bytes of spills) at two CTAs per SM: 50 against 60 ms
// (round 1, with the unrolled loops, measured no gain)
// DEFER (classes with >= 2 p functions, contractions of <= 96 primitive quartets): the loop-regime primitive quartets are
// not evaluated here - no Boys loop is compiled into this kernel - but marked, one bit per (bra primitive, ket primitive)
// step in three words per thread, and the masks go to global memory (dmask[tile][word][thread]) for k_eri_grad_deferred.
template <bool LA, bool LB, bool LC, bool LD, bool ZS, bool DEFER>
__global__ void __launch_bounds__(256, 2)
k_eri_grad_chunks(const Tile* __restrict__ tiles, const int2* __restrict__ chunks, SystemDev s, const double* __restrict__ pd,
                  int nterms, const double* __restrict__ At, const double* __restrict__ Bt, double* __restrict__ padj,
                  double* __restrict__ pasum, unsigned* __restrict__ dmask, unsigned char* __restrict__ dbin, int tile0) {
  constexpr int L = (int)LA + (int)LB + (int)LC + (int)LD;
  const double tmid = boys_midrange_from<L>();
  extern __shared__ double sm[];
  const int2 ch = chunks[blockIdx.x];
  const int bp2 = tiles[ch.x].bp2;
  const BlockPair b2 = s.bp[bp2];
  const int tid = threadIdx.x;
  const int il = tid >> 6, jl = (tid >> 4) & 3, kl = (tid >> 2) & 3, ll = tid & 3, bl = tid >> 4, kt = tid & 15;
  const int KKk = b2.KK;
  double* sket = sm;
  // ket adjoints accumulate in per-thread shared-memory slots; long contractions are walked in chunks of kGradKetChunk
  // ket primitive pairs so that the slots stay within shared memory (STO-3G: 9 pairs = one chunk)
  double* sacc = sket + KKk * kPairFields * 16;     // [(min(KKk, kGradKetChunk)*5 + 2)][256]
  const int KC = KKk < kGradKetChunk ? KKk : kGradKetChunk;
  {
    const double* gk = pd + b2.off * 16;
    for (int x = tid; x < KKk * kPairFields * 16; x += 256) sket[x] = gk[x];
  }
  const int n = s.N;
  const size_t nn = (size_t)n * n;
  const int ic = b2.axI, id = b2.axJ;
  const int k = b2.sI + kl, l = b2.sJ + ll;
  const bool kvalid = kl < b2.nI && ll < b2.nJ;
  const long long off5k = b2.off / kPairFields * kAdjFields;
  double qcs = 0.0, qds = 0.0;
  int nzero = 0;          // valid quartets of this thread whose two-particle weight is exactly zero (nothing to add)
  for (int k0 = 0; k0 < KKk; k0 += KC) {
    const int kn = KKk - k0 < KC ? KKk - k0 : KC;
    for (int x = 0; x < kn * kAdjFields; ++x) sacc[x * 256 + tid] = 0.0;
    __syncthreads();      // (first round: also publishes the staged ket records)
    for (int ti = 0; ti < ch.y; ++ti) {
      const Tile tl = tiles[ch.x + ti];
      if (tile_screened(s, tl)) continue;
      const BlockPair b1 = s.bp[tl.bp1];
      const bool valid = kvalid && il < b1.nI && jl < b1.nJ;
      const int KKb = b1.KK;
      const int ia = b1.axI, ib = b1.axJ;
      const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
      double W = 0.0;
      if (valid) {
        const int i = b1.sI + il, j = b1.sJ + jl;
        for (int t = 0; t < nterms; ++t) {
          const double* A = At + t * nn;
          const double* B = Bt + t * nn;
          W += 8.0 * (A[i * n + j] * B[k * n + l] + A[k * n + l] * B[i * n + j]) -
               2.0 * (A[i * n + k] * B[j * n + l] + A[j * n + k] * B[i * n + l] + A[i * n + l] * B[j * n + k] +
                      A[j * n + l] * B[i * n + k]);
        }
        W *= (tl.bp1 == tl.bp2 ? 0.5 : 1.0) * (b1.bI == b1.bJ ? 0.5 : 1.0) * (b2.bI == b2.bJ ? 0.5 : 1.0);
        if (k0 == 0 && W == 0.0) ++nzero;
      }
      // a warp (2 bra lanes x 16 ket lanes) none of whose quartets carries weight has nothing to add for this tile
      if (!__any_sync(0xffffffffu, W != 0.0)) continue;
      const long long off5b = b1.off / kPairFields * kAdjFields;
      const double* gbra = pd + b1.off * 16;
      double pas = 0.0, pbs = 0.0;
      unsigned m0 = 0u, m1 = 0u, m2 = 0u;      // DEFER: marked steps pb * KKk + pk
      const double binscale = 31.99 / tmid;      // ... their Boys-argument bins go straight to global memory (one byte each)
      unsigned char* tbin = DEFER ? dbin + (long long)(ch.x + ti - tile0) * kDefMaxSteps * 256 + tid : nullptr;      // (class-relative tile index)
      for (int pb = 0; pb < KKb; ++pb) {
        const double* brec = gbra + pb * kPairFields * 16;
        PairAdj ba; ba.P[0] = ba.P[1] = ba.P[2] = ba.g = ba.K = 0.0;
        // adjoints of bra.P along the axes of the four p functions, summed over the ket primitives and folded into ba.P
        // once per bra primitive (the axis is CTA-uniform: no select inside the ket loop)
        double bsl[4] = {0.0, 0.0, 0.0, 0.0};
        const PairPrim bra = load_pair(brec, bl);
        if (W != 0.0 && !(ZS && pair_underflowed(bra.K))) {
          double bs[4] = {0.0, 0.0, 0.0, 0.0};
          if (LA) bs[0] = brec[ia * 16 + bl];
          if (LB) bs[1] = brec[ib * 16 + bl];
          if (LC) bs[2] = brec[ic * 16 + bl];
          if (LD) bs[3] = brec[id * 16 + bl];
          for (int pk = 0; pk < kn; ++pk) {
            const double* krec = sket + (k0 + pk) * kPairFields * 16;
            const PairPrim ket = load_pair(krec, kt);
            if (ZS && pair_underflowed(ket.K)) continue;
            double ds[4] = {0.0, 0.0, 0.0, 0.0};
            if (LA) ds[0] = bs[0] - krec[ia * 16 + kt];
            if (LB) ds[1] = bs[1] - krec[ib * 16 + kt];
            if (LC) ds[2] = bs[2] - krec[ic * 16 + kt];
            if (LD) ds[3] = bs[3] - krec[id * 16 + kt];
            QuartetAdj q;
            if (DEFER) {
              const double tq = prim_quartet_t(bra, ket);
              if (!(tq >= tmid)) {      // loop regime: marked for the second kernel
                const int st = pb * KKk + k0 + pk;
                const unsigned bit = 1u << (st & 31);
                if (st < 32) m0 |= bit; else if (st < 64) m1 |= bit; else m2 |= bit;
                tbin[st * 256] = (unsigned char)(tq > 0.0 ? min(31, (int)(tq * binscale)) : 0);
                continue;
              }
              prim_quartet<LA, LB, LC, LD, true, kBoysNoLoops>(bra, ket, ds, dl, c_boys, W, &q);
            } else {
              prim_quartet<LA, LB, LC, LD, true>(bra, ket, ds, dl, c_boys, W, &q);
            }
            ba.P[0] += q.dP[0]; ba.P[1] += q.dP[1]; ba.P[2] += q.dP[2]; ba.g += q.gb; ba.K += q.Kb;
            if (LA) { bsl[0] += q.ds[0] + q.e[0]; pas += q.e[0]; }      // bra.pa = bra.P[ia] - A[ia]
            if (LB) { bsl[1] += q.ds[1] + q.e[1]; pbs += q.e[1]; }
            if (LC) { bsl[2] += q.ds[2]; qcs += q.e[2]; }
            if (LD) { bsl[3] += q.ds[3]; qds += q.e[3]; }
            // ket adjoints: per-thread shared-memory slots [P0 P1 P2 g K] of this ket primitive; the axis components go
            // to the row of their axis (CTA-uniform row index)
            double* a = sacc + (pk * kAdjFields) * 256 + tid;
            a[0] -= q.dP[0]; a[256] -= q.dP[1]; a[512] -= q.dP[2]; a[768] += q.gk; a[1024] += q.Kk;
            if (LA) a[ia * 256] -= q.ds[0];
            if (LB) a[ib * 256] -= q.ds[1];
            if (LC) a[ic * 256] += q.e[2] - q.ds[2];                     // ket.pa = ket.P[ic] - C[ic]
            if (LD) a[id * 256] += q.e[3] - q.ds[3];
          }
        }
        if (LA) add3(ba.P, ia, bsl[0]);
        if (LB) add3(ba.P, ib, bsl[1]);
        if (LC) add3(ba.P, ic, bsl[2]);
        if (LD) add3(ba.P, id, bsl[3]);
        // bra adjoints of this primitive pair: reduce over the 16 ket lanes of the half-warp
        double r[kAdjFields] = {ba.P[0], ba.P[1], ba.P[2], ba.g, ba.K};
#pragma unroll
        for (int f = 0; f < kAdjFields; ++f) {
#pragma unroll
          for (int o = 8; o >= 1; o >>= 1) r[f] += __shfl_xor_sync(0xffffffffu, r[f], o);
        }
        if (kt == 0) {
#pragma unroll
          for (int f = 0; f < kAdjFields; ++f)
            if (r[f] != 0.0) acc_add(padj, (off5b + (long long)pb * kAdjFields + f) * 16 + bl, r[f], s.det);
        }
      }
      if (LA || LB) {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) { pas += __shfl_xor_sync(0xffffffffu, pas, o); pbs += __shfl_xor_sync(0xffffffffu, pbs, o); }
        if (kt == 0) {
          if (pas != 0.0) acc_add(pasum, ((long long)tl.bp1 * 2 + 0) * 16 + bl, pas, s.det);
          if (pbs != 0.0) acc_add(pasum, ((long long)tl.bp1 * 2 + 1) * 16 + bl, pbs, s.det);
        }
      }
      if (DEFER && __any_sync(0xffffffffu, (m0 | m1 | m2) != 0u)) {      // (the mask buffer is zeroed by the host)
        unsigned* mp = dmask + ((long long)(ch.x + ti - tile0) * kDeferWords) * 256 + tid;
        mp[0] = m0; mp[256] = m1; mp[512] = m2;
      }
    }
    const bool last = k0 + kn >= KKk;
    int nslots = kn * kAdjFields;
    if (last) {           // the two contracted (Q-C), (Q-D) sums ride along with the last round
      sacc[(nslots + 0) * 256 + tid] = qcs;
      sacc[(nslots + 1) * 256 + tid] = qds;
      nslots += 2;
    }
    __syncthreads();
    // ket adjoints: sum over the 16 bra lanes
    for (int slot = tid >> 4; slot < nslots; slot += 16) {
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < 16; ++b) sum += sacc[slot * 256 + b * 16 + kt];
      if (sum != 0.0) {
        if (slot < kn * kAdjFields) acc_add(padj, (off5k + (long long)k0 * kAdjFields + slot) * 16 + kt, sum, s.det);
        else acc_add(pasum, ((long long)bp2 * 2 + (slot - kn * kAdjFields)) * 16 + kt, sum, s.det);
      }
    }
    __syncthreads();
  }
  if (s.zero_weight) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nzero += __shfl_xor_sync(0xffffffffu, nzero, o);
    if ((tid & 31) == 0 && nzero) atomicAdd(s.zero_weight, (unsigned long long)nzero);
  }
}

// ---------------------------------------------------------------------------------------------
// ERI gradient for SPARSE weights (selected by the host when the density matrix is mostly exact zeros: well separated
// fragments, where ~80 % of the quartets carry exactly zero weight and the survivors sit in the same few lanes of every
// tile).  Same chunks (tiles sharing the ket block pair) and the same primitive code as k_eri_grad_chunks, but the CTA
// first COMPACTS its work: every thread evaluates the weight of its quartet in every tile of the chunk and appends the
// (tile, lane) pairs with weight to a list in shared memory; then the 256 threads take list entries round robin, so all
// lanes stay busy.  A thread is no longer tied to a ket lane or a bra lane, so nothing can be reduced across threads:
// bra adjoints go to global memory with one atomic each as they are completed, ket adjoints accumulate in the thread's
// private shared-memory slots during a quartet and are flushed with atomics after it (~90 atomics per contracted
// quartet of 81 primitive quartets).  With dense weights the tile-synchronous kernel is the better one - hence two.
// ---------------------------------------------------------------------------------------------
template <bool LA, bool LB, bool LC, bool LD>
__global__ void __launch_bounds__(256, ((int)LA + (int)LB + (int)LC + (int)LD <= 3) ? 2 : 1)
k_eri_grad_sparse(const Tile* __restrict__ tiles, const int2* __restrict__ chunks, SystemDev s, const double* __restrict__ pd,
                  int nterms, const double* __restrict__ At, const double* __restrict__ Bt, double* __restrict__ padj,
                  double* __restrict__ pasum) {
  extern __shared__ double sm[];
  __shared__ int s_count;
  const int2 ch = chunks[blockIdx.x];          // at most kSparseChunk tiles
  const int bp2 = tiles[ch.x].bp2;
  const BlockPair b2 = s.bp[bp2];
  const int tid = threadIdx.x;
  const int KKk = b2.KK;
  double* sket = sm;
  double* sacc = sket + KKk * kPairFields * 16;     // [min(KKk, kGradKetChunk)*5][256] per-thread scratch
  const int KC = KKk < kGradKetChunk ? KKk : kGradKetChunk;
  unsigned short* items = (unsigned short*)(sacc + KC * kAdjFields * 256);            // [kSparseChunk * 256]
  {
    const double* gk = pd + b2.off * 16;
    for (int x = tid; x < KKk * kPairFields * 16; x += 256) sket[x] = gk[x];
  }
  if (tid == 0) s_count = 0;
  const int n = s.N;
  const size_t nn = (size_t)n * n;
  const int ic = b2.axI, id = b2.axJ;
  const long long off5k = b2.off / kPairFields * kAdjFields;
  const double wket = b2.bI == b2.bJ ? 0.5 : 1.0;
  // weight of the quartet of lane t (= il*64 + jl*16 + kl*4 + ll) in tile tl, exactly as in k_eri_grad_chunks
  auto weight = [&](const Tile tl, const BlockPair& b1, int t, bool* valid) -> double {
    const int il = t >> 6, jl = (t >> 4) & 3, kl = (t >> 2) & 3, ll = t & 3;
    *valid = kl < b2.nI && ll < b2.nJ && il < b1.nI && jl < b1.nJ;
    if (!*valid) return 0.0;
    const int i = b1.sI + il, j = b1.sJ + jl, k = b2.sI + kl, l = b2.sJ + ll;
    if (s.wmask) {       // every product of the weight carries one of these six B factors: all zero -> W is exactly 0
      const unsigned char* m = s.wmask;
      if (!(m[k * n + l] | m[i * n + j] | m[j * n + l] | m[i * n + l] | m[j * n + k] | m[i * n + k])) return 0.0;
    }
    double W = 0.0;
    for (int q = 0; q < nterms; ++q) {
      const double* A = At + q * nn;
      const double* B = Bt + q * nn;
      W += 8.0 * (A[i * n + j] * B[k * n + l] + A[k * n + l] * B[i * n + j]) -
           2.0 * (A[i * n + k] * B[j * n + l] + A[j * n + k] * B[i * n + l] + A[i * n + l] * B[j * n + k] +
                  A[j * n + l] * B[i * n + k]);
    }
    return W * ((tl.bp1 == tl.bp2 ? 0.5 : 1.0) * (b1.bI == b1.bJ ? 0.5 : 1.0) * wket);
  };
  __syncthreads();
  // phase 1: which (tile, lane) pairs carry weight
  int nzero = 0;
  for (int ti = 0; ti < ch.y; ++ti) {
    const Tile tl = tiles[ch.x + ti];
    if (tile_screened(s, tl)) continue;
    const BlockPair b1 = s.bp[tl.bp1];
    bool valid;
    const double W = weight(tl, b1, tid, &valid);
    if (W != 0.0) items[atomicAdd(&s_count, 1)] = (unsigned short)((ti << 8) | tid);
    else if (valid) ++nzero;
  }
  __syncthreads();
  const int nitems = s_count;
  // phase 2: list entries round robin
  for (int k0 = 0; k0 < KKk; k0 += KC) {
    const int kn = KKk - k0 < KC ? KKk - k0 : KC;
    for (int x = 0; x < kn * kAdjFields; ++x) sacc[x * 256 + tid] = 0.0;
    for (int it = tid; it < nitems; it += 256) {
      const int item = items[it], ti = item >> 8, t = item & 255;
      const int bl = t >> 4, kt = t & 15;
      const Tile tl = tiles[ch.x + ti];
      const BlockPair b1 = s.bp[tl.bp1];
      bool valid;
      const double W = weight(tl, b1, t, &valid);
      const int KKb = b1.KK;
      const int ia = b1.axI, ib = b1.axJ;
      const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
      const long long off5b = b1.off / kPairFields * kAdjFields;
      const double* gbra = pd + b1.off * 16;
      double pas = 0.0, pbs = 0.0, qcs = 0.0, qds = 0.0;
      for (int pb = 0; pb < KKb; ++pb) {
        const double* brec = gbra + pb * kPairFields * 16;
        PairAdj ba; ba.P[0] = ba.P[1] = ba.P[2] = ba.g = ba.K = 0.0;
        double bsl[4] = {0.0, 0.0, 0.0, 0.0};
        const PairPrim bra = load_pair(brec, bl);
        if (s.zero_skip && pair_underflowed(bra.K)) continue;
        double bs[4] = {0.0, 0.0, 0.0, 0.0};
        if (LA) bs[0] = brec[ia * 16 + bl];
        if (LB) bs[1] = brec[ib * 16 + bl];
        if (LC) bs[2] = brec[ic * 16 + bl];
        if (LD) bs[3] = brec[id * 16 + bl];
        for (int pk = 0; pk < kn; ++pk) {
          const double* krec = sket + (k0 + pk) * kPairFields * 16;
          const PairPrim ket = load_pair(krec, kt);
          if (s.zero_skip && pair_underflowed(ket.K)) continue;
          double ds[4] = {0.0, 0.0, 0.0, 0.0};
          if (LA) ds[0] = bs[0] - krec[ia * 16 + kt];
          if (LB) ds[1] = bs[1] - krec[ib * 16 + kt];
          if (LC) ds[2] = bs[2] - krec[ic * 16 + kt];
          if (LD) ds[3] = bs[3] - krec[id * 16 + kt];
          QuartetAdj q;
          prim_quartet<LA, LB, LC, LD, true>(bra, ket, ds, dl, c_boys, W, &q);
          ba.P[0] += q.dP[0]; ba.P[1] += q.dP[1]; ba.P[2] += q.dP[2]; ba.g += q.gb; ba.K += q.Kb;
          if (LA) { bsl[0] += q.ds[0] + q.e[0]; pas += q.e[0]; }
          if (LB) { bsl[1] += q.ds[1] + q.e[1]; pbs += q.e[1]; }
          if (LC) { bsl[2] += q.ds[2]; qcs += q.e[2]; }
          if (LD) { bsl[3] += q.ds[3]; qds += q.e[3]; }
          double* a = sacc + (pk * kAdjFields) * 256 + tid;
          a[0] -= q.dP[0]; a[256] -= q.dP[1]; a[512] -= q.dP[2]; a[768] += q.gk; a[1024] += q.Kk;
          if (LA) a[ia * 256] -= q.ds[0];
          if (LB) a[ib * 256] -= q.ds[1];
          if (LC) a[ic * 256] += q.e[2] - q.ds[2];
          if (LD) a[id * 256] += q.e[3] - q.ds[3];
        }
        if (LA) add3(ba.P, ia, bsl[0]);
        if (LB) add3(ba.P, ib, bsl[1]);
        if (LC) add3(ba.P, ic, bsl[2]);
        if (LD) add3(ba.P, id, bsl[3]);
        const long long o = (off5b + (long long)pb * kAdjFields) * 16 + bl;
        if (ba.P[0] != 0.0) acc_add(padj, o, ba.P[0], s.det);
        if (ba.P[1] != 0.0) acc_add(padj, o + 16, ba.P[1], s.det);
        if (ba.P[2] != 0.0) acc_add(padj, o + 32, ba.P[2], s.det);
        if (ba.g != 0.0) acc_add(padj, o + 48, ba.g, s.det);
        if (ba.K != 0.0) acc_add(padj, o + 64, ba.K, s.det);
      }
      // this quartet's ket adjoints: flush the private slots and clear them for the next list entry
      for (int x = 0; x < kn * kAdjFields; ++x) {
        const double v = sacc[x * 256 + tid];
        if (v != 0.0) { acc_add(padj, (off5k + (long long)k0 * kAdjFields + x) * 16 + kt, v, s.det); sacc[x * 256 + tid] = 0.0; }
      }
      if (LA && pas != 0.0) acc_add(pasum, ((long long)tl.bp1 * 2 + 0) * 16 + bl, pas, s.det);
      if (LB && pbs != 0.0) acc_add(pasum, ((long long)tl.bp1 * 2 + 1) * 16 + bl, pbs, s.det);
      if (LC && qcs != 0.0) acc_add(pasum, ((long long)bp2 * 2 + 0) * 16 + kt, qcs, s.det);
      if (LD && qds != 0.0) acc_add(pasum, ((long long)bp2 * 2 + 1) * 16 + kt, qds, s.det);
    }
  }
  if (s.zero_weight) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nzero += __shfl_xor_sync(0xffffffffu, nzero, o);
    if ((tid & 31) == 0 && nzero) atomicAdd(s.zero_weight, (unsigned long long)nzero);
  }
}

// ---------------------------------------------------------------------------------------------
// ERI gradient, second kernel of the DEFERRED form: the loop-regime primitive quartets k_eri_grad_chunks<..., DEFER> marked.
// One CTA per tile; CTAs of tiles without marks leave after reading their 3 KB of masks.  Otherwise, as in k_eri_tiles_def:
// the marked quartets are counted per bin of the Boys argument, listed bin by bin, and taken round robin by the 256
// threads - whoever owns them - with the loop code of the Boys function; what differs is the delivery: a gradient quartet
// has 14 sums to deliver (5 + 2 adjoints of its bra primitive pair, 5 + 2 of its ket primitive pair) instead of one value.
// They go to fixed-point accumulators in shared memory (one per primitive-pair field of the tile's two block pairs: the
// order of the atomics does not matter), which are added limb by limb to the global accumulators at the end.
// ---------------------------------------------------------------------------------------------
template <bool LA, bool LB, bool LC, bool LD>
__global__ void __launch_bounds__(256, 2)
k_eri_grad_deferred(const Tile* __restrict__ tiles, int tile0, SystemDev s, const double* __restrict__ pd, int nterms,
                    const double* __restrict__ At, const double* __restrict__ Bt, double* __restrict__ padj,
                    double* __restrict__ pasum, const unsigned* __restrict__ dmask, const unsigned char* __restrict__ dbin) {
  constexpr int L = (int)LA + (int)LB + (int)LC + (int)LD;
  extern __shared__ double sm[];
  __shared__ int s_cnt[32], s_cur[32], s_total, s_any;
  const int tid = threadIdx.x;
  const long long tix = (long long)tile0 + blockIdx.x;
  const unsigned* mp = dmask + ((long long)blockIdx.x * kDeferWords) * 256 + tid;      // (dmask / dbin: class-relative tile index)
  const unsigned m0 = mp[0], m1 = mp[256], m2 = mp[512];
  if (tid == 0) s_any = 0;
  if (tid < 32) s_cnt[tid] = 0;
  __syncthreads();
  if ((m0 | m1 | m2) != 0u) s_any = 1;
  __syncthreads();
  if (!s_any) return;
  const Tile tl = tiles[tix];
  const TileCtx c = tile_ctx(tl, s);
  const int KKb = c.b1.KK, KKk = c.b2.KK;
  const int nfb = KKb * kAdjFields + 2, nfk = KKk * kAdjFields + 2;      // accumulator fields per bra / ket lane
  double* sbra = sm;
  double* sket = sbra + KKb * kPairFields * 16;
  double* sW = sket + KKk * kPairFields * 16;                             // [256] weights of the tile's quartets
  unsigned long long* accb = (unsigned long long*)(sW + 256);             // [16][nfb][2]
  unsigned long long* acck = accb + 16 * nfb * 2;                         // [16][nfk][2]
  unsigned short* items = (unsigned short*)(acck + 16 * nfk * 2);         // [256 * steps]
  stage_pairs(c, pd, sbra, sket);
  for (int x = tid; x < 16 * (nfb + nfk) * 2; x += 256) accb[x] = 0ull;
  // the weight of this thread's quartet, exactly as in k_eri_grad_chunks
  {
    const int n = s.N;
    const size_t nn = (size_t)n * n;
    double W = 0.0;
    if (c.valid && (m0 | m1 | m2) != 0u) {      // only the owners of marked quartets need their weight (108 scattered loads)
      const int i = c.b1.sI + c.il, j = c.b1.sJ + c.jl, k = c.b2.sI + c.kl, l = c.b2.sJ + c.ll;
      for (int t = 0; t < nterms; ++t) {
        const double* A = At + t * nn;
        const double* B = Bt + t * nn;
        W += 8.0 * (A[i * n + j] * B[k * n + l] + A[k * n + l] * B[i * n + j]) -
             2.0 * (A[i * n + k] * B[j * n + l] + A[j * n + k] * B[i * n + l] + A[i * n + l] * B[j * n + k] +
                    A[j * n + l] * B[i * n + k]);
      }
      W *= (tl.bp1 == tl.bp2 ? 0.5 : 1.0) * (c.b1.bI == c.b1.bJ ? 0.5 : 1.0) * (c.b2.bI == c.b2.bJ ? 0.5 : 1.0);
    }
    sW[tid] = W;
  }
  __syncthreads();
  const int ia = c.b1.axI, ib = c.b1.axJ, ic = c.b2.axI, id = c.b2.axJ;
  const AxisDeltas dl = quartet_deltas<LA, LB, LC, LD>(ia, ib, ic, id);
  const double tmid = boys_midrange_from<L>();
  const double binscale = 31.99 / tmid;
  auto bin_of = [&](double t) { return t > 0.0 ? min(31, (int)(t * binscale)) : 0; };
  // count the marks per argument bin; the first kernel left the bin of every marked step in dbin (one byte): listing
  // needs no second evaluation of t (ncu: recomputing it made the two listing loops 29 % of this kernel's instructions at
  // 6 lanes per instruction)
  const unsigned char* tbin = dbin + (long long)blockIdx.x * kDefMaxSteps * 256 + tid;
  auto bin_at = [&](int, int st) -> int { return tbin[st * 256] & 31; };
  (void)bin_of; (void)tmid;
  {
    int j = 0;
#pragma unroll
    for (int wd = 0; wd < 3; ++wd) {
      for (unsigned m = wd == 0 ? m0 : (wd == 1 ? m1 : m2); m; m &= m - 1, ++j) atomicAdd(&s_cnt[bin_at(j, wd * 32 + __ffs(m) - 1)], 1);
    }
  }
  __syncthreads();
  if (tid < 32) {
    const int cnt = s_cnt[tid];
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
    s_cur[tid] = incl - cnt;
    if (tid == 31) s_total = incl;
  }
  __syncthreads();
  const int total = s_total;
  {
    int j = 0;
#pragma unroll
    for (int wd = 0; wd < 3; ++wd) {
      for (unsigned m = wd == 0 ? m0 : (wd == 1 ? m1 : m2); m; m &= m - 1, ++j) {
        const int st = wd * 32 + __ffs(m) - 1;
        items[atomicAdd(&s_cur[bin_at(j, st)], 1)] = (unsigned short)((tid << 7) | st);
      }
    }
  }
  __syncthreads();
  // the listed quartets, round robin
  for (int x = tid; x < total; x += 256) {
    const int item = items[x];
    const int owner = item >> 7, st = item & 127;
    const int pb = st / KKk, pk = st - pb * KKk;
    const int bl = owner >> 4, kt = owner & 15;
    const double* brec = sbra + pb * kPairFields * 16;
    const double* krec = sket + pk * kPairFields * 16;
    const PairPrim bra = load_pair(brec, bl);
    const PairPrim ket = load_pair(krec, kt);
    double ds[4] = {0.0, 0.0, 0.0, 0.0};
    if (LA) ds[0] = brec[ia * 16 + bl] - krec[ia * 16 + kt];
    if (LB) ds[1] = brec[ib * 16 + bl] - krec[ib * 16 + kt];
    if (LC) ds[2] = brec[ic * 16 + bl] - krec[ic * 16 + kt];
    if (LD) ds[3] = brec[id * 16 + bl] - krec[id * 16 + kt];
    QuartetAdj q;
    prim_quartet<LA, LB, LC, LD, true, kBoysLoopsOnly>(bra, ket, ds, dl, c_boys, sW[owner], &q);
    // bra side (k_eri_grad_chunks: ba, bsl folded into ba.P along the axes, pas / pbs)
    double bP[3] = {q.dP[0], q.dP[1], q.dP[2]}, kP[3] = {-q.dP[0], -q.dP[1], -q.dP[2]};
    if (LA) { add3(bP, ia, q.ds[0] + q.e[0]); add3(kP, ia, -q.ds[0]); }
    if (LB) { add3(bP, ib, q.ds[1] + q.e[1]); add3(kP, ib, -q.ds[1]); }
    if (LC) { add3(bP, ic, q.ds[2]); add3(kP, ic, q.e[2] - q.ds[2]); }
    if (LD) { add3(bP, id, q.ds[3]); add3(kP, id, q.e[3] - q.ds[3]); }
    unsigned long long* ab = accb + ((size_t)bl * nfb + pb * kAdjFields) * 2;
    if (bP[0] != 0.0) det_add_shared(ab + 0, bP[0]);
    if (bP[1] != 0.0) det_add_shared(ab + 2, bP[1]);
    if (bP[2] != 0.0) det_add_shared(ab + 4, bP[2]);
    if (q.gb != 0.0) det_add_shared(ab + 6, q.gb);
    if (q.Kb != 0.0) det_add_shared(ab + 8, q.Kb);
    unsigned long long* as_ = accb + ((size_t)bl * nfb + KKb * kAdjFields) * 2;
    if (LA && q.e[0] != 0.0) det_add_shared(as_ + 0, q.e[0]);
    if (LB && q.e[1] != 0.0) det_add_shared(as_ + 2, q.e[1]);
    unsigned long long* ak = acck + ((size_t)kt * nfk + pk * kAdjFields) * 2;
    if (kP[0] != 0.0) det_add_shared(ak + 0, kP[0]);
    if (kP[1] != 0.0) det_add_shared(ak + 2, kP[1]);
    if (kP[2] != 0.0) det_add_shared(ak + 4, kP[2]);
    if (q.gk != 0.0) det_add_shared(ak + 6, q.gk);
    if (q.Kk != 0.0) det_add_shared(ak + 8, q.Kk);
    unsigned long long* aks = acck + ((size_t)kt * nfk + KKk * kAdjFields) * 2;
    if (LC && q.e[2] != 0.0) det_add_shared(aks + 0, q.e[2]);
    if (LD && q.e[3] != 0.0) det_add_shared(aks + 2, q.e[3]);
  }
  __syncthreads();
  // shared-memory limbs -> global accumulators (integer limbs as they are in fixed-point mode, a double otherwise)
  const long long off5b = c.b1.off / kPairFields * kAdjFields, off5k = c.b2.off / kPairFields * kAdjFields;
  for (int x = tid; x < 16 * (nfb + nfk); x += 256) {
    const bool isb = x < 16 * nfb;
    const int y = isb ? x : x - 16 * nfb;
    const int nf = isb ? nfb : nfk, KK = isb ? KKb : KKk;
    const int lane16 = y / nf, f = y - lane16 * nf;
    const unsigned long long hi = accb[2 * x], lo = accb[2 * x + 1];
    if ((hi | lo) == 0ull) continue;
    double* base;
    long long idx;
    if (f < KK * kAdjFields) { base = padj; idx = ((isb ? off5b : off5k) + f) * 16 + lane16; }
    else { base = pasum; idx = ((long long)(isb ? tl.bp1 : tl.bp2) * 2 + (f - KK * kAdjFields)) * 16 + lane16; }
    if (s.det) {
      unsigned long long* gp = (unsigned long long*)base + 2 * idx;
      if (hi) atomicAdd(gp, hi);
      if (lo) atomicAdd(gp + 1, lo);
    } else {
      atomicAdd(base + idx, (double)(long long)hi * (1.0 / 16777216.0) + (double)(long long)lo * (1.0 / 147573952589676412928.0));
    }
  }
}

// wmask[x] = 1 iff B_t[x] != 0 for some term t
__global__ void k_weight_mask(int n2, int nterms, const double* __restrict__ Bt, unsigned char* __restrict__ mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n2) return;
  bool on = false;
  for (int t = 0; t < nterms && !on; ++t) on = Bt[(size_t)t * n2 + x] != 0.0;
  mask[x] = on ? 1 : 0;
}

// number of nonzero elements of an n x n matrix (host decision: dense or sparse gradient kernel)
__global__ void k_count_nonzero(int n2, const double* __restrict__ A, int* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int nz = __syncthreads_count(x < n2 && A[x < n2 ? x : 0] != 0.0);
  if (threadIdx.x == 0 && nz) atomicAdd(out, nz);
}

// pair-field adjoints -> gradient w.r.t. exponents, NORMALISED coefficients and centres (sorted order)
__global__ void __launch_bounds__(128) k_pair_bwd(SystemDev s, const double* __restrict__ padj, const double* __restrict__ pasum,
                                                  double* __restrict__ g_alpha, double* __restrict__ g_coef,
                                                  double* __restrict__ g_xyz) {
  const int bpi = blockIdx.x;
  const BlockPair b = s.bp[bpi];
  const int lane = threadIdx.x & 15, il = lane >> 2, jl = lane & 3;
  if (!(il < b.nI && jl < b.nJ)) return;
  const int i = b.sI + il, j = b.sJ + jl;
  const long long off5 = b.off / kPairFields * kAdjFields;
  double Ab[3] = {0, 0, 0}, Bb[3] = {0, 0, 0};
  const long long oc = g_coef - g_alpha, ox = g_xyz - g_alpha;     // windows of one accumulator buffer (see k_one_electron_grad)
  const int first = threadIdx.x >> 4;
  for (int pp = first; pp < b.KK; pp += blockDim.x >> 4) {
    const int p = s.f_off[i] + pp / b.KJ, q = s.f_off[j] + pp % b.KJ;
    const long long a = (off5 + (long long)pp * kAdjFields) * 16 + lane;
    PairAdj adj;
    if (s.det) { adj.P[0] = det_value(padj, a); adj.P[1] = det_value(padj, a + 16); adj.P[2] = det_value(padj, a + 32); adj.g = det_value(padj, a + 48); adj.K = det_value(padj, a + 64); }
    else { adj.P[0] = padj[a]; adj.P[1] = padj[a + 16]; adj.P[2] = padj[a + 32]; adj.g = padj[a + 48]; adj.K = padj[a + 64]; }
    double ab = 0, bb = 0, cab = 0, cbb = 0;
    pair_backward(s.p_alpha[p], s.p_coef[p], s.f_xyz + 3 * i, s.f_type[i], s.p_alpha[q], s.p_coef[q], s.f_xyz + 3 * j,
                  s.f_type[j], adj, 0.0, 0.0, &ab, &bb, &cab, &cbb, Ab, Bb);
    acc_add(g_alpha, p, ab, s.det); acc_add(g_alpha, q, bb, s.det);
    acc_add(g_alpha, oc + p, cab, s.det); acc_add(g_alpha, oc + q, cbb, s.det);
  }
  if (first == 0) {   // explicit centre dependence of (P-A)_a, (P-B)_b summed over the contraction
    const long long pa = ((long long)bpi * 2 + 0) * 16 + lane, pb = ((long long)bpi * 2 + 1) * 16 + lane;
    if (s.f_type[i] > 0) Ab[s.f_type[i] - 1] -= s.det ? det_value(pasum, pa) : pasum[pa];
    if (s.f_type[j] > 0) Bb[s.f_type[j] - 1] -= s.det ? det_value(pasum, pb) : pasum[pb];
  }
#pragma unroll
  for (int x = 0; x < 3; ++x) { acc_add(g_alpha, ox + 3 * i + x, Ab[x], s.det); acc_add(g_alpha, ox + 3 * j + x, Bb[x], s.det); }
}

// =============================================================================================
// dense helpers for the SCF and its reverse sweep
// =============================================================================================
// C[M,N] = alpha * op(A) op(B) + beta * C, row-major; op = transpose if TA/TB
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, const double* __restrict__ A, int lda,
                                              const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc,
                                              double alpha, double beta) {
  __shared__ double sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx, kb = k0 + ty;
    sa[ty][tx] = (row < M && ka < K) ? (TA ? A[ka * lda + row] : A[row * lda + ka]) : 0.0;
    sb[ty][tx] = (kb < K && col < N) ? (TB ? B[col * ldb + kb] : B[kb * ldb + col]) : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[k][tx];
    __syncthreads();
  }
  if (row < M && col < N) C[row * ldc + col] = alpha * acc + (beta != 0.0 ? beta * C[row * ldc + col] : 0.0);
}

// symmetric eigensolver: parallel cyclic Jacobi (round-robin pairing), one CTA per matrix.
// A (n x n, row-major, symmetric) is destroyed; eigenvalues ascending in w, eigenvectors in the columns of Vout.
__global__ void __launch_bounds__(1024) k_eigh_jacobi(int n, double* __restrict__ Aall, double* __restrict__ Vall,
                                                      double* __restrict__ Voutall, double* __restrict__ wall, int max_sweeps) {
  extern __shared__ double sm[];
  const size_t nn = (size_t)n * n;
  double* A = Aall + blockIdx.x * nn;
  double* V = Vall + blockIdx.x * nn;
  double* Vout = Voutall + blockIdx.x * nn;
  double* w = wall + (size_t)blockIdx.x * n;
  const int m = (n + 1) / 2;        // pairs per round
  const int np = 2 * m;             // players (one dummy if n is odd)
  double* cs = sm;                  // [m]
  double* sn = sm + m;              // [m]
  int* pp = (int*)(sm + 2 * m);     // [m]
  int* qq = pp + m;                 // [m]
  __shared__ double red[32];
  __shared__ double s_off, s_diag;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int x = tid; x < (int)nn; x += nt) V[x] = (x / n == x % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    // convergence: off-diagonal Frobenius norm relative to the diagonal
    double off = 0.0, dg = 0.0;
    for (int x = tid; x < (int)nn; x += nt) { const double a = A[x]; if (x / n == x % n) dg += a * a; else off += a * a; }
    for (int o = 16; o >= 1; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
    if ((tid & 31) == 0) { red[tid >> 5] = off; }
    __syncthreads();
    if (tid == 0) { double t = 0; for (int x = 0; x < (nt + 31) / 32; ++x) t += red[x]; s_off = t; }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = dg; }
    __syncthreads();
    if (tid == 0) { double t = 0; for (int x = 0; x < (nt + 31) / 32; ++x) t += red[x]; s_diag = t; }
    __syncthreads();
    if (s_off <= 1e-30 * (s_diag + s_off)) break;
    for (int round = 0; round < np - 1; ++round) {
      if (tid < m) {
        int a = (tid == 0) ? np - 1 : (round + tid) % (np - 1);
        int b = (tid == 0) ? round : (round + np - 1 - tid) % (np - 1);
        int p = a < b ? a : b, q = a < b ? b : a;
        double c = 1.0, s_ = 0.0;
        if (q < n) {
          const double apq = A[p * n + q];
          if (fabs(apq) > 1e-300) {
            const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = rsqrt(t * t + 1.0); s_ = t * c;
          }
        } else { q = -1; }
        cs[tid] = c; sn[tid] = s_; pp[tid] = p; qq[tid] = q;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J   (row r, pair k)
      for (int x = tid; x < m * n; x += nt) {
        const int r = x / m, k = x % m;
        const int q = qq[k];
        if (q < 0 || sn[k] == 0.0) continue;
        const int p = pp[k]; const double c = cs[k], s_ = sn[k];
        const double arp = A[r * n + p], arq = A[r * n + q];
        A[r * n + p] = c * arp - s_ * arq; A[r * n + q] = s_ * arp + c * arq;
        const double vrp = V[r * n + p], vrq = V[r * n + q];
        V[r * n + p] = c * vrp - s_ * vrq; V[r * n + q] = s_ * vrp + c * vrq;
      }
      __syncthreads();
      // rows: A <- J^T A   (pair k, column cc)
      for (int x = tid; x < m * n; x += nt) {
        const int k = x / n, cc = x % n;
        const int q = qq[k];
        if (q < 0 || sn[k] == 0.0) continue;
        const int p = pp[k]; const double c = cs[k], s_ = sn[k];
        const double apc = A[p * n + cc], aqc = A[q * n + cc];
        A[p * n + cc] = c * apc - s_ * aqc; A[q * n + cc] = s_ * apc + c * aqc;
      }
      __syncthreads();
    }
  }
  // ascending order by rank
  for (int i = tid; i < n; i += nt) {
    const double wi = A[i * n + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) { const double wj = A[j * n + j]; rank += (wj < wi) || (wj == wi && j < i); }
    w[rank] = wi;
    for (int r = 0; r < n; ++r) Vout[r * n + rank] = V[r * n + i];
  }
}

// ---------------------------------------------------------------------------------------------
// fast path of the symmetric eigensolver (n <= kJacobiPackedMax): two kernels.
//  k_jacobi_packed : parallel cyclic Jacobi on A only, upper triangle PACKED in shared memory (n = 224 -> 202 KB),
//                    one CTA per matrix.  Per round the m = n/2 disjoint rotations are applied as 2x2 block updates
//                    B <- J_k^T B J_l for every pair of pairs (k < l): no global traffic inside the sweep.
//                    The (c, s) of every rotation are appended to a log in global memory.
//  k_jacobi_vectors: replays the log on the identity, one WARP per eigenvector-matrix row held in shared memory
//                    (rows are independent, the m rotations of a round are disjoint), and writes the columns in
//                    ascending eigenvalue order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri_packed(int i, int j, int n) {   // i <= j
  return i * n - ((i * (i - 1)) >> 1) + (j - i);
}
__device__ __forceinline__ int sym_packed(int i, int j, int n) { return i <= j ? tri_packed(i, j, n) : tri_packed(j, i, n); }
__device__ __forceinline__ void rr_pair(int round, int k, int np, int* p, int* q) {   // round-robin tournament
  const int a = (k == 0) ? np - 1 : (round + k) % (np - 1);
  const int b = (k == 0) ? round : (round + np - 1 - k) % (np - 1);
  *p = a < b ? a : b; *q = a < b ? b : a;
}

__global__ void __launch_bounds__(1024) k_jacobi_packed(int n, const double* __restrict__ Aall, unsigned short* pairtab,
                                                        double2* __restrict__ logall, int* __restrict__ metaall,
                                                        double* __restrict__ wall, int* __restrict__ rankall, int max_sweeps) {
  extern __shared__ double sm[];
  const int ne = n + (n & 1), m = ne >> 1, ntri = ne * (ne + 1) / 2;
  const double* A = Aall + (size_t)blockIdx.x * n * n;
  double2* log = logall + (size_t)blockIdx.x * max_sweeps * (ne - 1) * m;
  double* Ap = sm;                 // [ntri]
  double* cs = sm + ntri;          // [m]
  double* sn = cs + m;             // [m]
  int* pp = (int*)(sn + m);        // [m]
  int* qq = pp + m;                // [m]
  __shared__ double red[32];
  __shared__ double s_off, s_tot;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int x = tid; x < ne * ne; x += nt) {
    const int i = x / ne, j = x - i * ne;
    if (i <= j) Ap[tri_packed(i, j, ne)] = (i < n && j < n) ? A[i * n + j] : 0.0;
  }
  const int ntask = m * (m - 1) / 2;
  double prev_off = 1e300;
  // table of the pairs of pairs (k < l), task x = l(l-1)/2 + k (every CTA writes the same values)
  for (int x = tid; x < ntask; x += nt) {
    int l = (int)((1.0 + sqrt(1.0 + 8.0 * x)) * 0.5);
    while (l * (l - 1) / 2 > x) --l;
    while ((l + 1) * l / 2 <= x) ++l;
    pairtab[x] = (unsigned short)(((x - l * (l - 1) / 2) << 8) | l);
  }
  __syncthreads();
  int rounds = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    // off-diagonal weight
    double off = 0.0, tot = 0.0;
    for (int x = tid; x < ne * ne; x += nt) {
      const int i = x / ne, j = x - i * ne;
      if (i <= j) { const double a = Ap[tri_packed(i, j, ne)]; if (i == j) tot += a * a; else off += 2.0 * a * a; }
    }
    for (int o = 16; o >= 1; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    if ((tid & 31) == 0) red[tid >> 5] = off;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int x = 0; x < (nt + 31) / 32; ++x) t += red[x]; s_off = t; }
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = tot;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int x = 0; x < (nt + 31) / 32; ++x) t += red[x]; s_tot = t + s_off; }
    __syncthreads();
    // converged, or stagnating at the rounding floor (off-norm already ~1e-12 relative and no longer shrinking)
    if (s_off <= 1e-30 * s_tot || (s_off <= 1e-24 * s_tot && s_off > 0.25 * prev_off)) break;
    prev_off = s_off;
    for (int round = 0; round < ne - 1; ++round, ++rounds) {
      if (tid < m) {
        int p, q; rr_pair(round, tid, ne, &p, &q);
        const int ipp = tri_packed(p, p, ne), iqq = tri_packed(q, q, ne), ipq = tri_packed(p, q, ne);
        const double apq = Ap[ipq];
        double c = 1.0, s_ = 0.0;
        const double app = Ap[ipp], aqq = Ap[iqq];
        // a rotation by less than ~1e-18 rad changes nothing representable: drop the element instead (Numerical
        // Recipes' jacobi does the same); lets the last sweep of a warm-started solve skip its block updates
        if (fabs(apq) <= 1e-18 * (fabs(app) + fabs(aqq))) {
          Ap[ipq] = 0.0;
        } else {
          const double theta = (aqq - app) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          c = rsqrt(t * t + 1.0); s_ = t * c;
          Ap[ipp] = app - t * apq; Ap[iqq] = aqq + t * apq; Ap[ipq] = 0.0;
        }
        cs[tid] = c; sn[tid] = s_; pp[tid] = p; qq[tid] = q;
        log[(size_t)rounds * m + tid] = make_double2(c, s_);
      }
      __syncthreads();
      for (int x = tid; x < ntask; x += nt) {
        const unsigned short kl = pairtab[x];
        const int k = kl >> 8, l = kl & 255;
        const double ck = cs[k], sk = sn[k], cl = cs[l], sl = sn[l];
        if (sk == 0.0 && sl == 0.0) continue;
        const int p = pp[k], q = qq[k], r = pp[l], u = qq[l];
        const int ipr = sym_packed(p, r, ne), ipu = sym_packed(p, u, ne), iqr = sym_packed(q, r, ne), iqu = sym_packed(q, u, ne);
        const double bpr = Ap[ipr], bpu = Ap[ipu], bqr = Ap[iqr], bqu = Ap[iqu];
        // T = B J_l
        const double tpr = cl * bpr - sl * bpu, tpu = sl * bpr + cl * bpu;
        const double tqr = cl * bqr - sl * bqu, tqu = sl * bqr + cl * bqu;
        // B' = J_k^T T
        Ap[ipr] = ck * tpr - sk * tqr; Ap[ipu] = ck * tpu - sk * tqu;
        Ap[iqr] = sk * tpr + ck * tqr; Ap[iqu] = sk * tpu + ck * tqu;
      }
      __syncthreads();
    }
  }
  if (tid == 0) metaall[blockIdx.x] = rounds;
  // ascending order by rank (the padding index of an odd n is left out)
  for (int i = tid; i < n; i += nt) {
    const double wi = Ap[tri_packed(i, i, ne)];
    int rank = 0;
    for (int j = 0; j < n; ++j) { const double wj = Ap[tri_packed(j, j, ne)]; rank += (wj < wi) || (wj == wi && j < i); }
    wall[(size_t)blockIdx.x * n + rank] = wi;
    rankall[(size_t)blockIdx.x * n + i] = rank;
  }
}

__global__ void __launch_bounds__(256) k_jacobi_vectors(int n, const double2* __restrict__ logall, const int* __restrict__ metaall,
                                                        const int* __restrict__ rankall, double* __restrict__ Voutall, int max_sweeps) {
  extern __shared__ double sm[];
  const int ne = n + (n & 1), m = ne >> 1, np1 = ne - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  const int b = blockIdx.y;
  if (row >= n) return;
  const double2* log = logall + (size_t)b * max_sweeps * np1 * m;
  const int rounds = metaall[b];
  double* v = sm + (size_t)warp * ne;
  for (int x = lane; x < ne; x += 32) v[x] = (x == row) ? 1.0 : 0.0;
  __syncwarp();
  // each lane owns pairs k = lane, lane+32, ... (at most 4 for n <= 256); the tournament positions advance by one per round
  constexpr int KP = 4;
  int pa[KP], pb[KP];
  double2 cur[KP], nxt[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    const int k = lane + 32 * j;
    pa[j] = k % np1; pb[j] = (np1 - (k % np1)) % np1;     // round 0: a = k, b = -k (mod np1); k == 0 is the fixed player
    cur[j] = (k < m && rounds > 0) ? log[k] : make_double2(1.0, 0.0);
  }
  for (int rd = 0; rd < rounds; ++rd) {
    if (rd + 1 < rounds) {
#pragma unroll
      for (int j = 0; j < KP; ++j) { const int k = lane + 32 * j; nxt[j] = k < m ? log[(size_t)(rd + 1) * m + k] : make_double2(1.0, 0.0); }
    }
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const int k = lane + 32 * j;
      if (k < m && cur[j].y != 0.0) {
        const int a = (k == 0) ? ne - 1 : pa[j], bb = pb[j];
        const double va = v[a], vb = v[bb];
        // rotation acts on (p, q) = (min, max); swapping the roles flips the sign of s
        const double s_ = a < bb ? cur[j].y : -cur[j].y;
        v[a] = cur[j].x * va - s_ * vb; v[bb] = s_ * va + cur[j].x * vb;
      }
      pa[j] = pa[j] + 1 == np1 ? 0 : pa[j] + 1;
      pb[j] = pb[j] + 1 == np1 ? 0 : pb[j] + 1;
      cur[j] = nxt[j];
    }
    __syncwarp();
  }
  double* Vout = Voutall + (size_t)b * n * n + (size_t)row * n;
  const int* rank = rankall + (size_t)b * n;
  for (int x = lane; x < n; x += 32) Vout[rank[x]] = v[x];
}

// ---------------------------------------------------------------------------------------------
// Symmetric eigensolver on a thread-block CLUSTER (48 <= n <= kJacobiPackedMax): one-sided (Hestenes) Jacobi.
//   G = (A + sigma I) V,  V orthogonal (identity, or the previous SCF iteration's eigenvectors as a warm start);
//   plane rotations of column pairs, applied to G and V alike, make the columns of G mutually orthogonal; then
//   V^T (A + sigma I)^2 V is diagonal, i.e. V holds the eigenvectors (sigma = 1.5 max_i sum_j |a_ij| makes A + sigma I
//   positive definite with condition number <= 5) and lambda_i = g_i . v_i - sigma.
// A rotation touches only its two columns, so the columns are DISTRIBUTED: 16 blocks of w = ceil(n/16) columns, two
// blocks (2 w columns of G and of V, column-major) in the shared memory of each of the 8 CTAs of the cluster.  Per round
// a CTA orthogonalises every column of its top block against every column of its bottom block (w steps of w disjoint
// pairs, one warp per pair: 3 dot products, 1 rotation), then the blocks move one position round the ring (round-robin
// tournament over the 16 blocks, block 0 fixed) by writing them into the NEXT buffer of the neighbour CTA through
// distributed shared memory; 15 rounds = every pair of blocks once = one sweep (pairs inside a block in round 0).
// Against the single-CTA two-sided kernel above: the same ~223 steps per sweep, each ~8x shorter (8 SMs, no packed-index
// arithmetic), and a warm start needs no rotation of A into the old eigenbasis and no back-multiplication.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kOsCluster = 8;
constexpr int kOsThreads = 512;
constexpr int kOsMinN = 48;

// one warp: orthogonalise columns (gp, gq) and carry (vp, vq) along.  Every pair with |cos| > 1e-16 is rotated; the return
// value says whether the cosine BEFORE the rotation exceeded 1e-10.  A sweep in which no pair did leaves residual cosines
// of order cos^2 / (relative gap) <= 1e-20 / gap: the solver stops after it without a verification sweep.
// The rotation is orthogonal to the last bit by construction (cs = (1+t^2)^-1/2, sn = cs t) whatever the accuracy of t,
// which only sets the convergence rate: t is evaluated in single precision on operands scaled by a power of two.
// |g_p|^2 and |g_q|^2 are not recomputed per pair: every column carries its squared norm (na, nb: shared memory, recomputed
// exactly once per sweep, updated by the exact rotation identities in between), so a pair costs ONE dot product and ONE
// shuffle reduction.  The norms only enter the skip test and the (single-precision) rotation angle.
__device__ __forceinline__ int os_rotate(double* __restrict__ gp, double* __restrict__ gq, double* __restrict__ vp,
                                         double* __restrict__ vq, double* na, double* nb, int n, int lane) {
  double x[8], y[8];                              // n <= 256: at most 8 rows per lane, kept for the update
  double g = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = lane + 32 * i;
    x[i] = r < n ? gp[r] : 0.0; y[i] = r < n ? gq[r] : 0.0;
    g = fma(x[i], y[i], g);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  const double a = *na, b = *nb;
  const double ab = a * b, g2 = g * g;
  if (ab == 0.0 || g2 <= 1e-32 * ab) return 0;   // a padding column, or orthogonal to working precision
  // tan(theta) = 2g / (d + sign(d) sqrt(d^2 + 4 g^2)), d = b - a; operands scaled by the power of two of the larger one
  const double d = b - a, g2x = 2.0 * g;
  const int e = (__double2hiint(fmax(fabs(d), fabs(g2x))) >> 20) & 0x7ff;
  const double scale = __hiloint2double((2046 - e) << 20, 0);     // 2^-(e-1023)
  const float df = (float)(d * scale), gf = (float)(g2x * scale);
  const float hf = sqrtf(df * df + gf * gf);
  const double t = (double)__fdividef(gf, df + copysignf(hf, df));
  const double cs = rsqrt_(t * t + 1.0), sn = cs * t;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = lane + 32 * i;
    if (r < n) {
      gp[r] = cs * x[i] - sn * y[i]; gq[r] = sn * x[i] + cs * y[i];
      const double u = vp[r], v = vq[r];
      vp[r] = cs * u - sn * v; vq[r] = sn * u + cs * v;
    }
  }
  __syncwarp();
  if (lane == 0) {       // |c gp - s gq|^2 and |s gp + c gq|^2
    const double cc = cs * cs, ss = sn * sn, x2 = 2.0 * cs * sn * g;
    *na = cc * a - x2 + ss * b;
    *nb = ss * a + x2 + cc * b;
  }
  return g2 > 1e-20 * ab;
}

// the same with the TOP column (G and V) resident in the caller's registers for a whole round: only the bottom column moves
// through shared memory
__device__ __forceinline__ int os_rotate_top(double (&x)[8], double (&u)[8], double* __restrict__ gq, double* __restrict__ vq,
                                             double* na, double* nb, int n, int lane) {
  double y[8];
  double g = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = lane + 32 * i;
    y[i] = r < n ? gq[r] : 0.0;
    g = fma(x[i], y[i], g);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  const double a = *na, b = *nb;
  const double ab = a * b, g2 = g * g;
  if (ab == 0.0 || g2 <= 1e-32 * ab) return 0;
  const double d = b - a, g2x = 2.0 * g;
  const int e = (__double2hiint(fmax(fabs(d), fabs(g2x))) >> 20) & 0x7ff;
  const double scale = __hiloint2double((2046 - e) << 20, 0);
  const float df = (float)(d * scale), gf = (float)(g2x * scale);
  const float hf = sqrtf(df * df + gf * gf);
  const double t = (double)__fdividef(gf, df + copysignf(hf, df));
  const double cs = rsqrt_(t * t + 1.0), sn = cs * t;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = lane + 32 * i;
    if (r < n) {
      const double xn = cs * x[i] - sn * y[i];
      gq[r] = sn * x[i] + cs * y[i];
      x[i] = xn;
      const double v = vq[r];
      const double un = cs * u[i] - sn * v;
      vq[r] = sn * u[i] + cs * v;
      u[i] = un;
    }
  }
  __syncwarp();
  if (lane == 0) {
    const double cc = cs * cs, ss = sn * sn, x2 = 2.0 * cs * sn * g;
    *na = cc * a - x2 + ss * b;
    *nb = ss * a + x2 + cc * b;
  }
  return g2 > 1e-20 * ab;
}

// exact squared norm of one column (one warp)
__device__ __forceinline__ double os_norm2(const double* __restrict__ gcol, int n, int lane) {
  double a = 0.0;
  for (int r = lane; r < n; r += 32) a = fma(gcol[r], gcol[r], a);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

__global__ void __cluster_dims__(kOsCluster, 1, 1) __launch_bounds__(kOsThreads, 1)
k_eigh_onesided(int n, int w, const double* __restrict__ Aall, const double* __restrict__ V0all, double* __restrict__ Voutall,
                double* __restrict__ wall, double* __restrict__ lamall, int max_sweeps) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int mat = blockIdx.x / kOsCluster;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t nn = (size_t)n * n;
  const double* A = Aall + mat * nn;
  const double* V0 = V0all ? V0all + mat * nn : nullptr;
  const int blk = w * n;                       // doubles per block of G (or of V)
  // two buffers of {G top, G bottom, V top, V bottom}
  auto bufp = [&](int which) { return sm + (size_t)which * 4 * blk; };
  __shared__ double s_red[kOsThreads / 32];
  __shared__ double s_sigma;
  __shared__ int s_rot;
  __shared__ int s_cnt[2][kOsCluster];
  __shared__ double s_nrm[2][2][16];      // [buffer][top / bottom][column]: squared norms of the resident columns of G

  // sigma = 1.5 * max_i sum_j |a_ij|  (every CTA computes the same value)
  {
    double mx = 0.0;
    for (int r = tid; r < n; r += kOsThreads) {
      double sum = 0.0;
      for (int j = 0; j < n; ++j) sum += fabs(A[(size_t)j * n + r]);     // symmetric: column r read row-wise, coalesced
      mx = fmax(mx, sum);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
      double m2 = 0.0;
      for (int x = 0; x < kOsThreads / 32; ++x) m2 = fmax(m2, s_red[x]);
      s_sigma = m2 > 0.0 ? 1.5 * m2 : 1.0;
      s_rot = 0;
    }
    __syncthreads();
  }
  const double sigma = s_sigma;
  // initial blocks: top = block c, bottom = block c + 8; column j of block b is original column b*w + j (zero beyond n)
  {
    double* G = bufp(0);
    double* V = bufp(0) + 2 * blk;
    for (int x = tid; x < 2 * blk; x += kOsThreads) {
      const int slot = x / blk, rem = x - slot * blk, j = rem / n, r = rem - j * n;
      const int col = (slot == 0 ? c : c + kOsCluster) * w + j;
      V[x] = col < n ? (V0 ? V0[(size_t)r * n + col] : (r == col ? 1.0 : 0.0)) : 0.0;
    }
    __syncthreads();
    if (!V0) {
      for (int x = tid; x < 2 * blk; x += kOsThreads) {
        const int slot = x / blk, rem = x - slot * blk, j = rem / n, r = rem - j * n;
        const int col = (slot == 0 ? c : c + kOsCluster) * w + j;
        G[x] = col < n ? A[(size_t)col * n + r] + (r == col ? sigma : 0.0) : 0.0;
      }
    } else {
      // G[:, j] = A V[:, j] + sigma V[:, j]: thread = (row r, half of the 2w columns), A read through its symmetry
      const int half = tid / n, r = tid - half * n;            // needs 2 n <= 512 threads: n <= 256
      if (half < 2) {
        const int j0 = half * w;                               // columns j0 .. j0 + w - 1 of the 2w local columns
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0;
        for (int k = 0; k < n; ++k) {
          const double a = A[(size_t)k * n + r];
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j < w) acc[j] = fma(a, V[(j0 + j) * n + k], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) if (j < w) G[(j0 + j) * n + r] = acc[j] + sigma * V[(j0 + j) * n + r];
      }
    }
    __syncthreads();
  }
  cluster.sync();

  int cur = 0;
  const int m_in = (w + 1) / 2, np_in = 2 * m_in;        // round robin inside a block
  // destinations of this CTA's blocks in the ring (block 0 = top of CTA 0 stays)
  const int top_cta = c == 0 ? 0 : (c < kOsCluster - 1 ? c + 1 : kOsCluster - 1);
  const int top_slot = (c == kOsCluster - 1) ? 1 : 0;
  const int bot_cta = c == 0 ? 1 : c - 1;
  const int bot_slot = c == 0 ? 0 : 1;
  const int sweep_cap = max_sweeps < 0 ? -max_sweeps : max_sweeps;
  for (int sweep = 0; sweep < sweep_cap; ++sweep) {
    int rot = 0;
    for (int rnd = 0; rnd < 2 * kOsCluster - 1; ++rnd) {
      double* Gt = bufp(cur);
      double* Gb = Gt + blk;
      double* Vt = Gt + 2 * blk;
      double* Vb = Gt + 3 * blk;
      double* nt_ = s_nrm[cur][0];
      double* nb_ = s_nrm[cur][1];
      if (rnd == 0) {           // once per sweep: exact norms, then the pairs inside each of the two resident blocks
        for (int j = warp; j < 2 * w; j += kOsThreads / 32) {
          const double v = os_norm2((j < w ? Gt : Gb) + (j < w ? j : j - w) * n, n, lane);
          if (lane == 0) s_nrm[cur][j < w ? 0 : 1][j < w ? j : j - w] = v;
        }
        __syncthreads();
        for (int st = 0; st < np_in - 1; ++st) {
          if (warp < 2 * m_in) {
            const int which = warp / m_in, k = warp - which * m_in;
            int p, q; rr_pair(st, k, np_in, &p, &q);
            if (q < w) {
              double* G = which ? Gb : Gt;
              double* V = which ? Vb : Vt;
              double* nr = which ? nb_ : nt_;
              rot += os_rotate(G + p * n, G + q * n, V + p * n, V + q * n, nr + p, nr + q, n, lane);
            }
          }
          __syncthreads();
        }
      }
      {                                     // every top column against every bottom column; warp i keeps top column i in registers
        double xt[8], ut[8];
        if (warp < w) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = lane + 32 * i;
            xt[i] = r < n ? Gt[warp * n + r] : 0.0; ut[i] = r < n ? Vt[warp * n + r] : 0.0;
          }
        }
        for (int st = 0; st < w; ++st) {
          if (warp < w) {
            int q = warp + st; if (q >= w) q -= w;
            rot += os_rotate_top(xt, ut, Gb + q * n, Vb + q * n, nt_ + warp, nb_ + q, n, lane);
          }
          __syncthreads();
        }
        if (warp < w) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = lane + 32 * i;
            if (r < n) { Gt[warp * n + r] = xt[i]; Vt[warp * n + r] = ut[i]; }
          }
        }
        __syncthreads();
      }
      // move the blocks one position round the ring: into the other buffer of the destination CTA
      {
        double* dst_t = cluster.map_shared_rank(bufp(cur ^ 1), top_cta);
        double* dst_b = cluster.map_shared_rank(bufp(cur ^ 1), bot_cta);
        if ((blk & 1) == 0) {        // 16-byte remote stores (w * n even)
          double2* d0 = (double2*)(dst_t + top_slot * blk); double2* d1 = (double2*)(dst_t + (2 + top_slot) * blk);
          double2* d2 = (double2*)(dst_b + bot_slot * blk); double2* d3 = (double2*)(dst_b + (2 + bot_slot) * blk);
          const double2 *s0 = (const double2*)Gt, *s1 = (const double2*)Vt, *s2 = (const double2*)Gb, *s3 = (const double2*)Vb;
          for (int x = tid; x < blk / 2; x += kOsThreads) { d0[x] = s0[x]; d1[x] = s1[x]; d2[x] = s2[x]; d3[x] = s3[x]; }
        } else {
          for (int x = tid; x < blk; x += kOsThreads) {
            dst_t[top_slot * blk + x] = Gt[x]; dst_t[(2 + top_slot) * blk + x] = Vt[x];
            dst_b[bot_slot * blk + x] = Gb[x]; dst_b[(2 + bot_slot) * blk + x] = Vb[x];
          }
        }
      }
      if (tid < w) {            // the norms travel with their columns
        cluster.map_shared_rank(&s_nrm[cur ^ 1][top_slot][0], top_cta)[tid] = nt_[tid];
        cluster.map_shared_rank(&s_nrm[cur ^ 1][bot_slot][0], bot_cta)[tid] = nb_[tid];
      }
      if (rnd == 2 * kOsCluster - 2) {      // end of the sweep: publish this CTA's count of significant rotations
        if (lane == 0 && rot) atomicAdd(&s_rot, rot);
        __syncthreads();
        if (tid < kOsCluster) cluster.map_shared_rank(&s_cnt[sweep & 1][0], tid)[c] = s_rot;
      }
      cluster.sync();
      cur ^= 1;
    }
    int total = 0;
#pragma unroll
    for (int x = 0; x < kOsCluster; ++x) total += s_cnt[sweep & 1][x];
    if (tid == 0) s_rot = 0;
    __syncthreads();
    if (max_sweeps < 0 && c == 0 && tid == 0)
      printf("[dq eigh] n %d sweep %d: %d pairs above 1e-10%s\n", n, sweep + 1, total, V0 ? " (warm)" : "");
    if (total == 0) break;
  }

  // eigenvalues of the resident columns, global ranks, output in ascending order
  double* G = bufp(cur);
  double* V = bufp(cur) + 2 * blk;
  double* lam = lamall + (size_t)mat * 2 * kOsCluster * w;
  for (int j = warp; j < 2 * w; j += kOsThreads / 32) {
    double gv = 0.0, gg = 0.0;
    for (int r = lane; r < n; r += 32) { const double x = G[j * n + r]; gv += x * V[j * n + r]; gg += x * x; }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { gv += __shfl_xor_sync(0xffffffffu, gv, o); gg += __shfl_xor_sync(0xffffffffu, gg, o); }
    if (lane == 0) lam[c * 2 * w + j] = gg > 0.0 ? gv - sigma : 1e300;      // padding columns sort last
  }
  __threadfence();
  cluster.sync();
  const int ntot = 2 * kOsCluster * w;
  double* Vout = Voutall + mat * nn;
  double* wout = wall + (size_t)mat * n;
  for (int j = warp; j < 2 * w; j += kOsThreads / 32) {
    const int me = c * 2 * w + j;
    const double lj = __ldcg(lam + me);
    if (lj >= 1e299) continue;
    int rank = 0;
    for (int x = lane; x < ntot; x += 32) { const double lx = __ldcg(lam + x); rank += (lx < lj) || (lx == lj && x < me); }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0) wout[rank] = lj;
    for (int r = lane; r < n; r += 32) Vout[(size_t)r * n + rank] = V[j * n + r];
  }
}

// B[i,j] = A[i,j] * f(w[j]);  mode 0: w^-1/2
__global__ void k_scale_cols_rsqrt(int n, const double* __restrict__ A, const double* __restrict__ w, double* __restrict__ B) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < n * n) B[x] = A[x] * rsqrt(w[x % n]);
}

// F = H + 2 J - K from the tile accumulators (J: upper block triangle; K = Kacc + Kacc^T)
__global__ void k_fock_combine(int n, const int* __restrict__ f_block, const double* __restrict__ H, const double* __restrict__ Jacc,
                               const double* __restrict__ Kacc, double* __restrict__ F, double* __restrict__ G) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n * n) return;
  const int i = x / n, j = x % n;
  const double jv = f_block[i] <= f_block[j] ? Jacc[i * n + j] : Jacc[j * n + i];
  const double g = 2.0 * jv - (Kacc[i * n + j] + Kacc[j * n + i]);
  if (G) G[x] = g;
  if (F) F[x] = H[x] + g;
}

// out[0] = sum_x A[x] * (B[x] + C[x])   (single CTA, fixed-order tree: deterministic)
__global__ void __launch_bounds__(1024) k_dot3(int n, const double* __restrict__ A, const double* __restrict__ B,
                                               const double* __restrict__ C, double* __restrict__ out) {
  __shared__ double red[1024];
  double acc = 0.0;
  for (int x = threadIdx.x; x < n; x += blockDim.x) acc += A[x] * (B[x] + (C ? C[x] : 0.0));
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o >= 1; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = red[0];
}

// Y[o,v] = Y[v,o] = Z[o,v] / (eps_o - eps_v) for o < ne <= v, zero elsewhere
__global__ void k_make_y(int n, int ne, const double* __restrict__ Z, const double* __restrict__ eps, double* __restrict__ Y) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n * n) return;
  const int i = x / n, j = x % n;
  const bool oi = i < ne, oj = j < ne;
  double y = 0.0;
  if (oi != oj) { const int o = oi ? i : j, v = oi ? j : i; y = 0.5 * (Z[o * n + v] + Z[v * n + o]) / (eps[o] - eps[v]); }
  Y[x] = y;
}

// M[i,j] *= Phi_ij,  Phi = divided difference of lambda^-1/2  (Daleckii-Krein; equal eigenvalues -> derivative)
__global__ void k_phi_rsqrt(int n, const double* __restrict__ w, double* __restrict__ M) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n * n) return;
  const double a = w[x / n], b = w[x % n];
  const double fa = rsqrt(a), fb = rsqrt(b);
  // (fa - fb)/(a - b) = -1/(sqrt(a) sqrt(b) (sqrt(a)+sqrt(b))): no cancellation, valid for a == b
  M[x] *= -fa * fb / (sqrt(a) + sqrt(b));
}

// Y = a*X + b*Y^T-or-Y ...: small elementwise family
__global__ void k_axpby(int n, double a, const double* __restrict__ X, double b, double* __restrict__ Y) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < n) Y[x] = a * X[x] + (b != 0.0 ? b * Y[x] : 0.0);
}
__global__ void k_symmetrize(int n, double* __restrict__ A) {   // A <- (A + A^T)/2
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n * n) return;
  const int i = x / n, j = x % n;
  if (i < j) { const double v = 0.5 * (A[i * n + j] + A[j * n + i]); A[i * n + j] = v; A[j * n + i] = v; }
}
__global__ void k_add_transpose(int n, const double* __restrict__ A, double* __restrict__ B) {   // B += A + A^T
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n * n) return;
  B[x] += A[x] + A[(x % n) * n + x / n];
}
__global__ void k_exp(int n, const double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = exp(x[i]);
}
__global__ void k_mul(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] * b[i];
}

// Boys function on its own (parity probe for Tools.py:78-151): F, dF are [n][5]
__global__ void k_boys(int n, const double* __restrict__ t, double* __restrict__ F, double* __restrict__ dF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double f[5], d[5];
  boys_ref<4, true>(t[i], c_boys, f, d);
#pragma unroll
  for (int m = 0; m < 5; ++m) { F[5 * i + m] = f[m]; dF[5 * i + m] = d[m]; }
}

// FP64 FMA throughput probe (the FP64 roofline denominator; MEASURED_PEAKS.json has none)
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double* out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000000001, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456) out[0] = r;
}

// =============================================================================================
// launchers
// =============================================================================================
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

void launch_normalize(cudaStream_t st, int N, const int* f_off, const int* f_type, const double* alpha, const double* craw, double* cn) {
  k_normalize<<<cdiv(N, 128), 128, 0, st>>>(N, f_off, f_type, alpha, craw, cn); DQ_LAUNCHED();
}
void launch_normalize_bwd(cudaStream_t st, int N, const int* f_off, const int* f_type, const double* alpha, const double* craw,
                          const double* cbar, double* abar, double* crawbar) {
  k_normalize_bwd<<<cdiv(N, 128), 128, 0, st>>>(N, f_off, f_type, alpha, craw, cbar, abar, crawbar); DQ_LAUNCHED();
}
void launch_pairs(cudaStream_t st, const SystemDev& s, double* pd, double* bp_kmax) {
  cudaMemsetAsync(bp_kmax, 0, sizeof(double) * s.nbp, st);
  k_pairs<<<s.nbp, 128, 0, st>>>(s, pd, bp_kmax); DQ_LAUNCHED();
}
void launch_one_electron(cudaStream_t st, const SystemDev& s, double* S, double* T, double* V) {
  k_one_electron<<<cdiv((long long)s.N * (s.N + 1) / 2, 64), 64, 0, st>>>(s, S, T, V); DQ_LAUNCHED();
}
void launch_nuclear_grad(cudaStream_t st, const SystemDev& s, int kmax, const double* Hbar, double* g_atoms) {
  const long long nth = (long long)s.N * (s.N + 1) / 2 * kmax * kmax;
  k_nuclear_grad<<<cdiv(nth, 128), 128, 0, st>>>(s, kmax, Hbar, g_atoms); DQ_LAUNCHED();
}
void launch_one_electron_grad(cudaStream_t st, const SystemDev& s, int kmax, const double* Sbar, const double* Hbar,
                              double* g_alpha, double* g_coef, double* g_xyz) {
  const long long nth = (long long)s.N * (s.N + 1) / 2 * kmax * kmax;
  k_one_electron_grad<<<cdiv(nth, 64), 64, 0, st>>>(s, kmax, Sbar, Hbar, g_alpha, g_coef, g_xyz); DQ_LAUNCHED();
}

template <bool A, bool B, bool C, bool D>
static void eri_launch(cudaStream_t st, const Tile* tiles, int nt, const SystemDev& s, const double* pd, double* store, size_t smem) {
  // default: k_eri_tiles_def where it applies, else the plain kernel (Boys function evaluated in place; forced with
  // DQ_ERI_KERNEL=plain).  DQ_ERI_KERNEL=batched selects the warp-batched variant for A/B runs: measured on the dense (H2O)_32 cluster it raises the lanes per instruction from 17.7 to 23.5 but executes
  // 211 instead of 132 thread instructions per primitive quartet (arguments computed twice, task lists, values through
  // shared memory): 305 ms against 282 ms for the pass.
  // (read at every launch: the tests switch the variants inside one process)
  const bool plain = !(getenv("DQ_ERI_KERNEL") != nullptr && getenv("DQ_ERI_KERNEL")[0] == 'b');
  const bool nodef = getenv("DQ_ERI_KERNEL") != nullptr && getenv("DQ_ERI_KERNEL")[0] == 'p';      // A/B: "plain" = loops in place
  constexpr int L = (int)A + (int)B + (int)C + (int)D;
  // smem = two staged block pairs of kkmax primitive pairs each: kkmax = smem / (2 * 8 * 16 * 8)
  const int kkmax = (int)(smem / (2 * kPairFields * 16 * sizeof(double)));
  // (classes with fewer than two p functions have a Boys evaluation so cheap that the four CTA barriers of the deferred
  //  form cost more than its loops: ncu per class, (ss|ss) 15.1 against 12.3 ms, (ss|sp) 37.9 against 35.2, but (sp|sp)
  //  49.2 against 55.6, (sp|pp) 58.8 against 60.0, (pp|pp) 14.2 against 20.4)
  if (plain && !nodef && L >= 2 && kkmax * kkmax <= kDefMaxSteps) {
    // records + accumulators [256][2] + items [256 * steps]
    const size_t smem_def = smem + 512 * sizeof(unsigned long long) + (size_t)256 * kkmax * kkmax * sizeof(unsigned short) + 16;
    if (s.zero_skip) {
      ensure_dynamic_smem(k_eri_tiles_def<A, B, C, D, true>, 200 * 1024, true);
      k_eri_tiles_def<A, B, C, D, true><<<nt, 256, smem_def, st>>>(tiles, s, pd, store);
    } else {
      ensure_dynamic_smem(k_eri_tiles_def<A, B, C, D, false>, 200 * 1024, true);
      k_eri_tiles_def<A, B, C, D, false><<<nt, 256, smem_def, st>>>(tiles, s, pd, store);
    }
    DQ_LAUNCHED();
    return;
  }
  const size_t smem_b = smem + (size_t)8 * kBoysCap * (L + 2) * sizeof(double);
  if (!plain && smem_b <= 200 * 1024) {
    if (s.zero_skip) {
      ensure_dynamic_smem(k_eri_tiles_b<A, B, C, D, true>, 200 * 1024, true);
      k_eri_tiles_b<A, B, C, D, true><<<nt, 256, smem_b, st>>>(tiles, s, pd, store);
    } else {
      ensure_dynamic_smem(k_eri_tiles_b<A, B, C, D, false>, 200 * 1024, true);
      k_eri_tiles_b<A, B, C, D, false><<<nt, 256, smem_b, st>>>(tiles, s, pd, store);
    }
    DQ_LAUNCHED();
    return;
  }
  if (s.zero_skip) {
    ensure_dynamic_smem(k_eri_tiles<A, B, C, D, true>, 200 * 1024, true);
    k_eri_tiles<A, B, C, D, true><<<nt, 256, smem, st>>>(tiles, s, pd, store);
  } else {
    ensure_dynamic_smem(k_eri_tiles<A, B, C, D, false>, 200 * 1024, true);
    k_eri_tiles<A, B, C, D, false><<<nt, 256, smem, st>>>(tiles, s, pd, store);
  }
  DQ_LAUNCHED();
}
template <bool A, bool B, bool C, bool D>
static void grad_launch(cudaStream_t st, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* pd,
                        int nterms, const double* At, const double* Bt, double* padj, double* pasum, size_t smem, unsigned* dmask,
                        unsigned char* dbin, int tile0) {
  if (dmask) {
    if (s.zero_skip) {
      ensure_dynamic_smem(k_eri_grad_chunks<A, B, C, D, true, true>, 220 * 1024, true);
      k_eri_grad_chunks<A, B, C, D, true, true><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, nterms, At, Bt, padj, pasum, dmask, dbin, tile0);
    } else {
      ensure_dynamic_smem(k_eri_grad_chunks<A, B, C, D, false, true>, 220 * 1024, true);
      k_eri_grad_chunks<A, B, C, D, false, true><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, nterms, At, Bt, padj, pasum, dmask, dbin, tile0);
    }
  } else if (s.zero_skip) {
    ensure_dynamic_smem(k_eri_grad_chunks<A, B, C, D, true, false>, 220 * 1024, true);   // 2 CTAs x ~105 KB
    k_eri_grad_chunks<A, B, C, D, true, false><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, nterms, At, Bt, padj, pasum, nullptr, nullptr, 0);
  } else {
    ensure_dynamic_smem(k_eri_grad_chunks<A, B, C, D, false, false>, 220 * 1024, true);
    k_eri_grad_chunks<A, B, C, D, false, false><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, nterms, At, Bt, padj, pasum, nullptr, nullptr, 0);
  }
  DQ_LAUNCHED();
}

#define DQ_CLASS_SWITCH(cls, CALL)                                                              \
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

template <bool A, bool B, bool C, bool D>
static void eri_compact_launch(cudaStream_t st, const Tile* tiles, int nt, const SystemDev& s, const double* pd, double* store, size_t smem) {
  ensure_dynamic_smem(k_eri_tiles_compact<A, B, C, D>, smem > 64 * 1024 ? smem : 64 * 1024, getenv("DQ_CARVEOUT") != nullptr);
  k_eri_tiles_compact<A, B, C, D><<<nt, 256, smem, st>>>(tiles, s, pd, store); DQ_LAUNCHED();
}
void launch_eri_tiles_compact(cudaStream_t st, int cls, const Tile* tiles, int nt, const SystemDev& s, const double* pd, double* store,
                              int kkmax) {
  if (nt <= 0) return;
  const size_t smem = (size_t)2 * kkmax * kPairFields * 16 * sizeof(double) + (size_t)16 * 256 * sizeof(double) +
                      (size_t)2 * 16 * kkmax * sizeof(unsigned short) + 16;
#define CALL(a, b, c, d) eri_compact_launch<a, b, c, d>(st, tiles, nt, s, pd, store, smem)
  DQ_CLASS_SWITCH(cls, CALL)
#undef CALL
}
void launch_eri_tiles(cudaStream_t st, int cls, const Tile* tiles, int nt, const SystemDev& s, const double* pd, double* store,
                      int kkmax) {
  if (nt <= 0) return;
  const size_t smem = (size_t)2 * kkmax * kPairFields * 16 * sizeof(double);
#define CALL(a, b, c, d) eri_launch<a, b, c, d>(st, tiles, nt, s, pd, store, smem)
  DQ_CLASS_SWITCH(cls, CALL)
#undef CALL
}
template <bool A, bool B, bool C, bool D>
static void grad_sparse_launch(cudaStream_t st, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* pd,
                               int nterms, const double* At, const double* Bt, double* padj, double* pasum, size_t smem) {
  ensure_dynamic_smem(k_eri_grad_sparse<A, B, C, D>, 220 * 1024, true);
  k_eri_grad_sparse<A, B, C, D><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, nterms, At, Bt, padj, pasum); DQ_LAUNCHED();
}
void launch_eri_grad_sparse(cudaStream_t st, int cls, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s,
                            const double* pd, int nterms, const double* At, const double* Bt, double* padj, double* pasum, int kkmax) {
  if (nchunks <= 0) return;
  const size_t smem = eri_grad_smem_bytes(kkmax) - (size_t)2 * 256 * sizeof(double) + (size_t)kSparseChunk * 256 * sizeof(unsigned short);
#define CALL(a, b, c, d) grad_sparse_launch<a, b, c, d>(st, tiles, chunks, nchunks, s, pd, nterms, At, Bt, padj, pasum, smem)
  DQ_CLASS_SWITCH(cls, CALL)
#undef CALL
}
template <bool A, bool B, bool C, bool D>
static void grad_deferred_launch(cudaStream_t st, const Tile* tiles, int tile0, int nt, const SystemDev& s, const double* pd, int nterms,
                                 const double* At, const double* Bt, double* padj, double* pasum, const unsigned* dmask,
                                 const unsigned char* dbin, size_t smem) {
  ensure_dynamic_smem(k_eri_grad_deferred<A, B, C, D>, smem > 100 * 1024 ? smem : 100 * 1024, true);
  k_eri_grad_deferred<A, B, C, D><<<nt, 256, smem, st>>>(tiles, tile0, s, pd, nterms, At, Bt, padj, pasum, dmask, dbin); DQ_LAUNCHED();
}
// second kernel of the deferred gradient form over the tiles [tile0, tile0 + nt) of class `cls` (indices into `tiles` = the
// ket-ordered list the first kernel walked; dmask is indexed the same way)
void launch_eri_grad_deferred(cudaStream_t st, int cls, const Tile* tiles, int tile0, int nt, const SystemDev& s, const double* pd,
                              int nterms, const double* At, const double* Bt, double* padj, double* pasum, const unsigned* dmask,
                              const unsigned char* dbin, int kkmax) {
  if (nt <= 0) return;
  const size_t smem = ((size_t)2 * kkmax * kPairFields * 16 + 256) * sizeof(double) +
                      (size_t)2 * 16 * (kkmax * kAdjFields + 2) * 2 * sizeof(unsigned long long) +
                      (size_t)256 * kkmax * kkmax * sizeof(unsigned short) + 16;
#define CALL(a, b, c, d) grad_deferred_launch<a, b, c, d>(st, tiles, tile0, nt, s, pd, nterms, At, Bt, padj, pasum, dmask, dbin, smem)
  DQ_CLASS_SWITCH(cls, CALL)
#undef CALL
}
void launch_weight_mask(cudaStream_t st, int n2, int nterms, const double* Bt, unsigned char* mask) {
  k_weight_mask<<<cdiv(n2, 256), 256, 0, st>>>(n2, nterms, Bt, mask); DQ_LAUNCHED();
}
void launch_count_nonzero(cudaStream_t st, int n2, const double* A, int* out) {
  cudaMemsetAsync(out, 0, sizeof(int), st);
  k_count_nonzero<<<cdiv(n2, 256), 256, 0, st>>>(n2, A, out); DQ_LAUNCHED();
}
void launch_eri_grad_chunks(cudaStream_t st, int cls, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s,
                            const double* pd, int nterms, const double* At, const double* Bt, double* padj, double* pasum, int kkmax,
                            unsigned* dmask, unsigned char* dbin, int tile0) {
  if (nchunks <= 0) return;
  const size_t smem = eri_grad_smem_bytes(kkmax);
#define CALL(a, b, c, d) grad_launch<a, b, c, d>(st, tiles, chunks, nchunks, s, pd, nterms, At, Bt, padj, pasum, smem, dmask, dbin, tile0)
  DQ_CLASS_SWITCH(cls, CALL)
#undef CALL
}
size_t eri_grad_smem_bytes(int kkmax) {
  const int kc = kkmax < kGradKetChunk ? kkmax : kGradKetChunk;
  return ((size_t)kkmax * kPairFields * 16 + (size_t)(kc * kAdjFields + 2) * 256) * sizeof(double);
}
void launch_scatter_packed(cudaStream_t st, const Tile* tiles, int nt, const SystemDev& s, const int* f_orig, const double* store,
                           double* packed) {
  if (nt <= 0) return;
  k_scatter_packed<<<nt, 256, 0, st>>>(tiles, s, f_orig, store, packed); DQ_LAUNCHED();
}
void launch_jk_chunks(cudaStream_t st, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* store,
                      const double* D, double* Jacc, double* Kacc) {
  if (nchunks <= 0) return;
  const size_t smem2 = ((size_t)(s.det ? 24 : 16) * (s.N | 1) + 16) * sizeof(double) + 256 * 16;
  const size_t smem3 = smem2 + (size_t)8 * kJkSlots * 2 * kTile * sizeof(double) + 8 * kJkSlots * sizeof(unsigned long long);
  static const bool ldg_kernel = getenv("DQ_JK_LDG") != nullptr;      // A/B switch: the register-path (LDG) kernel
  // A/B switch DQ_JK_PRIVATE=1: private K strips per warp (no atomics in the loop, but 198 KB of shared memory = one CTA
  // of 8 warps per SM: measured 0.96 ms per build on the dense (H2O)_32 cluster against 0.74 ms for the shared strips)
  static const bool private_strips = getenv("DQ_JK_PRIVATE") != nullptr;
  const size_t ring = (size_t)8 * kJkSlots * 2 * kTile * sizeof(double) + 8 * kJkSlots * sizeof(unsigned long long);
  const size_t smemp = ((size_t)72 * (s.N | 1) + 16) * sizeof(double) + 256 * 16 + ring;     // private strips per warp
  if (!ldg_kernel && private_strips && smemp <= 227 * 1024) {
    ensure_dynamic_smem(k_jk_stream<true>, smemp, true);
    k_jk_stream<true><<<nchunks, 256, smemp, st>>>(tiles, chunks, s, store, D, Jacc, Kacc); DQ_LAUNCHED();
    return;
  }
  if (!ldg_kernel && smem3 <= 227 * 1024) {
    ensure_dynamic_smem(k_jk_stream<false>, smem3, true);
    k_jk_stream<false><<<nchunks, 256, smem3, st>>>(tiles, chunks, s, store, D, Jacc, Kacc); DQ_LAUNCHED();
    return;
  }
  ensure_dynamic_smem(k_jk_cols, smem2);
  k_jk_cols<<<nchunks, 256, smem2, st>>>(tiles, chunks, s, store, D, Jacc, Kacc); DQ_LAUNCHED();
}
size_t jk_direct_smem_bytes(int kkmax, int N, int det) {
  return ((size_t)2 * kkmax * kPairFields * 16 + 2 * 16 * 17 + 16 + (size_t)(det ? 24 : 16) * (N | 1)) * sizeof(double);
}
template <bool A, bool B, bool C, bool D>
static void jk_direct_launch(cudaStream_t st, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* pd,
                             const double* Dm, double* Jacc, double* Kacc, int kkmax, size_t smem) {
  if (s.zero_skip) {
    ensure_dynamic_smem(k_jk_direct<A, B, C, D, true>, smem > 100 * 1024 ? smem : 100 * 1024, true);
    k_jk_direct<A, B, C, D, true><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, Dm, Jacc, Kacc, kkmax);
  } else {
    ensure_dynamic_smem(k_jk_direct<A, B, C, D, false>, smem > 100 * 1024 ? smem : 100 * 1024, true);
    k_jk_direct<A, B, C, D, false><<<nchunks, 256, smem, st>>>(tiles, chunks, s, pd, Dm, Jacc, Kacc, kkmax);
  }
  DQ_LAUNCHED();
}
void launch_jk_direct(cudaStream_t st, int cls, const Tile* tiles, const int2* chunks, int nchunks, const SystemDev& s, const double* pd,
                      const double* D, double* Jacc, double* Kacc, int kkmax) {
  if (nchunks <= 0) return;
  const size_t smem = jk_direct_smem_bytes(kkmax, s.N, s.det);
#define CALL(a, b, c, d) jk_direct_launch<a, b, c, d>(st, tiles, chunks, nchunks, s, pd, D, Jacc, Kacc, kkmax, smem)
  DQ_CLASS_SWITCH(cls, CALL)
#undef CALL
}
void launch_det_finalize(cudaStream_t st, long long n, const double* in, double* out, double beta) {
  if (n <= 0) return;
  k_det_finalize<<<cdiv(n, 256), 256, 0, st>>>(n, in, out, beta); DQ_LAUNCHED();
}
void launch_pair_bwd(cudaStream_t st, const SystemDev& s, const double* padj, const double* pasum, double* g_alpha, double* g_coef,
                     double* g_xyz) {
  k_pair_bwd<<<s.nbp, 128, 0, st>>>(s, padj, pasum, g_alpha, g_coef, g_xyz); DQ_LAUNCHED();
}

void launch_gemm(cudaStream_t st, bool ta, bool tb, int M, int N, int K, const double* A, int lda, const double* B, int ldb,
                 double* C, int ldc, double alpha, double beta) {
  dim3 grid(cdiv(N, 16), cdiv(M, 16));
  if (!ta && !tb) k_gemm<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
  else if (ta && !tb) k_gemm<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
  else if (!ta && tb) k_gemm<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
  else k_gemm<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
  DQ_LAUNCHED();
}
constexpr int kJacobiPackedMax = 236;   // packed upper triangle + rotation arrays must fit 227 KB of shared memory
constexpr int kJacobiSweeps = 30;
size_t eigh_workspace_bytes(int n, int batch) {
  const int ne = n + (n & 1), m = ne / 2;
  size_t b = 0;
  b += (size_t)batch * kJacobiSweeps * (ne - 1 > 0 ? ne - 1 : 1) * m * sizeof(double2);   // rotation log
  b += (size_t)batch * n * sizeof(int) + (size_t)batch * sizeof(int) + 256;                // ranks, round counts
  b += (size_t)(m * (m - 1) / 2 + 1) * sizeof(unsigned short) + 256;                        // pair-of-pairs table
  b += (size_t)batch * 2 * kOsCluster * ((n + 2 * kOsCluster - 1) / (2 * kOsCluster)) * sizeof(double) + 512;   // cluster solver: eigenvalues by position
  return b;
}
// cluster (one-sided) solver, optionally warm-started from V0 (n x n, orthogonal, e.g. the previous eigenvectors)
bool eigh_cluster_applicable(int n) {
  static const bool off = getenv("DQ_EIGH_PACKED") != nullptr;      // A/B switch: the single-CTA two-sided kernels
  return !off && n >= kOsMinN && n <= kJacobiPackedMax;
}
void launch_eigh_cluster(cudaStream_t st, int n, int batch, const double* A, const double* V0, double* Vout, double* w, void* work) {
  const int ne = n + (n & 1), m = ne / 2;
  char* base = (char*)work;
  base += (size_t)batch * kJacobiSweeps * (ne - 1) * m * sizeof(double2);
  base += (size_t)batch * n * sizeof(int);
  base += (((size_t)batch * sizeof(int) + 255) / 256) * 256;
  base += (((size_t)(m * (m - 1) / 2 + 1) * sizeof(unsigned short) + 255) / 256) * 256;
  base = (char*)(((size_t)base + 255) / 256 * 256);          // (the int arrays above leave it 4-byte aligned for odd n)
  double* lam = (double*)base;
  const int wcol = (n + 2 * kOsCluster - 1) / (2 * kOsCluster);
  const size_t smem = (size_t)8 * wcol * n * sizeof(double);
  ensure_dynamic_smem(k_eigh_onesided, smem);
  static const int sweeps = getenv("DQ_EIGH_DEBUG") ? -40 : 40;      // negative: print the per-sweep rotation counts
  k_eigh_onesided<<<batch * kOsCluster, kOsThreads, smem, st>>>(n, wcol, A, V0, Vout, w, lam, sweeps); DQ_LAUNCHED();
}
bool eigh_cluster_applicable(int n);
void launch_eigh_cluster(cudaStream_t st, int n, int batch, const double* A, const double* V0, double* Vout, double* w, void* work);
void launch_eigh(cudaStream_t st, int n, int batch, double* A, double* Vwork, double* Vout, double* w, void* work) {
  if (work != nullptr && eigh_cluster_applicable(n)) { launch_eigh_cluster(st, n, batch, A, nullptr, Vout, w, work); return; }
  if (n > kJacobiPackedMax || n < 2 || work == nullptr) {
    const int m = (n + 1) / 2;
    const size_t smem = (size_t)2 * m * sizeof(double) + (size_t)2 * m * sizeof(int);
    int threads = 1024;
    if (n <= 16) threads = 128; else if (n <= 32) threads = 256; else if (n <= 64) threads = 512;
    k_eigh_jacobi<<<batch, threads, smem, st>>>(n, A, Vwork, Vout, w, 60); DQ_LAUNCHED();
    return;
  }
  const int ne = n + (n & 1), m = ne / 2;
  char* base = (char*)work;
  double2* log = (double2*)base; base += (size_t)batch * kJacobiSweeps * (ne - 1) * m * sizeof(double2);
  int* rank = (int*)base; base += (size_t)batch * n * sizeof(int);
  int* meta = (int*)base; base += (((size_t)batch * sizeof(int) + 255) / 256) * 256;
  unsigned short* tab = (unsigned short*)base;
  const size_t smem1 = (size_t)(ne * (ne + 1) / 2 + 2 * m) * sizeof(double) + (size_t)2 * m * sizeof(int);
  ensure_dynamic_smem(k_jacobi_packed, smem1);
  int threads = 1024;
  if (n <= 16) threads = 64; else if (n <= 32) threads = 128; else if (n <= 64) threads = 256; else if (n <= 128) threads = 512;
  k_jacobi_packed<<<batch, threads, smem1, st>>>(n, A, tab, log, meta, w, rank, kJacobiSweeps); DQ_LAUNCHED();
  int wpb = n < 4 ? n : 4;               // warps (rows) per CTA: many small CTAs spread the log reads over the SMs
  const size_t smem2 = (size_t)wpb * ne * sizeof(double);
  ensure_dynamic_smem(k_jacobi_vectors, smem2);
  dim3 grid((n + wpb - 1) / wpb, batch);
  k_jacobi_vectors<<<grid, wpb * 32, smem2, st>>>(n, log, meta, rank, Vout, kJacobiSweeps); DQ_LAUNCHED();
}
void launch_scale_cols_rsqrt(cudaStream_t st, int n, const double* A, const double* w, double* B) {
  k_scale_cols_rsqrt<<<cdiv((long long)n * n, 256), 256, 0, st>>>(n, A, w, B); DQ_LAUNCHED();
}
void launch_fock_combine(cudaStream_t st, int n, const int* f_block, const double* H, const double* Jacc, const double* Kacc,
                         double* F, double* G) {
  k_fock_combine<<<cdiv((long long)n * n, 256), 256, 0, st>>>(n, f_block, H, Jacc, Kacc, F, G); DQ_LAUNCHED();
}
void launch_dot3(cudaStream_t st, int n, const double* A, const double* B, const double* C, double* out) {
  k_dot3<<<1, 1024, 0, st>>>(n, A, B, C, out); DQ_LAUNCHED();
}
void launch_make_y(cudaStream_t st, int n, int ne, const double* Z, const double* eps, double* Y) {
  k_make_y<<<cdiv((long long)n * n, 256), 256, 0, st>>>(n, ne, Z, eps, Y); DQ_LAUNCHED();
}
void launch_phi_rsqrt(cudaStream_t st, int n, const double* w, double* M) {
  k_phi_rsqrt<<<cdiv((long long)n * n, 256), 256, 0, st>>>(n, w, M); DQ_LAUNCHED();
}
void launch_axpby(cudaStream_t st, int n, double a, const double* X, double b, double* Y) {
  k_axpby<<<cdiv(n, 256), 256, 0, st>>>(n, a, X, b, Y); DQ_LAUNCHED();
}
void launch_symmetrize(cudaStream_t st, int n, double* A) {
  k_symmetrize<<<cdiv((long long)n * n, 256), 256, 0, st>>>(n, A); DQ_LAUNCHED();
}
void launch_add_transpose(cudaStream_t st, int n, const double* A, double* B) {
  k_add_transpose<<<cdiv((long long)n * n, 256), 256, 0, st>>>(n, A, B); DQ_LAUNCHED();
}
void launch_exp(cudaStream_t st, int n, const double* x, double* y) {
  k_exp<<<cdiv(n, 256), 256, 0, st>>>(n, x, y); DQ_LAUNCHED();
}
void launch_mul(cudaStream_t st, int n, const double* a, const double* b, double* y) {
  k_mul<<<cdiv(n, 256), 256, 0, st>>>(n, a, b, y); DQ_LAUNCHED();
}
void launch_boys(cudaStream_t st, int n, const double* t, double* F, double* dF) {
  k_boys<<<cdiv(n, 128), 128, 0, st>>>(n, t, F, dF); DQ_LAUNCHED();
}
void launch_fp64_peak(cudaStream_t st, int blocks, int iters, double* out) {
  k_fp64_peak<<<blocks, 256, 0, st>>>(iters, out); DQ_LAUNCHED();
}

}  // namespace dq
