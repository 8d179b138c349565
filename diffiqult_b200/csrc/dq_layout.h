// diffiqult_b200 — device data layout shared by the kernels (dq_kernels.cu) and the host driver (dq_api.cu).
//
// Basis functions are permuted into SORTED order (key: type s/px/py/pz, contraction length) and cut into
// FUNCTION BLOCKS of at most 4 functions of one type and one contraction length.  A TILE is a quadruple of
// blocks (I,J | K,L) with I<=J, K<=L, pair(I,J) <= pair(K,L): 4^4 = 256 contracted quartets, one per thread
// of a 256-thread CTA, all of one angular class, so a CTA never diverges on the integral class.
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace dq {

constexpr int kBlk = 4;                 // functions per block
constexpr int kTile = kBlk * kBlk * kBlk * kBlk;   // 256 quartets per tile
constexpr int kPairLanes = kBlk * kBlk;  // 16 function pairs per block pair
constexpr int kPairFields = 8;          // P[3], gamma, 1/gamma, Kab, (P-A)_axis(a), (P-B)_axis(b) per primitive pair
constexpr int kAdjFields = 5;           // Pbar[3], gammabar, Kbar per primitive pair

struct BlockPair {       // one per (I<=J)
  int sI, nI, sJ, nJ;    // first sorted function and count of each block
  int axI, axJ;          // axis (0..2) if the block is p-type, else -1
  int KI, KJ;            // contraction lengths
  int KK;                // KI*KJ primitive pairs
  int bI, bJ;            // block ids
  int pad_;
  long long off;         // offset (in units of 16 doubles) of this block pair's primitive-pair records
};

// pair-data record layout: pd[(off + pp*kPairFields + field)*16 + lane], lane = il*4+jl
// adjoint layout:          padj[(off5 + pp*kAdjFields + field)*16 + lane] with off5 = off/kPairFields*kAdjFields
// contracted (pa,pb) sums: pasum[(bp*2 + which)*16 + lane]

struct Tile { int bp1, bp2; };

struct SystemDev {
  int N, nprim, natoms, nblocks, nbp;
  const int* f_off;        // [N+1] first primitive of each sorted function
  const int* f_type;       // [N]   0 s, 1..3 p
  const int* f_block;      // [N]   block id
  const double* f_xyz;     // [N][3]
  const double* p_alpha;   // [nprim] exponents
  const double* p_coef;    // [nprim] NORMALISED contraction coefficients
  const int* Z;            // [natoms]
  const double* C;         // [natoms][3]
  const BlockPair* bp;     // [nbp]
  const double* bp_kmax;   // [nbp]  max |K_ab| over the primitive-pair records of each block pair
  double screen;           // opt-in tile screening threshold (0: off)
  unsigned long long* skipped;   // counter of skipped tiles (screening on)
  int zero_skip;                 // 1 (default): primitive pairs whose Gaussian product underflowed to exactly 0 are left out of the loops
  unsigned long long* underflowed;   // counter: such primitive pairs (valid lanes), written by k_pairs
  const unsigned char* wmask;    // [N*N] gradient pass: is B_t[i,j] nonzero for some term t (every product in a weight has a B factor)
  int det;                       // 1: accumulators are pairs of integer limbs (order-independent sums, dq_math.cuh det_add), 0: doubles + floating-point atomics
  unsigned long long* zero_weight;   // counter of contracted quartets the gradient pass skipped because their weight is exactly 0
};

// Batch of molecules of ONE structure (dq_batch.cu).  Functions stay in the caller's order; function pairs x = (i <= j)
// in the reference's packed order (Tools.py:17-22).  Inputs are per molecule ([mol][...]); the primitive-pair records and
// their adjoints are molecule-fastest ([...][mol]) because the quartet kernels put neighbouring molecules in neighbouring lanes.
struct BatchDev {
  int nmol, N, nprim, natoms, M, nrec, neri, kkmax, ne;
  int det;                 // fixed-point (order-independent) gradient accumulators, see SystemDev::det
  const int* f_off;        // [N+1]
  const int* f_type;       // [N]  0 s, 1..3 p
  const int* Z;            // [natoms]
  const int2* pairs;       // [M]  (i, j), i <= j
  const int* pair_off;     // [M+1] first primitive-pair record of each function pair
  const int* rec_pair;     // [nrec] function pair of each record
  const double* atoms;     // [nmol][natoms*3]
  const double* xyz;       // [nmol][N*3]
  const double* craw;      // [nmol][nprim] raw contraction coefficients
  double* alpha;           // [nmol][nprim] exponents (after exp() when the caller passes ln alpha)
  double* cn;              // [nmol][nprim] normalised coefficients
  double* prec;            // [nrec][kPairFields][nmol]
};

}  // namespace dq
