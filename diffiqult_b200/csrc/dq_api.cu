// diffiqult_b200 — host driver behind the C ABI (include/diffiqult_b200.h).
//
// Everything numerical runs in the kernels of dq_kernels.cu; this file sorts the basis into function blocks,
// builds the tile lists, owns device memory, sequences the launches of the SCF (Energy.py:138-195) and of its
// exact reverse sweep (DESIGN.md "gradient"), and moves results back in the caller's (reference) ordering.
// There is no CPU fallback: without a CUDA device every entry point returns an error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/diffiqult_b200.h"
#include "dq_launch.h"
#include "dq_layout.h"

namespace dq {

static thread_local std::string t_err;
static thread_local dq_stats t_stats;

struct Fail { int code; };
static void fail(int code, const char* fmt, const char* a = "") {
  char buf[512];
  snprintf(buf, sizeof buf, fmt, a);
  t_err = buf;
  throw Fail{code};
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) fail(-2, "CUDA error: %s", cudaGetErrorString(e_)); } while (0)
// launch errors (bad configuration, attribute not granted) do not surface through the stream: poll them where a phase ends,
// so that nothing stale is left for the caller's next CUDA call to trip over
static void check_launches(const char* where) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", where, cudaGetErrorString(e));
    fail(-2, "CUDA error after %s", buf);
  }
}

// A Boys-function loop that ran out of terms raised the device flag (non-finite argument; the reference exits the process
// there, Tools.py:124-125, 150-151): report it as status -5 and clear the flag.  Call after the stream has been drained.
static void check_boys(cudaStream_t st) {
  int* flag = boys_fail_flag();
  if (!flag) return;
  int h = 0;
  if (cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return;
  if (h) {
    cudaMemsetAsync(flag, 0, sizeof h, st);
    cudaStreamSynchronize(st);
    t_err = "Boys function: series / continued fraction did not converge (non-finite argument); the reference exits here (Tools.py:124-125, 150-151)";
    throw Fail{-5};
  }
}

// ---------------------------------------------------------------------------------------------
// Device-memory pool: freed blocks are kept (per device, by exact size) and handed out again, so that repeated
// calls on the same problem -- every BFGS step, every bench step -- do not pay cudaMalloc/cudaFree of the 2.6 GB tile
// store.  Blocks return to the pool only after the owning call has synchronised its stream.
struct Pool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;
  size_t cached_bytes = 0;
  void* get(int dev, size_t bytes) {
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.find({dev, bytes});
      if (it != free_blocks.end()) { void* p = it->second; free_blocks.erase(it); cached_bytes -= bytes; return p; }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {          // out of memory: drop the cache and retry once
      cudaGetLastError();
      release();
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return nullptr;
    return p;
  }
  void put(int dev, size_t bytes, void* p) {
    std::lock_guard<std::mutex> g(mu);
    free_blocks.insert({{dev, bytes}, p});
    cached_bytes += bytes;
  }
  void release() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
};
static Pool g_pool;

static void sync_side_streams();
struct Workspace {
  cudaStream_t st = 0;
  int dev = 0;
  std::vector<std::pair<void*, size_t>> owned;
  ~Workspace() {
    cudaStreamSynchronize(st);
    sync_side_streams();      // an error between ForkJoin::begin and ::end leaves class kernels running on the side streams
    for (auto& b : owned) g_pool.put(dev, b.second, b.first);
  }
  template <class T> T* alloc(size_t n, bool zero = false) {
    if (n == 0) n = 1;
    const size_t bytes = (n * sizeof(T) + 255) / 256 * 256;
    void* p = g_pool.get(dev, bytes);
    if (!p) fail(-2, "CUDA error: %s", "out of device memory");
    owned.push_back({p, bytes});
    if (zero) CU(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    return (T*)p;
  }
  template <class T> T* upload(const std::vector<T>& h) {
    T* d = alloc<T>(h.size());
    if (!h.empty()) CU(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return d;
  }
};

struct Timer {
  cudaEvent_t a, b; cudaStream_t st;
  explicit Timer(cudaStream_t s) : st(s) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
  double stop() { cudaEventRecord(b, st); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); cudaEventDestroy(a); cudaEventDestroy(b); return ms; }
};

// The class launches of an ERI pass are independent of each other: they are issued on a few side streams forked from the
// caller's stream and joined back, so that the last wave of one class overlaps the first of the next (with chunks of 12
// tiles a gradient CTA lives ~1 ms, and a rank of an 8-GPU job has only ~10 waves per class).
struct ForkJoin {
  static constexpr int kSide = 4;
  cudaStream_t side[kSide] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[kSide] = {nullptr, nullptr, nullptr, nullptr};
  int dev = -1;
  bool open = false;        // begin() without end(): work may be in flight on the side streams
  void init(int d) {
    if (dev == d) return;
    for (int i = 0; i < kSide; ++i) {
      cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    dev = d;
  }
  void begin(cudaStream_t main) { open = true; cudaEventRecord(fork, main); for (int i = 0; i < kSide; ++i) cudaStreamWaitEvent(side[i], fork, 0); }
  cudaStream_t lane(int i) const { return side[i % kSide]; }
  void end(cudaStream_t main) { for (int i = 0; i < kSide; ++i) { cudaEventRecord(join[i], side[i]); cudaStreamWaitEvent(main, join[i], 0); } open = false; }
  void drain() { if (open) { for (int i = 0; i < kSide; ++i) cudaStreamSynchronize(side[i]); open = false; } }
};
static thread_local ForkJoin t_fork;
static void sync_side_streams() { t_fork.drain(); }

// Tile lists depend only on the basis STRUCTURE (types, contraction lengths) and on (world, rank), not on the
// parameter values: each host thread keeps the last plan, so the repeated calls of an optimisation or of the bench
// neither rebuild nor re-upload the 1.27 M-entry tile list.
struct TilePlan {
  int dev = 0, world = 1, rank = 0;
  std::vector<int> ltype, contr;
  std::vector<Tile> tiles;      // this rank's tiles, grouped by class
  std::vector<int2> chunks;     // runs of tiles sharing the bra block pair (J/K contraction)
  int cls_start[17];
  int chunk_cls_start[17];      // chunk ranges per class
  long long n_quartets = 0, n_prim_quartets = 0;
  // the same tiles re-ordered for the gradient pass: per class by KET block pair, cut into chunks sharing the ket
  std::vector<Tile> gtiles;
  std::vector<int2> gchunks;
  int gchunk_cls_start[17];
  std::vector<int2> dchunks;    // integral-direct J/K: runs of <= kDirectChunk tiles sharing the bra block pair (a CTA evaluates AND contracts them)
  int dchunk_cls_start[17];
  int2* d_dchunks = nullptr;
  std::vector<int2> gchunks32;  // the same order cut into chunks of <= 20 tiles (sparse-weight gradient kernel: its work list is in shared memory)
  int gchunk32_cls_start[17];
  int2* d_gchunks32 = nullptr;
  Tile* d_tiles = nullptr;
  int2* d_chunks = nullptr;
  Tile* d_gtiles = nullptr;
  int2* d_gchunks = nullptr;
  ~TilePlan() {
    if (d_tiles) cudaFree(d_tiles);
    if (d_chunks) cudaFree(d_chunks);
    if (d_gtiles) cudaFree(d_gtiles);
    if (d_gchunks) cudaFree(d_gchunks);
    if (d_gchunks32) cudaFree(d_gchunks32);
    if (d_dchunks) cudaFree(d_dchunks);
  }
};
static thread_local std::shared_ptr<TilePlan> t_plan;

// the sorted / blocked system on the device plus the maps back to the caller's ordering
struct Prepared {
  Workspace ws;
  SystemDev sd;
  int N = 0, nprim = 0, ne = 0, nblocks = 0, nbp = 0, kmax = 0, kkmax = 0;
  std::vector<int> f_orig;      // sorted function -> original function
  std::vector<int> p_orig;      // sorted primitive -> original primitive
  std::vector<BlockPair> bps;
  std::shared_ptr<TilePlan> tp;
  long long pd_len = 0;         // doubles in the pair-record buffer
  long long n_quartets = 0, n_prim_quartets = 0;
  // device arrays
  double *d_alpha = nullptr, *d_lnalpha = nullptr, *d_craw = nullptr, *d_cn = nullptr, *d_xyz = nullptr, *d_pd = nullptr;
  int *d_foff = nullptr, *d_ftype = nullptr, *d_fblock = nullptr, *d_forig = nullptr;
  Tile* d_tiles = nullptr;
  int2* d_chunks = nullptr; int nchunks = 0;
  const int* cls_start = nullptr;
  size_t ntiles = 0;
  double* d_store = nullptr;
  unsigned long long* d_skipped = nullptr;
  unsigned long long* d_underflowed = nullptr;
  bool eri_compact = false;     // value pass: compact the surviving primitive pairs per tile (many underflowed Gaussian products)
  bool direct = false;          // integral-direct Fock builds: every J/K contraction re-evaluates the tiles (k_jk_direct), no tile store
  bool log_alpha = false;
  bool plan_only = false;       // host planning only (dq_plan): no CUDA call is made
  int world = 1, rank = 0;
  dq_allreduce_fn allreduce = nullptr; void* ar_user = nullptr;
};

static void need_device(int dev) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); fail(-3, "no CUDA device: diffiqult_b200 has no CPU path"); }
  if (dev < 0 || dev >= ndev) fail(-1, "device ordinal out of range");
  CU(cudaSetDevice(dev));
}

static void check_system(const dq_system* s) {
  if (!s) fail(-1, "null system");
  if (s->nbasis <= 0 || s->nprim <= 0) fail(-1, "empty basis (nbasis == 0: a 'shifted' System_mol never fills list_contr)");
  int np = 0;
  for (int i = 0; i < s->nbasis; ++i) {
    if (s->contr[i] <= 0) fail(-1, "contr[i] must be positive");
    if (s->ltype[i] < 0 || s->ltype[i] > 3) fail(-1, "ltype must be 0 (s) or 1..3 (p): only s and p Gaussians are supported");
    np += s->contr[i];
  }
  if (np != s->nprim) fail(-1, "sum(contr) != nprim");
}

static void prepare(Prepared& P, const dq_system* s, const dq_options* opt, bool need_tiles, bool normalise) {
  check_system(s);
  int dev = opt ? opt->device : 0;
  if (!P.plan_only) need_device(dev);
  P.ws.dev = dev;
  P.ws.st = opt ? (cudaStream_t)opt->stream : 0;
  cudaStream_t st = P.ws.st;
  P.world = (opt && opt->world_size > 1) ? opt->world_size : 1;
  P.rank = (opt && opt->world_size > 1) ? opt->rank : 0;
  P.allreduce = opt ? opt->allreduce : nullptr;
  P.ar_user = opt ? opt->allreduce_user : nullptr;
  if (P.world > 1 && !P.allreduce) fail(-1, "world_size > 1 needs an allreduce callback");
  static thread_local int boys_dev = -1;
  if (!P.plan_only && boys_dev != dev) { if (upload_boys_constants() != 0) fail(-2, "constant upload failed"); boys_dev = dev; }

  const int N = s->nbasis;
  P.N = N; P.nprim = s->nprim; P.ne = s->ne; P.log_alpha = s->log_alpha != 0;
  std::vector<int> off(N + 1, 0);
  for (int i = 0; i < N; ++i) off[i + 1] = off[i] + s->contr[i];
  // sort functions by (type, contraction length); stable so that equal keys keep the caller's order
  P.f_orig.resize(N);
  for (int i = 0; i < N; ++i) P.f_orig[i] = i;
  // Inside a (type, contraction length) group the order decides which four functions share a block, i.e. which quartets
  // share a warp.  All lanes of a warp walk the same primitive indices, so a warp's Boys arguments t = rho |PQ|^2 are alike -
  // and its lanes take the same branch of the Boys function with similar trip counts - when the functions of a block have
  // like exponents and nearby centres: sort by the largest exponent (quarter-octave buckets), then along a Morton curve.
  // The tile plan depends on the group sizes only, so it stays cached while parameters move.  DQ_SORT=0: caller's order.
  static const int sort_mode = getenv("DQ_SORT") ? atoi(getenv("DQ_SORT")) : 1;
  std::vector<long long> key2(N, 0);
  if (sort_mode >= 1) {
    double lo[3] = {1e300, 1e300, 1e300};
    for (int i = 0; i < N; ++i) for (int c = 0; c < 3; ++c) lo[c] = std::min(lo[c], s->xyz[3 * i + c]);
    for (int i = 0; i < N; ++i) {
      double amax = 0.0;
      for (int p = 0; p < s->contr[i]; ++p) {
        const double a = s->log_alpha ? exp(s->alpha[off[i] + p]) : s->alpha[off[i] + p];
        amax = std::max(amax, a);
      }
      long long bucket = amax > 0.0 ? (long long)floor(4.0 * log2(amax)) + 4096 : 0;
      unsigned long long morton = 0;
      unsigned g[3];
      for (int c = 0; c < 3; ++c) { const double q = (s->xyz[3 * i + c] - lo[c]) / 2.0; g[c] = q > 0 ? (q < 1023 ? (unsigned)q : 1023u) : 0u; }   // 2-bohr cells
      for (int b = 9; b >= 0; --b) for (int c = 2; c >= 0; --c) morton = (morton << 1) | ((g[c] >> b) & 1u);
      key2[i] = sort_mode == 1 ? (bucket << 32) | (long long)morton : (long long)morton;
    }
  }
  std::stable_sort(P.f_orig.begin(), P.f_orig.end(), [&](int a, int b) {
    if (s->ltype[a] != s->ltype[b]) return s->ltype[a] < s->ltype[b];
    if (s->contr[a] != s->contr[b]) return s->contr[a] < s->contr[b];
    return key2[a] < key2[b];
  });
  std::vector<int> foff(N + 1, 0), ftype(N), fblock(N);
  std::vector<double> xyz(3 * N), alpha(s->nprim), craw(s->nprim);
  P.p_orig.resize(s->nprim);
  P.kmax = 0;
  for (int si = 0; si < N; ++si) {
    const int o = P.f_orig[si];
    ftype[si] = s->ltype[o];
    foff[si + 1] = foff[si] + s->contr[o];
    P.kmax = std::max(P.kmax, (int)s->contr[o]);
    for (int c = 0; c < 3; ++c) xyz[3 * si + c] = s->xyz[3 * o + c];
    for (int p = 0; p < s->contr[o]; ++p) {
      P.p_orig[foff[si] + p] = off[o] + p;
      alpha[foff[si] + p] = s->alpha[off[o] + p];
      craw[foff[si] + p] = s->coef[off[o] + p];
    }
  }
  // function blocks
  struct Blk { int s, n, type, K; };
  std::vector<Blk> blks;
  for (int si = 0; si < N;) {
    const int ty = ftype[si], K = foff[si + 1] - foff[si];
    int n = 1;
    while (n < kBlk && si + n < N && ftype[si + n] == ty && foff[si + n + 1] - foff[si + n] == K) ++n;
    for (int q = 0; q < n; ++q) fblock[si + q] = (int)blks.size();
    blks.push_back({si, n, ty, K});
    si += n;
  }
  const int nb = (int)blks.size();
  P.nblocks = nb; P.nbp = nb * (nb + 1) / 2;
  P.bps.resize(P.nbp);
  long long offp = 0;
  P.kkmax = 0;
  {
    int x = 0;
    for (int I = 0; I < nb; ++I)
      for (int J = I; J < nb; ++J, ++x) {
        BlockPair& b = P.bps[x];
        b.sI = blks[I].s; b.nI = blks[I].n; b.sJ = blks[J].s; b.nJ = blks[J].n;
        b.axI = blks[I].type - 1; b.axJ = blks[J].type - 1;
        b.KI = blks[I].K; b.KJ = blks[J].K; b.KK = b.KI * b.KJ; b.bI = I; b.bJ = J; b.pad_ = 0;
        b.off = offp; offp += (long long)b.KK * kPairFields;
        P.kkmax = std::max(P.kkmax, b.KK);
      }
  }
  P.pd_len = offp * 16;

  // device copies
  Workspace& ws = P.ws;
  if (!P.plan_only) {
  P.d_foff = ws.upload(foff); P.d_ftype = ws.upload(ftype); P.d_fblock = ws.upload(fblock); P.d_forig = ws.upload(P.f_orig);
  P.d_xyz = ws.upload(xyz); P.d_craw = ws.upload(craw);
  std::vector<int> Z(s->natoms > 0 ? s->natoms : 1, 0);
  std::vector<double> Cn(s->natoms > 0 ? 3 * s->natoms : 3, 0.0);
  for (int a = 0; a < s->natoms; ++a) { Z[a] = s->charges[a]; for (int c = 0; c < 3; ++c) Cn[3 * a + c] = s->atoms[3 * a + c]; }
  int* dZ = ws.upload(Z); double* dC = ws.upload(Cn);
  BlockPair* dbp = ws.upload(P.bps);
  if (P.log_alpha) {
    P.d_lnalpha = ws.upload(alpha);
    P.d_alpha = ws.alloc<double>(s->nprim);
    launch_exp(st, s->nprim, P.d_lnalpha, P.d_alpha);           // Energy.py:120-121
  } else {
    P.d_alpha = ws.upload(alpha);
  }
  if (normalise) {
    P.d_cn = ws.alloc<double>(s->nprim);
    launch_normalize(st, N, P.d_foff, P.d_ftype, P.d_alpha, P.d_craw, P.d_cn);   // Energy.py:132
  } else {
    P.d_cn = P.d_craw;
  }
  SystemDev& sd = P.sd;
  sd.N = N; sd.nprim = s->nprim; sd.natoms = s->natoms; sd.nblocks = nb; sd.nbp = P.nbp;
  sd.f_off = P.d_foff; sd.f_type = P.d_ftype; sd.f_block = P.d_fblock; sd.f_xyz = P.d_xyz;
  sd.p_alpha = P.d_alpha; sd.p_coef = P.d_cn; sd.Z = dZ; sd.C = dC; sd.bp = dbp;
  sd.bp_kmax = nullptr; sd.screen = 0.0; sd.skipped = nullptr; sd.zero_weight = nullptr; sd.underflowed = nullptr; sd.wmask = nullptr;
  sd.zero_skip = getenv("DQ_NO_ZERO_SKIP") ? 0 : 1;
  {
    const char* acc = getenv("DQ_ACCUM");
    sd.det = acc ? (acc[1] == 'i' ? 1 : 0) : ((opt && opt->accumulate == 2) ? 0 : 1);     // "fixed" / "float"
  }
  }

  if (!need_tiles) return;
  // tiles of this rank, grouped by angular class (pI pJ | pK pL)
  std::shared_ptr<TilePlan> tp = t_plan;
  const bool hit = !P.plan_only && tp && tp->dev == dev && tp->world == P.world && tp->rank == P.rank &&
                   tp->ltype.size() == (size_t)N && std::equal(tp->ltype.begin(), tp->ltype.end(), s->ltype) &&
                   std::equal(tp->contr.begin(), tp->contr.end(), s->contr);
  if (!hit) {
    tp = std::make_shared<TilePlan>();
    tp->dev = dev; tp->world = P.world; tp->rank = P.rank;
    tp->ltype.assign(s->ltype, s->ltype + N); tp->contr.assign(s->contr, s->contr + N);
    std::vector<std::vector<Tile>> by_cls(16);
    std::vector<long long> seen(16, 0);
    for (int b1 = 0; b1 < P.nbp; ++b1) {
      const BlockPair& x = P.bps[b1];
      for (int b2 = b1; b2 < P.nbp; ++b2) {
        const BlockPair& y = P.bps[b2];
        const int cls = (x.axI >= 0) * 8 + (x.axJ >= 0) * 4 + (y.axI >= 0) * 2 + (y.axJ >= 0);
        if (seen[cls]++ % P.world != P.rank) continue;
        by_cls[cls].push_back({b1, b2});
        const long long nq = (long long)x.nI * x.nJ * y.nI * y.nJ;
        tp->n_quartets += nq; tp->n_prim_quartets += nq * x.KK * y.KK;
      }
    }
    for (int c = 0; c < 16; ++c) { tp->cls_start[c] = (int)tp->tiles.size(); tp->tiles.insert(tp->tiles.end(), by_cls[c].begin(), by_cls[c].end()); }
    tp->cls_start[16] = (int)tp->tiles.size();
    constexpr int kChunk = 256;
    for (int c = 0; c < 16; ++c) {
      tp->chunk_cls_start[c] = (int)tp->chunks.size();
      for (int t = tp->cls_start[c]; t < tp->cls_start[c + 1];) {
        int e = t + 1;
        while (e < tp->cls_start[c + 1] && e - t < kChunk && tp->tiles[e].bp1 == tp->tiles[t].bp1) ++e;
        tp->chunks.push_back(make_int2(t, e - t));
        t = e;
      }
    }
    tp->chunk_cls_start[16] = (int)tp->chunks.size();
    // a value tile lives 50-100 us: 16 tiles per CTA amortise the per-chunk staging of the bra records and the D / K strips
    // without leaving long tails
    const int kDirectChunk = getenv("DQ_DIRECT_CHUNK") ? std::max(1, atoi(getenv("DQ_DIRECT_CHUNK"))) : 16;
    for (int c = 0; c < 16; ++c) {
      tp->dchunk_cls_start[c] = (int)tp->dchunks.size();
      for (int t = tp->cls_start[c]; t < tp->cls_start[c + 1];) {
        int e = t + 1;
        while (e < tp->cls_start[c + 1] && e - t < kDirectChunk && tp->tiles[e].bp1 == tp->tiles[t].bp1) ++e;
        tp->dchunks.push_back(make_int2(t, e - t));
        t = e;
      }
    }
    tp->dchunk_cls_start[16] = (int)tp->dchunks.size();
    const int kGradChunk = getenv("DQ_GRAD_CHUNK") ? std::max(1, atoi(getenv("DQ_GRAD_CHUNK"))) : 12;
    tp->gtiles = tp->tiles;
    for (int c = 0; c < 16; ++c) {
      std::stable_sort(tp->gtiles.begin() + tp->cls_start[c], tp->gtiles.begin() + tp->cls_start[c + 1],
                       [](const Tile& a, const Tile& b) { return a.bp2 < b.bp2; });
      tp->gchunk_cls_start[c] = (int)tp->gchunks.size();
      for (int t = tp->cls_start[c]; t < tp->cls_start[c + 1];) {
        int e = t + 1;
        while (e < tp->cls_start[c + 1] && e - t < kGradChunk && tp->gtiles[e].bp2 == tp->gtiles[t].bp2) ++e;
        tp->gchunks.push_back(make_int2(t, e - t));
        t = e;
      }
    }
    tp->gchunk_cls_start[16] = (int)tp->gchunks.size();
    for (int c = 0; c < 16; ++c) {
      tp->gchunk32_cls_start[c] = (int)tp->gchunks32.size();
      for (int t = tp->cls_start[c]; t < tp->cls_start[c + 1];) {
        int e = t + 1;
        while (e < tp->cls_start[c + 1] && e - t < 20 && tp->gtiles[e].bp2 == tp->gtiles[t].bp2) ++e;     // kSparseChunk
        tp->gchunks32.push_back(make_int2(t, e - t));
        t = e;
      }
    }
    tp->gchunk32_cls_start[16] = (int)tp->gchunks32.size();
    if (!P.plan_only) {
      CU(cudaMalloc(&tp->d_dchunks, std::max<size_t>(tp->dchunks.size(), 1) * sizeof(int2)));
      CU(cudaMemcpyAsync(tp->d_dchunks, tp->dchunks.data(), tp->dchunks.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
      CU(cudaMalloc(&tp->d_gchunks32, std::max<size_t>(tp->gchunks32.size(), 1) * sizeof(int2)));
      CU(cudaMemcpyAsync(tp->d_gchunks32, tp->gchunks32.data(), tp->gchunks32.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
      CU(cudaMalloc(&tp->d_gtiles, std::max<size_t>(tp->gtiles.size(), 1) * sizeof(Tile)));
      CU(cudaMalloc(&tp->d_gchunks, std::max<size_t>(tp->gchunks.size(), 1) * sizeof(int2)));
      CU(cudaMemcpyAsync(tp->d_gtiles, tp->gtiles.data(), tp->gtiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(tp->d_gchunks, tp->gchunks.data(), tp->gchunks.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
      CU(cudaMalloc(&tp->d_tiles, std::max<size_t>(tp->tiles.size(), 1) * sizeof(Tile)));
      CU(cudaMalloc(&tp->d_chunks, std::max<size_t>(tp->chunks.size(), 1) * sizeof(int2)));
      CU(cudaMemcpyAsync(tp->d_tiles, tp->tiles.data(), tp->tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(tp->d_chunks, tp->chunks.data(), tp->chunks.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
      CU(cudaStreamSynchronize(st));
      t_plan = tp;
    }
  }
  P.tp = tp;
  P.n_quartets = tp->n_quartets; P.n_prim_quartets = tp->n_prim_quartets;
  P.cls_start = tp->cls_start; P.ntiles = tp->tiles.size();
  if (P.plan_only) return;
  P.d_tiles = tp->d_tiles; P.d_chunks = tp->d_chunks; P.nchunks = (int)tp->chunks.size();
  if (((size_t)16 * N + 2 * 16 * 17 + 16) * sizeof(double) > 220 * 1024) fail(-1, "basis too large for the J/K kernel's shared-memory row strips");
  P.d_pd = ws.alloc<double>(P.pd_len);
  double* d_kmax = ws.alloc<double>(P.nbp);
  P.d_skipped = ws.alloc<unsigned long long>(1, true);
  P.sd.bp_kmax = d_kmax;
  P.sd.screen = (opt && opt->screen_threshold > 0.0) ? opt->screen_threshold : 0.0;
  P.sd.skipped = P.d_skipped;
  P.d_underflowed = ws.alloc<unsigned long long>(1, true);
  P.sd.underflowed = P.d_underflowed;
  launch_pairs(st, P.sd, P.d_pd, d_kmax);
  P.sd.underflowed = nullptr;
  {
    const char* force = getenv("DQ_ERI_KERNEL");
    if (force) P.eri_compact = force[0] == 'c';
    else if (P.sd.zero_skip) {
      unsigned long long uf = 0;
      CU(cudaMemcpyAsync(&uf, P.d_underflowed, sizeof uf, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      long long tot = 0;
      for (const BlockPair& b : P.bps) tot += (long long)b.KK * b.nI * b.nJ;
      P.eri_compact = (double)uf > 0.25 * (double)tot;
      if (uf == 0) P.sd.zero_skip = 0;      // nothing underflowed (compact molecules): the kernels without the skip tests
    }
    // the compacting kernel adds 32 KB of partial-sum cells and two entry lists to the staged records
    const size_t compact_smem = (size_t)2 * P.kkmax * kPairFields * 16 * sizeof(double) + 16 * 256 * sizeof(double) +
                                (size_t)2 * 16 * P.kkmax * sizeof(unsigned short) + 16;
    if (compact_smem > 220 * 1024) P.eri_compact = false;
  }
  if ((size_t)2 * P.kkmax * kPairFields * 16 * sizeof(double) > 200 * 1024) fail(-1, "contraction too long for the ERI kernel's shared-memory staging");
}

// decide stored vs integral-direct (dq_options.jk_mode) and allocate the tile store or the transient batch buffer
static void setup_jk_mode(Prepared& P, const dq_options* opt) {
  const int mode = opt ? opt->jk_mode : 0;
  const size_t store_bytes = (size_t)P.ntiles * kTile * sizeof(double);
  if (mode == 2) P.direct = true;
  else if (mode == 1) P.direct = false;
  else {
    size_t fre = 0, tot = 0;
    CU(cudaMemGetInfo(&fre, &tot));
    P.direct = store_bytes > (fre + g_pool.cached_bytes) / 10 * 7;     // keep 30 % headroom for the SCF history
  }
  if (P.direct && jk_direct_smem_bytes(P.kkmax, P.N, P.sd.det) > 220 * 1024)
    fail(-1, "basis / contraction too large for the integral-direct kernel's shared-memory staging");
  t_stats.jk_direct = P.direct ? 1 : 0;
}

// The ERI pass re-reads the primitive-pair records (15 MB for (H2O)_32) once per tile while it streams 2.6 GB of tiles out.
// The tiles go out with evict-first stores, but how much of the record set they still push out of L2 depends on where the
// two buffers landed physically: the pass took 230 ms in three processes of four and 260-285 ms in the fourth.  For the
// duration of the pass the records are therefore pinned with an L2 access-policy window (persisting hits) on its stream.
struct L2Pin {
  cudaStream_t st; bool on = false;
  L2Pin(cudaStream_t s, const void* base, size_t bytes) : st(s) {
    static thread_local int limit_dev = -1;
    int dev = 0; cudaGetDevice(&dev);
    int maxwin = 0, maxpersist = 0;
    cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceGetAttribute(&maxpersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    if (getenv("DQ_NO_L2_PIN") || maxwin <= 0 || maxpersist <= 0 || bytes == 0) return;
    if (limit_dev != dev) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)std::min<long long>(maxpersist, 48ll << 20)); limit_dev = dev; }
    cudaStreamAttrValue a = {};
    a.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    a.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)maxwin);
    a.accessPolicyWindow.hitRatio = 1.0f;
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    on = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) == cudaSuccess;
    if (!on) cudaGetLastError();
  }
  ~L2Pin() {
    if (!on) return;
    cudaStreamAttrValue a = {};
    a.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
  }
};

static void compute_eri_store(Prepared& P) {
  if (P.direct) return;
  P.d_store = P.ws.alloc<double>((size_t)P.ntiles * kTile);
  L2Pin pin(P.ws.st, P.d_pd, (size_t)P.pd_len * sizeof(double));
  // one stream: a value tile lives ~45 us, so the tails of the class launches are negligible, and concurrent class kernels
  // made the pass time erratic (129-148 ms from run to run)
  for (int c = 15; c >= 0; --c) {        // p-rich (long) classes first
    if (P.cls_start[c + 1] == P.cls_start[c]) continue;
    if (P.eri_compact)
      launch_eri_tiles_compact(P.ws.st, c, P.d_tiles + P.cls_start[c], P.cls_start[c + 1] - P.cls_start[c], P.sd, P.d_pd,
                               P.d_store + (size_t)P.cls_start[c] * kTile, P.kkmax);
    else
      launch_eri_tiles(P.ws.st, c, P.d_tiles + P.cls_start[c], P.cls_start[c + 1] - P.cls_start[c], P.sd, P.d_pd,
                       P.d_store + (size_t)P.cls_start[c] * kTile, P.kkmax);
  }
  t_stats.eri_kernel_compact = P.eri_compact ? 1 : 0;
}

// unsort an N x N matrix (rows and/or columns) into the caller's ordering
static void unsort_matrix(const Prepared& P, const std::vector<double>& in, double* out, bool rows, bool cols) {
  const int n = P.N;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      out[(rows ? P.f_orig[i] : i) * n + (cols ? P.f_orig[j] : j)] = in[i * n + j];
}

// G[D] = 2J - K from the stored tiles (+ allreduce of the partial J|K over ranks); F = H + G if F != null
struct JK {
  double *JKacc;   // [2*n*n] J accumulators then K accumulators (one buffer: one allreduce)
  double *JKraw = nullptr;   // fixed-point accumulation: [2*2*n*n] limb pairs the kernel adds into, converted into JKacc
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;   // brackets of the contraction kernels, read by ms() after a sync
  double ms() {
    double tot = 0.0;
    for (auto& e : ev) { float m = 0; if (cudaEventElapsedTime(&m, e.first, e.second) == cudaSuccess) tot += m; }
    return tot;
  }
  ~JK() { for (auto& e : ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } }
};
static void fock_build(Prepared& P, JK& jk, const double* D, const double* H, double* F, double* G) {
  const int n = P.N;
  cudaStream_t st = P.ws.st;
  const bool det = P.sd.det != 0;
  double* acc = det ? jk.JKraw : jk.JKacc;                       // what the contraction kernel adds into
  const size_t kofs = (size_t)n * n * (det ? 2 : 1);             // K accumulators follow the J accumulators
  CU(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * kofs, st));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  jk.ev.push_back({e0, e1});
  CU(cudaEventRecord(e0, st));
  if (!P.direct) {
    launch_jk_chunks(st, P.d_tiles, P.d_chunks, P.nchunks, P.sd, P.d_store, D, acc, acc + kofs);
  } else {
    // integral-direct: one fused evaluate-and-contract kernel per angular class (p-rich, long classes first), the class
    // launches on side streams so that their tails overlap
    const TilePlan& tp = *P.tp;
    t_fork.init(P.ws.dev);
    t_fork.begin(st);
    for (int c = 15, lane = 0; c >= 0; --c) {
      const int nch = tp.dchunk_cls_start[c + 1] - tp.dchunk_cls_start[c];
      if (nch <= 0) continue;
      launch_jk_direct(t_fork.lane(lane++), c, P.d_tiles, tp.d_dchunks + tp.dchunk_cls_start[c], nch, P.sd, P.d_pd, D, acc, acc + kofs,
                       P.kkmax);
    }
    t_fork.end(st);
  }
  CU(cudaEventRecord(e1, st));
  if (det) launch_det_finalize(st, (long long)2 * n * n, jk.JKraw, jk.JKacc, 0.0);
  if (P.world > 1) P.allreduce(jk.JKacc, (int64_t)2 * n * n, (void*)st, P.ar_user);
  launch_fock_combine(st, n, P.d_fblock, H, jk.JKacc, jk.JKacc + (size_t)n * n, F, G);
  t_stats.fock_builds++;
}

static double nuclear_repulsion(const dq_system* s) {   // Energy.py:21-27 (O(natoms^2) scalar work on the host)
  double e = 0.0;
  for (int i = 0; i < s->natoms; ++i)
    for (int j = i + 1; j < s->natoms; ++j) {
      double r2 = 0;
      for (int c = 0; c < 3; ++c) { const double d = s->atoms[3 * j + c] - s->atoms[3 * i + c]; r2 += d * d; }
      e = e + s->charges[i] * s->charges[j] / sqrt(r2);
    }
  return e;
}

// ---------------------------------------------------------------------------------------------
// SCF (+ optional gradient)
// ---------------------------------------------------------------------------------------------
struct ScfOut { double E = 0; int niter = 0; bool converged = false; std::vector<double> esteps, C, eps; };

static void run_rhf(const dq_system* s, const dq_options* opt, ScfOut& out, bool want_grad, double* g_alpha, double* g_coef,
                    double* g_xyz, double* g_atoms = nullptr) {
  memset(&t_stats, 0, sizeof t_stats);
  reset_launch_count();
  Prepared P;
  if (opt && opt->max_scf < 1) fail(-1, "max_scf must be >= 1 (Energy.py:164: the SCF loop needs at least one iteration)");
  const int max_scf = opt ? opt->max_scf : 30;
  const double tol = (opt && opt->e_tol > 0) ? opt->e_tol : 1e-8;
  if (s->ne <= 0 || s->ne > s->nbasis) fail(-1, "ne (doubly-occupied orbitals) must be in 1..nbasis");
  cudaStream_t st0 = opt ? (cudaStream_t)opt->stream : 0;
  need_device(opt ? opt->device : 0);
  Timer t_total(st0);
  Timer t_setup(st0);
  prepare(P, s, opt, true, true);
  Workspace& ws = P.ws;
  cudaStream_t st = ws.st;
  const int n = P.N, ne = P.ne;
  const size_t nn = (size_t)n * n;
  t_stats.ms_setup = t_setup.stop();
  t_stats.n_tiles = (int64_t)P.ntiles; t_stats.n_quartets = P.n_quartets; t_stats.n_prim_quartets = P.n_prim_quartets;

  Timer t_1e(st);
  double *S = ws.alloc<double>(nn), *T = ws.alloc<double>(nn), *V = ws.alloc<double>(nn), *H = ws.alloc<double>(nn);
  launch_one_electron(st, P.sd, S, T, V);
  launch_axpby(st, (int)nn, 1.0, T, 0.0, H);
  launch_axpby(st, (int)nn, 1.0, V, 1.0, H);                                  // Energy.py:137
  t_stats.ms_one_electron = t_1e.stop();
  check_launches("setup / one-electron kernels");

  // SCF buffers
  double *L = ws.alloc<double>(nn), *wS = ws.alloc<double>(n), *X = ws.alloc<double>(nn), *t1 = ws.alloc<double>(nn),
         *t2 = ws.alloc<double>(nn), *t3 = ws.alloc<double>(nn), *vwork = ws.alloc<double>(nn);
  void* ework = ws.alloc<char>(eigh_workspace_bytes(n, 1));
  JK jk; jk.JKacc = ws.alloc<double>(2 * nn);
  if (P.sd.det) jk.JKraw = ws.alloc<double>(4 * nn);
  double *F = ws.alloc<double>(nn), *Fp = ws.alloc<double>(nn), *U = ws.alloc<double>(nn), *C = ws.alloc<double>(nn),
         *D = ws.alloc<double>(nn), *eps = ws.alloc<double>(n), *dE = ws.alloc<double>(1);
  // history for the reverse sweep: F_{t-1}, U_t, eps_t, D_t, F_t  (device slabs of 16 iterations: no per-iteration cudaMalloc)
  std::vector<double*> hFprev, hU, hEps, hD, hF;
  double* slab = nullptr; int slab_left = 0;
  auto hist = [&](size_t cnt) { double* p = slab; slab += cnt; return p; };

  // everything of an SCF iteration that precedes the Fock build: F' = X^T F X, eigensolve, C = X U, D = C_occ C_occ^T
  auto scf_pre = [&](int it, cudaStream_t sx) {
    if (want_grad) {
      if (slab_left == 0) { slab = ws.alloc<double>((size_t)16 * (4 * nn + n)); slab_left = 16; }
      --slab_left;
      double* f = hist(nn); CU(cudaMemcpyAsync(f, F, nn * sizeof(double), cudaMemcpyDeviceToDevice, sx)); hFprev.push_back(f);
    }
    launch_gemm(sx, true, false, n, n, n, X, n, F, n, t1, n, 1.0, 0.0);       // Energy.py:166 (the duplicated block
    launch_gemm(sx, false, false, n, n, n, t1, n, X, n, Fp, n, 1.0, 0.0);     //   :170-173 recomputes the same C)
    // Energy.py:167-168.  Warm start: successive Fock matrices are close.  Cluster solver: the previous eigenvectors are
    // the starting basis (2-4 sweeps instead of 7).  Single-CTA solver: F' is first rotated into the previous eigenbasis
    // (nearly diagonal -> 2-4 sweeps instead of ~9) and U = U_prev V.
    if (it == 0 || getenv("DQ_COLD_EIGH")) {
      launch_eigh(sx, n, 1, Fp, vwork, U, eps, ework);
    } else if (eigh_cluster_applicable(n)) {
      // cluster solver: the previous eigenvectors go in as the starting basis, the new ones come out directly
      launch_symmetrize(sx, n, Fp);
      launch_eigh_cluster(sx, n, 1, Fp, U, t3, eps, ework);
      CU(cudaMemcpyAsync(U, t3, nn * sizeof(double), cudaMemcpyDeviceToDevice, sx));
    } else {
      launch_gemm(sx, true, false, n, n, n, U, n, Fp, n, t1, n, 1.0, 0.0);      // U_prev^T F'
      launch_gemm(sx, false, false, n, n, n, t1, n, U, n, t2, n, 1.0, 0.0);     // U_prev^T F' U_prev
      launch_symmetrize(sx, n, t2);
      launch_eigh(sx, n, 1, t2, vwork, t3, eps, ework);
      launch_gemm(sx, false, false, n, n, n, U, n, t3, n, t1, n, 1.0, 0.0);     // U_prev V
      CU(cudaMemcpyAsync(U, t1, nn * sizeof(double), cudaMemcpyDeviceToDevice, sx));
    }
    launch_gemm(sx, false, false, n, n, n, X, n, U, n, C, n, 1.0, 0.0);       // Energy.py:169
    launch_gemm(sx, false, true, n, n, ne, C, n, C, n, D, n, 1.0, 0.0);       // Energy.py:174-180
  };

  // S^-1/2 and the whole first iteration up to its Fock build depend only on S and H: they run on a side stream,
  // concurrently with the ERI pass (single-CTA eigensolver kernels next to 1.27 M tile CTAs).
  cudaStream_t side = nullptr;
  cudaEvent_t ev_1e = nullptr, ev_side = nullptr;
  {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
    CU(cudaEventCreateWithFlags(&ev_1e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
  }
  struct SideGuard { cudaStream_t s; cudaEvent_t a, b; ~SideGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); cudaEventDestroy(a); cudaEventDestroy(b); } } side_guard{side, ev_1e, ev_side};
  CU(cudaEventRecord(ev_1e, st));
  CU(cudaStreamWaitEvent(side, ev_1e, 0));
  // S^-1/2 = L diag(w^-1/2) L^T                                               // Energy.py:138-144
  CU(cudaMemcpyAsync(t1, S, nn * sizeof(double), cudaMemcpyDeviceToDevice, side));
  launch_eigh(side, n, 1, t1, vwork, L, wS, ework);
  launch_scale_cols_rsqrt(side, n, L, wS, t2);
  launch_gemm(side, false, true, n, n, n, t2, n, L, n, X, n, 1.0, 0.0);
  // readguess (Energy.py:148-157): D0 = C0_occ C0_occ^T from the caller's coefficients, first Fock matrix H + G[D0]
  const bool have_guess = opt && opt->guess_C != nullptr;
  double* D0 = nullptr;
  if (have_guess) {
    std::vector<double> c0(nn);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) c0[(size_t)i * n + j] = opt->guess_C[(size_t)P.f_orig[i] * n + j];
    double* dC0 = ws.upload(c0);
    CU(cudaStreamSynchronize(st));      // c0 is a stack-scoped staging vector
    D0 = ws.alloc<double>(nn);
    launch_gemm(side, false, true, n, n, ne, dC0, n, dC0, n, D0, n, 1.0, 0.0);
  } else {
    CU(cudaMemcpyAsync(F, H, nn * sizeof(double), cudaMemcpyDeviceToDevice, side));      // Energy.py:159 (core guess)
    scf_pre(0, side);
  }
  CU(cudaEventRecord(ev_side, side));

  // The cluster eigensolver (8 CTAs x 196 KB of shared memory) does not co-schedule well with the tile kernels: overlapped,
  // the two cold solves (~4 ms) cost the ERI pass 13 ms; they run first instead.  The single-CTA solver (DQ_EIGH_PACKED)
  // overlaps for free.
  if (eigh_cluster_applicable(n) && !getenv("DQ_SIDE_OVERLAP")) CU(cudaStreamWaitEvent(st, ev_side, 0));
  Timer t_eri(st);
  setup_jk_mode(P, opt);
  compute_eri_store(P);                                                        // Energy.py:136
  t_stats.ms_eri = t_eri.stop();
  check_launches("ERI tile kernels");

  Timer t_scf(st);
  CU(cudaStreamWaitEvent(st, ev_side, 0));
  if (have_guess) {                      // the first Fock matrix needs the ERIs: no overlap for iteration 0
    fock_build(P, jk, D0, H, F, nullptr);                                      // Energy.py:157
    scf_pre(0, st);
  }
  double oldE = 1e8, e_elec = 0.0;
  out.esteps.clear();
  out.converged = false;
  const bool trace = getenv("DQ_TRACE") != nullptr;
  for (int it = 0; it < max_scf; ++it) {
    double tr0 = 0, tr1 = 0, tr2 = 0, tr3 = 0;
    auto now = [&]() { cudaStreamSynchronize(st); timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    if (trace) tr0 = tr1 = now();
    if (it > 0) scf_pre(it, st);
    if (trace) tr2 = now();
    fock_build(P, jk, D, H, F, nullptr);                                       // Energy.py:188
    launch_dot3(st, (int)nn, D, H, F, dE);                                     // Energy.py:189
    CU(cudaMemcpyAsync(&e_elec, dE, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (trace) { tr3 = now(); fprintf(stderr, "[dq trace] it %d: pre %.3f ms, eigh %.3f ms, rest %.3f ms\n", it, tr1 - tr0, tr2 - tr1, tr3 - tr2); }
    out.esteps.push_back(e_elec);
    out.niter = it + 1;
    if (want_grad) {
      double *u = hist(nn), *e = hist(n), *d = hist(nn), *f = hist(nn);
      CU(cudaMemcpyAsync(u, U, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CU(cudaMemcpyAsync(e, eps, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CU(cudaMemcpyAsync(d, D, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CU(cudaMemcpyAsync(f, F, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
      hU.push_back(u); hEps.push_back(e); hD.push_back(d); hF.push_back(f);
    }
    if (fabs(e_elec - oldE) < tol) { out.converged = true; break; }          // Energy.py:192-194
    oldE = e_elec;
  }
  out.E = e_elec + nuclear_repulsion(s);                                       // Energy.py:196, 316/324
  {
    std::vector<double> hc(nn), he(n);
    CU(cudaMemcpyAsync(hc.data(), C, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(he.data(), eps, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out.C.resize(nn); out.eps = he;
    unsort_matrix(P, hc, out.C.data(), true, false);
  }
  t_stats.ms_scf = t_scf.stop();
  t_stats.scf_iterations = out.niter;
  check_launches("SCF kernels");

  if (want_grad) {
    // ---- exact reverse sweep of the k iterations actually performed
    Timer t_rev(st);
    const int k = out.niter;
    double *Dbar = ws.alloc<double>(nn), *Hbar = ws.alloc<double>(nn), *Xbar = ws.alloc<double>(nn, true), *Pm = ws.alloc<double>(nn),
           *Fbar = ws.alloc<double>(nn), *Y = ws.alloc<double>(nn);
    double* At = ws.alloc<double>((size_t)(k + 1) * nn);
    double* Bt = ws.alloc<double>((size_t)(k + 1) * nn);
    int nterms = 0;
    launch_axpby(st, (int)nn, 2.0, hF[k - 1], 0.0, Dbar);
    launch_axpby(st, (int)nn, 2.0, hD[k - 1], 0.0, Hbar);
    CU(cudaMemcpyAsync(At, hD[k - 1], nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(Bt, hD[k - 1], nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
    nterms = 1;
    for (int t = k; t >= 1; --t) {
      double* Ut = hU[t - 1];
      launch_gemm(st, false, true, n, n, ne, Ut, n, Ut, n, Pm, n, 1.0, 0.0);      // P_t = U_occ U_occ^T
      launch_gemm(st, false, false, n, n, n, Dbar, n, X, n, t1, n, 1.0, 0.0);     // t1 = Dbar X
      launch_gemm(st, false, false, n, n, n, t1, n, Pm, n, t2, n, 1.0, 0.0);      // t2 = Dbar X P
      launch_add_transpose(st, n, t2, Xbar);                                       // Xbar += t2 + t2^T
      launch_gemm(st, false, false, n, n, n, X, n, t1, n, t2, n, 1.0, 0.0);       // Pbar = X Dbar X
      launch_gemm(st, true, false, n, n, n, Ut, n, t2, n, t1, n, 1.0, 0.0);       // U^T Pbar
      launch_gemm(st, false, false, n, n, n, t1, n, Ut, n, t2, n, 1.0, 0.0);      // Z = U^T Pbar U
      launch_make_y(st, n, ne, t2, hEps[t - 1], Y);
      launch_gemm(st, false, false, n, n, n, Ut, n, Y, n, t1, n, 1.0, 0.0);       // U Y
      launch_gemm(st, false, true, n, n, n, t1, n, Ut, n, t2, n, 1.0, 0.0);       // Abar = U Y U^T
      launch_gemm(st, false, false, n, n, n, t2, n, X, n, t1, n, 1.0, 0.0);       // t1 = Abar X
      launch_gemm(st, false, false, n, n, n, t1, n, hFprev[t - 1], n, t3, n, 1.0, 0.0);   // Abar X F_{t-1}
      launch_add_transpose(st, n, t3, Xbar);
      launch_gemm(st, false, false, n, n, n, X, n, t1, n, Fbar, n, 1.0, 0.0);     // Fbar = X Abar X
      launch_symmetrize(st, n, Fbar);
      launch_axpby(st, (int)nn, 1.0, Fbar, 1.0, Hbar);
      if (t > 1) {
        CU(cudaMemcpyAsync(At + (size_t)nterms * nn, Fbar, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(Bt + (size_t)nterms * nn, hD[t - 2], nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        ++nterms;
        fock_build(P, jk, Fbar, nullptr, nullptr, Dbar);                           // Dbar_{t-1} = G[Fbar]
      } else if (have_guess) {             // F_0 = H + G[D0] with a parameter-independent D0: one more two-electron term
        CU(cudaMemcpyAsync(At + (size_t)nterms * nn, Fbar, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(Bt + (size_t)nterms * nn, D0, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        ++nterms;
      }
    }
    // Sbar = L [ (L^T Xbar L) o Phi ] L^T
    double* Sbar = ws.alloc<double>(nn);
    launch_gemm(st, true, false, n, n, n, L, n, Xbar, n, t1, n, 1.0, 0.0);
    launch_gemm(st, false, false, n, n, n, t1, n, L, n, t2, n, 1.0, 0.0);
    launch_phi_rsqrt(st, n, wS, t2);
    launch_gemm(st, false, false, n, n, n, L, n, t2, n, t1, n, 1.0, 0.0);
    launch_gemm(st, false, true, n, n, n, t1, n, L, n, Sbar, n, 1.0, 0.0);
    t_stats.ms_reverse = t_rev.stop();
    check_launches("reverse-sweep kernels");

    Timer t_g(st);
    if (eri_grad_smem_bytes(P.kkmax) > 220 * 1024) fail(-1, "contraction too long for the ERI-gradient kernel's shared-memory accumulators");
    const long long adj_len = P.pd_len / kPairFields * kAdjFields;
    const bool det = P.sd.det != 0;
    const int lm = det ? 2 : 1;                  // fixed-point accumulation: every accumulator is a pair of integer limbs
    double* padj = ws.alloc<double>(adj_len * lm, true);
    double* pasum = ws.alloc<double>((size_t)P.nbp * 2 * 16 * lm, true);
    // gbuf = [g_alpha | g_coef(normalised) | g_xyz]  (one buffer: one allreduce); gacc: what the kernels add into
    const long long glen = (long long)2 * P.nprim + 3 * n;
    double* gbuf = ws.alloc<double>(glen, true);
    double* gacc = det ? ws.alloc<double>(2 * glen, true) : gbuf;
    double *ga = gbuf, *gc = gbuf + P.nprim, *gx = gbuf + 2 * P.nprim;
    double *aa = gacc, *ac = gacc + P.nprim, *ax = gacc + 2 * P.nprim;      // element windows (the kernels index from aa)
    unsigned long long* d_zero = ws.alloc<unsigned long long>(1, true);
    P.sd.zero_weight = d_zero;
    Timer t_gk(st);
    // dense or sparse weights?  The density matrix of well separated fragments is exactly zero between them (the cluster
    // eigensolver keeps exact zeros), and then most quartets carry no weight: the lane-independent kernel pays off.
    bool sparse = false;
    {
      const char* force = getenv("DQ_GRAD_KERNEL");
      if (force) sparse = force[0] == 's';
      else {
        int* d_nz = ws.alloc<int>(1);
        int h_nz = 0;
        launch_count_nonzero(st, (int)nn, hD[k - 1], d_nz);
        CU(cudaMemcpyAsync(&h_nz, d_nz, sizeof h_nz, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        sparse = (double)h_nz < 0.25 * (double)nn;
      }
    }
    t_stats.grad_kernel_sparse = sparse ? 1 : 0;
    if (sparse) {
      unsigned char* d_mask = ws.alloc<unsigned char>(nn);
      launch_weight_mask(st, (int)nn, nterms, Bt, d_mask);
      P.sd.wmask = d_mask;
    }
    // deferred loop-regime quartets (dense weights, short contractions): masks of the marked primitive quartets, 3 words per
    // thread and tile
    unsigned* dmask = nullptr;
    unsigned char* dbin = nullptr;
    size_t def_off[16] = {0};
    const char* defer_env = getenv("DQ_GRAD_DEFER");      // default on; DQ_GRAD_DEFER=0: Boys loops in place in every class
    if (!sparse && !(defer_env && defer_env[0] == '0') && P.kkmax * P.kkmax <= 96) {
      size_t ndef = 0;      // tiles of the classes that defer (>= 2 p functions): buffers are laid out class by class
      for (int c = 0; c < 16; ++c) {
        def_off[c] = ndef;
        if (((c >> 3) & 1) + ((c >> 2) & 1) + ((c >> 1) & 1) + (c & 1) >= 2) ndef += (size_t)(P.cls_start[c + 1] - P.cls_start[c]);
      }
      // 27 KB per deferring tile: fine for (H2O)_32 (17 GB of 180), not for much larger clusters - then the loops stay in place
      size_t free_b = 0, total_b = 0;
      if (ndef > 0 && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && ndef * (size_t)(96 + 12) * 256 > free_b / 3) ndef = 0;
      if (ndef > 0) {
        dmask = ws.alloc<unsigned>(ndef * 3 * 256);
        CU(cudaMemsetAsync(dmask, 0, ndef * 3 * 256 * sizeof(unsigned), st));
        dbin = ws.alloc<unsigned char>(ndef * 96 * 256);      // one byte per (tile, step, thread): only marked entries are written and read
      }
    }
    t_fork.init(ws.dev);
    t_fork.begin(st);
    for (int c = 15, lane = 0; c >= 0; --c) {
      if (sparse) {
        const int nch = P.tp->gchunk32_cls_start[c + 1] - P.tp->gchunk32_cls_start[c];
        if (nch <= 0) continue;
        launch_eri_grad_sparse(t_fork.lane(lane++), c, P.tp->d_gtiles, P.tp->d_gchunks32 + P.tp->gchunk32_cls_start[c], nch, P.sd,
                               P.d_pd, nterms, At, Bt, padj, pasum, P.kkmax);
        continue;
      }
      const int nch = P.tp->gchunk_cls_start[c + 1] - P.tp->gchunk_cls_start[c];
      if (nch <= 0) continue;
      const int npf = ((c >> 3) & 1) + ((c >> 2) & 1) + ((c >> 1) & 1) + (c & 1);      // p functions of the class
      unsigned* dm = (dmask && npf >= 2) ? dmask + def_off[c] * 3 * 256 : nullptr;
      unsigned char* db = dm ? dbin + def_off[c] * 96 * 256 : nullptr;
      cudaStream_t ls = t_fork.lane(lane++);
      launch_eri_grad_chunks(ls, c, P.tp->d_gtiles, P.tp->d_gchunks + P.tp->gchunk_cls_start[c], nch, P.sd, P.d_pd,
                             nterms, At, Bt, padj, pasum, P.kkmax, dm, db, P.cls_start[c]);
      if (dm)      // the marked loop-regime quartets of this class (same stream: after its first kernel)
        launch_eri_grad_deferred(ls, c, P.tp->d_gtiles, P.cls_start[c], P.cls_start[c + 1] - P.cls_start[c], P.sd, P.d_pd, nterms, At, Bt,
                                 padj, pasum, dm, db, P.kkmax);
    }
    t_fork.end(st);
    t_stats.ms_eri_grad = t_gk.stop();
    check_launches("ERI-gradient kernels");
    launch_pair_bwd(st, P.sd, padj, pasum, aa, ac, ax);
    if (det) launch_det_finalize(st, glen, gacc, gbuf, 0.0);
    if (P.world > 1) P.allreduce(gbuf, (int64_t)glen, (void*)st, P.ar_user);
    // the one-electron part is replicated (every rank adds the same numbers after the allreduce)
    if (det) CU(cudaMemsetAsync(gacc, 0, sizeof(double) * 2 * glen, st));
    launch_one_electron_grad(st, P.sd, P.kmax, Sbar, Hbar, aa, ac, ax);
    if (det) launch_det_finalize(st, glen, gacc, gbuf, 1.0);
    // nuclear coordinates (Energy.py:125-130: only V and E_nuc depend on them; the Gaussians do not follow the nuclei)
    double* natom_buf = nullptr;
    if (g_atoms && s->natoms > 0) {
      const long long na = 3LL * s->natoms;
      double* nacc = ws.alloc<double>(na * lm, true);
      launch_nuclear_grad(st, P.sd, P.kmax, Hbar, nacc);
      natom_buf = nacc;
      if (det) { natom_buf = ws.alloc<double>(na); launch_det_finalize(st, na, nacc, natom_buf, 0.0); }
    }
    double* graw = ws.alloc<double>(P.nprim);
    launch_normalize_bwd(st, n, P.d_foff, P.d_ftype, P.d_alpha, P.d_craw, gc, ga, graw);
    if (P.log_alpha) launch_mul(st, P.nprim, ga, P.d_alpha, ga);                   // d/d ln(alpha) = alpha d/d alpha
    unsigned long long h_zero = 0;
    CU(cudaMemcpyAsync(&h_zero, d_zero, sizeof h_zero, cudaMemcpyDeviceToHost, st));
    std::vector<double> ha(P.nprim), hc(P.nprim), hx(3 * n);
    CU(cudaMemcpyAsync(ha.data(), ga, P.nprim * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hc.data(), graw, P.nprim * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hx.data(), gx, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (natom_buf) CU(cudaMemcpyAsync(g_atoms, natom_buf, 3 * s->natoms * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (natom_buf) {      // + dE_nuc/dR_i = sum_j Z_i Z_j (R_j - R_i) / |R_i - R_j|^3          (Energy.py:21-27)
      for (int i = 0; i < s->natoms; ++i)
        for (int j = 0; j < s->natoms; ++j) {
          if (j == i) continue;
          double d[3], r2 = 0;
          for (int c = 0; c < 3; ++c) { d[c] = s->atoms[3 * j + c] - s->atoms[3 * i + c]; r2 += d[c] * d[c]; }
          const double f = s->charges[i] * s->charges[j] / (r2 * sqrt(r2));
          for (int c = 0; c < 3; ++c) g_atoms[3 * i + c] += f * d[c];
        }
    }
    for (int p = 0; p < P.nprim; ++p) { if (g_alpha) g_alpha[P.p_orig[p]] = ha[p]; if (g_coef) g_coef[P.p_orig[p]] = hc[p]; }
    if (g_xyz) for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) g_xyz[3 * P.f_orig[i] + c] = hx[3 * i + c];
    t_stats.ms_grad_other = t_g.stop() - t_stats.ms_eri_grad;
    t_stats.grad_terms = nterms;
    t_stats.n_grad_quartets_zero_weight = (int64_t)h_zero;
  }
  if (P.sd.screen > 0.0 && !P.direct) {
    unsigned long long sk = 0;
    CU(cudaMemcpyAsync(&sk, P.d_skipped, sizeof sk, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    t_stats.n_tiles_skipped = (int64_t)sk;
  }
  {
    unsigned long long uf = 0;
    CU(cudaMemcpyAsync(&uf, P.d_underflowed, sizeof uf, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    t_stats.n_prim_pairs_underflowed = (int64_t)uf;
    long long tot = 0;
    for (const BlockPair& b : P.bps) tot += (long long)b.KK * b.nI * b.nJ;
    t_stats.n_prim_pairs = tot;
  }
  CU(cudaStreamSynchronize(st));
  check_launches("gradient assembly");
  check_boys(st);
  t_stats.ms_jk = jk.ms();
  t_stats.ms_total = t_total.stop();
  t_stats.launches = launch_count();
}

// ---------------------------------------------------------------------------------------------
// batch of molecules of one structure (dq_batch.cu): host side
// ---------------------------------------------------------------------------------------------
static void run_batch(const dq_system* shape, int nmol, const double* atoms, const double* alpha, const double* coef, const double* xyz,
                      const dq_options* opt, bool want_grad, double* E, int32_t* status, int32_t* niter, double* g_alpha,
                      double* g_coef, double* g_xyz) {
  memset(&t_stats, 0, sizeof t_stats);
  reset_batch_launch_count();
  check_system(shape);
  if (nmol <= 0) fail(-1, "nmol must be positive");
  const int n = shape->nbasis, nprim = shape->nprim, natoms = shape->natoms, ne = shape->ne;
  if (ne <= 0 || ne > n) fail(-1, "ne (doubly-occupied orbitals) must be in 1..nbasis");
  if (n > 40) fail(-4, "the batched small-molecule path holds the SCF in shared memory: nbasis <= 40 (use the per-molecule entry points)");
  const int dev = opt ? opt->device : 0;
  need_device(dev);
  if (opt && opt->max_scf < 1) fail(-1, "max_scf must be >= 1 (Energy.py:164: the SCF loop needs at least one iteration)");
  const int max_scf = opt ? opt->max_scf : 30;
  const double tol = (opt && opt->e_tol > 0) ? opt->e_tol : 1e-8;
  cudaStream_t st = opt ? (cudaStream_t)opt->stream : 0;
  static thread_local int boys_dev = -1;
  if (boys_dev != dev) { if (upload_boys_constants_batch(boys_fail_flag()) != 0) fail(-2, "constant upload failed"); boys_dev = dev; }
  int det = (opt && opt->accumulate == 2) ? 0 : 1;
  if (const char* acc = getenv("DQ_ACCUM")) det = acc[1] == 'i' ? 1 : 0;
  Timer t_total(st);

  // structure tables
  const int nn = n * n, M = n * (n + 1) / 2, neri = M * (M + 1) / 2;
  std::vector<int> foff(n + 1, 0), ftype(n), Z(natoms > 0 ? natoms : 1, 0);
  for (int i = 0; i < n; ++i) { foff[i + 1] = foff[i] + shape->contr[i]; ftype[i] = shape->ltype[i]; }
  for (int a = 0; a < natoms; ++a) Z[a] = shape->charges[a];
  std::vector<int2> pairs(M);
  std::vector<int> poff(M + 1, 0), recp;
  int kkmax = 0;
  {
    int x = 0;
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j, ++x) {
        pairs[x] = make_int2(i, j);
        const int kk = shape->contr[i] * shape->contr[j];
        poff[x + 1] = poff[x] + kk;
        kkmax = std::max(kkmax, kk);
        for (int q = 0; q < kk; ++q) recp.push_back(x);
      }
  }
  const int nrec = poff[M];
  if (kkmax > 40) fail(-4, "the batched path keeps ket adjoints in shared memory: at most 40 primitive pairs per function pair");
  std::vector<std::vector<int2>> qcls(16);
  long long npq = 0;
  for (int x = 0; x < M; ++x)
    for (int y = x; y < M; ++y) {
      const int cls = (ftype[pairs[x].x] > 0) * 8 + (ftype[pairs[x].y] > 0) * 4 + (ftype[pairs[y].x] > 0) * 2 + (ftype[pairs[y].y] > 0);
      qcls[cls].push_back(make_int2(x, y));
      npq += (long long)(poff[x + 1] - poff[x]) * (poff[y + 1] - poff[y]);
    }
  std::vector<int2> qlist; int qstart[17];
  for (int c = 0; c < 16; ++c) { qstart[c] = (int)qlist.size(); qlist.insert(qlist.end(), qcls[c].begin(), qcls[c].end()); }
  qstart[16] = (int)qlist.size();

  // molecules are processed in slabs so that the SCF history of a slab stays within a fixed budget
  const size_t per_mol = sizeof(double) * ((size_t)nrec * (kPairFields + kAdjFields) + (size_t)neri + (size_t)8 * nn +
                                           (want_grad ? (size_t)max_scf * (4 * nn + n) + (size_t)(max_scf + 1) * 2 * nn : 0) + 4 * nprim);
  size_t fre = 0, tot = 0;
  CU(cudaMemGetInfo(&fre, &tot));
  const size_t budget = std::min<size_t>((fre + g_pool.cached_bytes) / 2, (size_t)16 << 30);
  const int slab = (int)std::max<size_t>(1, std::min<size_t>((size_t)nmol, budget / std::max<size_t>(per_mol, 1)));

  Workspace ws; ws.dev = dev; ws.st = st;
  int* d_foff = ws.upload(foff); int* d_ftype = ws.upload(ftype); int* d_Z = ws.upload(Z);
  int2* d_pairs = ws.upload(pairs); int* d_poff = ws.upload(poff); int* d_recp = ws.upload(recp);
  int2* d_qlist = ws.upload(qlist);
  CU(cudaStreamSynchronize(st));           // the staging vectors above live on this stack frame
  const size_t hl = (size_t)max_scf * (4 * nn + n);
  for (int m0 = 0; m0 < nmol; m0 += slab) {
    const int nm = std::min(slab, nmol - m0);
    Workspace w2; w2.dev = dev; w2.st = st;
    auto up = [&](const double* h, size_t cnt) {
      double* d = w2.alloc<double>(cnt);
      CU(cudaMemcpyAsync(d, h, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
      return d;
    };
    BatchDev b;
    b.det = det;
    b.nmol = nm; b.N = n; b.nprim = nprim; b.natoms = natoms; b.M = M; b.nrec = nrec; b.neri = neri; b.kkmax = kkmax; b.ne = ne;
    b.f_off = d_foff; b.f_type = d_ftype; b.Z = d_Z; b.pairs = d_pairs; b.pair_off = d_poff; b.rec_pair = d_recp;
    b.atoms = up(atoms + (size_t)m0 * natoms * 3, (size_t)nm * std::max(natoms, 1) * 3);
    b.xyz = up(xyz + (size_t)m0 * n * 3, (size_t)nm * n * 3);
    b.craw = up(coef + (size_t)m0 * nprim, (size_t)nm * nprim);
    const double* d_alpha_in = up(alpha + (size_t)m0 * nprim, (size_t)nm * nprim);
    b.alpha = w2.alloc<double>((size_t)nm * nprim); b.cn = w2.alloc<double>((size_t)nm * nprim);
    b.prec = w2.alloc<double>((size_t)nrec * kPairFields * nm);
    double *S = w2.alloc<double>((size_t)nm * nn), *H = w2.alloc<double>((size_t)nm * nn), *eri = w2.alloc<double>((size_t)nm * neri);
    double *Eel = w2.alloc<double>(nm), *esteps = w2.alloc<double>((size_t)nm * max_scf), *C = w2.alloc<double>((size_t)nm * nn),
           *eps = w2.alloc<double>((size_t)nm * n);
    int *d_niter = w2.alloc<int>(nm), *d_conv = w2.alloc<int>(nm), *d_nterms = w2.alloc<int>(nm);
    double *hist = nullptr, *terms = nullptr, *Hbar = nullptr, *Sbar = nullptr;
    if (want_grad) {
      hist = w2.alloc<double>((size_t)nm * hl); terms = w2.alloc<double>((size_t)nm * (max_scf + 1) * 2 * nn);
      Hbar = w2.alloc<double>((size_t)nm * nn); Sbar = w2.alloc<double>((size_t)nm * nn);
    }
    Timer t_e(st);
    launch_batch_prepare(st, b, d_alpha_in, shape->log_alpha != 0);
    launch_batch_one_electron(st, b, S, H);
    for (int c = 0; c < 16; ++c) launch_batch_eri(st, c, b, d_qlist + qstart[c], qstart[c + 1] - qstart[c], eri);
    t_stats.ms_eri += t_e.stop();
    Timer t_s(st);
    if (launch_batch_scf(st, b, S, H, eri, max_scf, tol, want_grad ? 1 : 0, hist, Eel, d_niter, d_conv, esteps, C, eps, terms, d_nterms,
                         Hbar, Sbar) != 0)
      fail(-4, "basis too large for the shared-memory SCF kernel of the batched path");
    t_stats.ms_scf += t_s.stop();
    std::vector<double> hE(nm); std::vector<int> hn(nm), hc(nm);
    CU(cudaMemcpyAsync(hE.data(), Eel, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hn.data(), d_niter, nm * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hc.data(), d_conv, nm * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (want_grad) {
      Timer t_g(st);
      const int lm = det ? 2 : 1;            // fixed-point accumulators are pairs of integer limbs
      double* padj = w2.alloc<double>((size_t)nrec * kAdjFields * nm * lm, true);
      double* pasum = w2.alloc<double>((size_t)M * 2 * nm * lm, true);
      const size_t na = (size_t)nm * nprim, nx = (size_t)nm * n * 3;
      double *ga = w2.alloc<double>(na, true), *gc = w2.alloc<double>(na, true), *gx = w2.alloc<double>(nx, true),
             *graw = w2.alloc<double>(na);
      double *aa = ga, *ac = gc, *ax = gx;
      if (det) { aa = w2.alloc<double>(2 * na, true); ac = w2.alloc<double>(2 * na, true); ax = w2.alloc<double>(2 * nx, true); }
      for (int c = 0; c < 16; ++c)
        launch_batch_eri_grad(st, c, b, d_qlist + qstart[c], qstart[c + 1] - qstart[c], terms, d_nterms, max_scf, padj, pasum);
      launch_batch_grad_accumulate(st, b, padj, pasum, Sbar, Hbar, aa, ac, ax);
      if (det) {
        launch_det_finalize(st, (long long)na, aa, ga, 0.0); launch_det_finalize(st, (long long)na, ac, gc, 0.0);
        launch_det_finalize(st, (long long)nx, ax, gx, 0.0);
      }
      launch_batch_grad_finish(st, b, shape->log_alpha != 0, ga, gc, graw);
      if (g_alpha) CU(cudaMemcpyAsync(g_alpha + (size_t)m0 * nprim, ga, (size_t)nm * nprim * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (g_coef) CU(cudaMemcpyAsync(g_coef + (size_t)m0 * nprim, graw, (size_t)nm * nprim * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (g_xyz) CU(cudaMemcpyAsync(g_xyz + (size_t)m0 * n * 3, gx, (size_t)nm * n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
      t_stats.ms_eri_grad += t_g.stop();
    }
    CU(cudaStreamSynchronize(st));
    for (int m = 0; m < nm; ++m) {
      dq_system one = *shape;
      one.atoms = atoms + (size_t)(m0 + m) * natoms * 3;
      if (E) E[m0 + m] = hE[m] + nuclear_repulsion(&one);                        // Energy.py:196
      if (status) status[m0 + m] = hc[m] ? 0 : 1;                                // 1 = the reference's 99999 sentinel
      if (niter) niter[m0 + m] = hn[m];
    }
  }
  check_launches("batched kernels");
  check_boys(st);
  t_stats.n_quartets = (int64_t)qlist.size() * nmol;
  t_stats.n_prim_quartets = npq * nmol;
  t_stats.ms_total = t_total.stop();
  t_stats.launches = batch_launch_count();
}

}  // namespace dq

// =============================================================================================
// C ABI
// =============================================================================================
using namespace dq;

#define DQ_TRY try {
#define DQ_CATCH } catch (const Fail& f) { return f.code; } catch (const std::exception& e) { t_err = e.what(); return -9; } return 0;

extern "C" {

const char* dq_last_error(void) { return t_err.c_str(); }
void dq_get_stats(dq_stats* out) { if (out) *out = t_stats; }
void dq_release_cache(void) { g_pool.release(); t_plan.reset(); }
int dq_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

// host-only view of the work decomposition (no GPU needed): out = {function blocks, block pairs, tiles of this rank,
// contracted quartets of this rank, primitive quartets of this rank, tiles of all ranks}
int dq_plan(const dq_system* sys, int32_t world_size, int32_t rank, int64_t* out) {
  DQ_TRY
  if (world_size < 1 || rank < 0 || rank >= world_size) fail(-1, "bad world_size / rank");
  Prepared P;
  P.plan_only = true;
  dq_options o; memset(&o, 0, sizeof o);
  o.world_size = world_size; o.rank = rank;
  o.allreduce = (dq_allreduce_fn)(void*)1;   // never called in plan mode
  prepare(P, sys, &o, true, false);
  out[0] = P.nblocks; out[1] = P.nbp; out[2] = (int64_t)P.ntiles; out[3] = P.n_quartets; out[4] = P.n_prim_quartets;
  out[5] = (int64_t)P.nbp * (P.nbp + 1) / 2;
  DQ_CATCH
}

int dq_normalization(const dq_system* sys, const dq_options* opt, double* out) {
  DQ_TRY
  memset(&t_stats, 0, sizeof t_stats); reset_launch_count();
  Prepared P;
  prepare(P, sys, opt, false, true);
  std::vector<double> h(P.nprim);
  CU(cudaMemcpyAsync(h.data(), P.d_cn, P.nprim * sizeof(double), cudaMemcpyDeviceToHost, P.ws.st));
  CU(cudaStreamSynchronize(P.ws.st));
  for (int p = 0; p < P.nprim; ++p) out[P.p_orig[p]] = h[p];
  t_stats.launches = launch_count();
  DQ_CATCH
}

int dq_one_electron(const dq_system* sys, const dq_options* opt, double* S, double* T, double* V) {
  DQ_TRY
  memset(&t_stats, 0, sizeof t_stats); reset_launch_count();
  Prepared P;
  prepare(P, sys, opt, false, false);
  const size_t nn = (size_t)P.N * P.N;
  double *dS = P.ws.alloc<double>(nn), *dT = P.ws.alloc<double>(nn), *dV = P.ws.alloc<double>(nn);
  launch_one_electron(P.ws.st, P.sd, dS, dT, dV);
  std::vector<double> h(nn);
  double* outs[3] = {S, T, V}; double* devs[3] = {dS, dT, dV};
  for (int m = 0; m < 3; ++m) {
    CU(cudaMemcpyAsync(h.data(), devs[m], nn * sizeof(double), cudaMemcpyDeviceToHost, P.ws.st));
    CU(cudaStreamSynchronize(P.ws.st));
    if (outs[m]) unsort_matrix(P, h, outs[m], true, true);
  }
  check_boys(P.ws.st);
  t_stats.launches = launch_count();
  DQ_CATCH
}

int dq_eri_packed(const dq_system* sys, const dq_options* opt, double* eri) {
  DQ_TRY
  memset(&t_stats, 0, sizeof t_stats); reset_launch_count();
  Prepared P;
  prepare(P, sys, opt, true, false);
  Timer t(P.ws.st);
  compute_eri_store(P);
  t_stats.ms_eri = t.stop();
  const long long n = P.N, m = n * (n + 1) / 2, len = m * (m + 1) / 2;
  double* dpk = P.ws.alloc<double>(len, true);
  launch_scatter_packed(P.ws.st, P.d_tiles, (int)P.ntiles, P.sd, P.d_forig, P.d_store, dpk);
  if (P.world > 1) P.allreduce(dpk, len, (void*)P.ws.st, P.ar_user);
  CU(cudaMemcpyAsync(eri, dpk, len * sizeof(double), cudaMemcpyDeviceToHost, P.ws.st));
  CU(cudaStreamSynchronize(P.ws.st));
  check_boys(P.ws.st);
  t_stats.n_tiles = (int64_t)P.ntiles; t_stats.n_quartets = P.n_quartets; t_stats.n_prim_quartets = P.n_prim_quartets;
  t_stats.launches = launch_count();
  DQ_CATCH
}

int dq_fock(const dq_system* sys, const dq_options* opt, const double* D, double* F, double* Hcore) {
  DQ_TRY
  memset(&t_stats, 0, sizeof t_stats); reset_launch_count();
  Prepared P;
  prepare(P, sys, opt, true, false);
  cudaStream_t st = P.ws.st;
  const int n = P.N; const size_t nn = (size_t)n * n;
  setup_jk_mode(P, opt);
  compute_eri_store(P);
  double *S = P.ws.alloc<double>(nn), *T = P.ws.alloc<double>(nn), *V = P.ws.alloc<double>(nn), *H = P.ws.alloc<double>(nn);
  launch_one_electron(st, P.sd, S, T, V);
  launch_axpby(st, (int)nn, 1.0, T, 0.0, H);
  launch_axpby(st, (int)nn, 1.0, V, 1.0, H);
  std::vector<double> hd(nn);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) hd[i * n + j] = D[P.f_orig[i] * n + P.f_orig[j]];
  double* dD = P.ws.upload(hd);
  double* dF = P.ws.alloc<double>(nn);
  JK jk; jk.JKacc = P.ws.alloc<double>(2 * nn);
  if (P.sd.det) jk.JKraw = P.ws.alloc<double>(4 * nn);
  fock_build(P, jk, dD, H, dF, nullptr);
  std::vector<double> h(nn);
  CU(cudaMemcpyAsync(h.data(), dF, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  unsort_matrix(P, h, F, true, true);
  if (Hcore) {
    CU(cudaMemcpyAsync(h.data(), H, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    unsort_matrix(P, h, Hcore, true, true);
  }
  check_boys(st);
  t_stats.launches = launch_count();
  DQ_CATCH
}

int dq_rhf(const dq_system* sys, const dq_options* opt, double* E, int32_t* niter, double* esteps, double* C, double* eps) {
  try {
    ScfOut o;
    run_rhf(sys, opt, o, false, nullptr, nullptr, nullptr);
    if (E) *E = o.E;
    if (niter) *niter = o.niter;
    if (esteps) for (size_t i = 0; i < o.esteps.size(); ++i) esteps[i] = o.esteps[i];
    if (C) memcpy(C, o.C.data(), o.C.size() * sizeof(double));
    if (eps) memcpy(eps, o.eps.data(), o.eps.size() * sizeof(double));
    return o.converged ? 0 : 1;
  } catch (const Fail& f) { return f.code; } catch (const std::exception& e) { t_err = e.what(); return -9; }
}

int dq_rhf_grad(const dq_system* sys, const dq_options* opt, double* E, int32_t* niter, double* g_alpha, double* g_coef,
                double* g_xyz) {
  try {
    ScfOut o;
    run_rhf(sys, opt, o, true, g_alpha, g_coef, g_xyz);
    if (E) *E = o.E;
    if (niter) *niter = o.niter;
    return o.converged ? 0 : 1;
  } catch (const Fail& f) { return f.code; } catch (const std::exception& e) { t_err = e.what(); return -9; }
}

int dq_rhf_grad_ex(const dq_system* sys, const dq_options* opt, double* E, int32_t* niter, double* g_alpha, double* g_coef,
                   double* g_xyz, double* g_atoms) {
  try {
    ScfOut o;
    run_rhf(sys, opt, o, true, g_alpha, g_coef, g_xyz, g_atoms);
    if (E) *E = o.E;
    if (niter) *niter = o.niter;
    return o.converged ? 0 : 1;
  } catch (const Fail& f) { return f.code; } catch (const std::exception& e) { t_err = e.what(); return -9; }
}

int dq_rhf_grad_batch(const dq_system* shape, int32_t nmol, const double* atoms, const double* alpha, const double* coef,
                      const double* xyz, const dq_options* opt, int32_t want_grad, double* E, int32_t* status, int32_t* niter,
                      double* g_alpha, double* g_coef, double* g_xyz) {
  DQ_TRY
  run_batch(shape, nmol, atoms, alpha, coef, xyz, opt, want_grad != 0, E, status, niter, g_alpha, g_coef, g_xyz);
  DQ_CATCH
}

int dq_eigh(int32_t n, int32_t batch, const double* A, double* w, double* V, int32_t device) {
  DQ_TRY
  need_device(device);
  reset_launch_count();
  Workspace ws; ws.dev = device;
  const size_t nn = (size_t)n * n * batch;
  double *dA = ws.alloc<double>(nn), *dW = ws.alloc<double>(nn), *dV = ws.alloc<double>(nn), *dw = ws.alloc<double>((size_t)n * batch);
  CU(cudaMemcpy(dA, A, nn * sizeof(double), cudaMemcpyHostToDevice));
  void* ework = ws.alloc<char>(eigh_workspace_bytes(n, batch));
  launch_eigh(0, n, batch, dA, dW, dV, dw, ework);
  CU(cudaMemcpy(w, dw, (size_t)n * batch * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(V, dV, nn * sizeof(double), cudaMemcpyDeviceToHost));
  t_stats.launches = launch_count();
  DQ_CATCH
}

int dq_boys(int32_t n, const double* t, double* F, double* dF, int32_t device) {
  DQ_TRY
  need_device(device);
  if (upload_boys_constants() != 0) fail(-2, "constant upload failed");
  Workspace ws; ws.dev = device;
  double *dt = ws.alloc<double>(n), *dF_ = ws.alloc<double>((size_t)5 * n), *ddF = ws.alloc<double>((size_t)5 * n);
  CU(cudaMemcpy(dt, t, n * sizeof(double), cudaMemcpyHostToDevice));
  launch_boys(0, n, dt, dF_, ddF);
  CU(cudaMemcpy(F, dF_, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(dF, ddF, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost));
  check_boys(0);
  DQ_CATCH
}

int dq_fp64_peak(int32_t device, double* tflops) {
  DQ_TRY
  need_device(device);
  cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
  Workspace ws; ws.dev = device;
  double* out = ws.alloc<double>(1);
  const int blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  launch_fp64_peak(0, blocks, 1024, out);
  CU(cudaDeviceSynchronize());
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    Timer t(0);
    launch_fp64_peak(0, blocks, iters, out);
    const double ms = t.stop();
    const double fl = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
    best = std::max(best, fl / (ms * 1e-3) * 1e-12);
  }
  *tflops = best;
  DQ_CATCH
}

}  // extern "C"
