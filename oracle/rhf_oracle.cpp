// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of DiffiQult's RHF hot path.
//
// A line-by-line restatement (not a copy: the reference is Python) of the reference's
// algorithm for the path named by BASELINE.json: Boys function, normalisation, S/T/V,
// packed ERI vector, Fock build, Roothaan SCF, and the forward-mode parameter gradient.
// Every function cites the reference file:line it follows.  It is generic over the scalar
// type: `double` gives the float path, `Fwd` (value + P first-order tangents) gives what
// the reference obtains from algopy.UTPM (third-party, un-vendored, unpinned; semantics
// restated in oracle/fwdmode.py, whose header lists the reference call sites).
//
// Pinned (tests/test_oracle.py) against: the reference's PyQuante fixtures and its recorded
// CH4 run (tests/golden/pyquante_fixtures.npz, molden_ch4.npz) and against outputs of the
// reference's own source executed under Python 3 (tests/golden/reference_*.npz, produced by
// tests/golden/make_golden.py through oracle/ref_loader.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (diffiqult_b200/) never does.
//
// Build: g++ -O2 -fopenmp -shared -fPIC -o oracle/_build/librhf_oracle.so oracle/rhf_oracle.cpp
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

// ------------------------------------------------------------------------------------
// forward-mode scalar: value + g_ndir tangents (first order, all directions at once)
// ------------------------------------------------------------------------------------
#ifndef ORC_PMAX
#define ORC_PMAX 256
#endif
static int g_ndir = 0;  // number of live tangent directions (set per gradient call)

struct Fwd {
  double v;
  double d[ORC_PMAX];
  Fwd() : v(0.0) { for (int p = 0; p < g_ndir; ++p) d[p] = 0.0; }
  Fwd(double x) : v(x) { for (int p = 0; p < g_ndir; ++p) d[p] = 0.0; }
  Fwd(const Fwd& o) : v(o.v) { for (int p = 0; p < g_ndir; ++p) d[p] = o.d[p]; }
  Fwd& operator=(const Fwd& o) { v = o.v; for (int p = 0; p < g_ndir; ++p) d[p] = o.d[p]; return *this; }
};
static inline Fwd operator+(const Fwd& a, const Fwd& b) { Fwd r; r.v = a.v + b.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] + b.d[p]; return r; }
static inline Fwd operator-(const Fwd& a, const Fwd& b) { Fwd r; r.v = a.v - b.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] - b.d[p]; return r; }
static inline Fwd operator*(const Fwd& a, const Fwd& b) { Fwd r; r.v = a.v * b.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] * b.v + b.d[p] * a.v; return r; }
static inline Fwd operator/(const Fwd& a, const Fwd& b) { Fwd r; r.v = a.v / b.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = (a.d[p] - b.d[p] * r.v) / b.v; return r; }
static inline Fwd operator+(const Fwd& a, double b) { Fwd r(a); r.v += b; return r; }
static inline Fwd operator+(double b, const Fwd& a) { return a + b; }
static inline Fwd operator-(const Fwd& a, double b) { Fwd r(a); r.v -= b; return r; }
static inline Fwd operator-(double b, const Fwd& a) { Fwd r; r.v = b - a.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = -a.d[p]; return r; }
static inline Fwd operator-(const Fwd& a) { return 0.0 - a; }
static inline Fwd operator*(const Fwd& a, double b) { Fwd r; r.v = a.v * b; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] * b; return r; }
static inline Fwd operator*(double b, const Fwd& a) { return a * b; }
static inline Fwd operator/(const Fwd& a, double b) { Fwd r; r.v = a.v / b; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] / b; return r; }
static inline Fwd operator/(double b, const Fwd& a) { Fwd r; r.v = b / a.v; double f = -r.v / a.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] * f; return r; }
// comparisons look at the value coefficient only (oracle/fwdmode.py header)
static inline bool operator<(const Fwd& a, const Fwd& b) { return a.v < b.v; }
static inline bool operator<(const Fwd& a, double b) { return a.v < b; }
static inline bool operator>(const Fwd& a, double b) { return a.v > b; }
static inline double val(double x) { return x; }
static inline double val(const Fwd& x) { return x.v; }
static inline Fwd exp(const Fwd& a) { Fwd r; r.v = std::exp(a.v); for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] * r.v; return r; }
static inline Fwd log(const Fwd& a) { Fwd r; r.v = std::log(a.v); for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] / a.v; return r; }
static inline Fwd sqrt(const Fwd& a) { Fwd r; r.v = std::sqrt(a.v); double f = 0.5 / r.v; for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] * f; return r; }
static inline Fwd pow(const Fwd& a, double e) { Fwd r; r.v = std::pow(a.v, e); double f = e * std::pow(a.v, e - 1.0); for (int p = 0; p < g_ndir; ++p) r.d[p] = a.d[p] * f; return r; }
static inline Fwd fabs(const Fwd& a) { return a.v >= 0 ? a : -a; }
using std::exp; using std::log; using std::sqrt; using std::pow; using std::fabs;

// ------------------------------------------------------------------------------------
// Tools.py
// ------------------------------------------------------------------------------------
// Tools.py:17-22  matrix_tovector(i,j,n) = i*n - i(i+1)/2 + j   (i <= j)
static inline long tri_index(long i, long j, long n) { return i * n - i * (i + 1) / 2 + j; }

// Tools.py:7-14  vec_tomatrix(x,n): inverse of tri_index
static inline void tri_unindex(long x, long n, int* i, int* j) {
  long r = x; long ii = 0;
  for (ii = 0; ii < n; ++ii) { r -= n - ii; if (r < 0) break; }
  *i = (int)ii; *j = (int)(n + r);
}

// Tools.py:25-51  eri_index
static inline long eri_index(int i, int j, int k, int l, long n) {
  int kk = k, ll = l, ii = i, jj = j;
  if (k >= l) { kk = l; ll = k; }
  if (i >= j) { ii = j; jj = i; }
  long x = tri_index(ii, jj, n), y = tri_index(kk, ll, n), m = n * (n + 1) / 2;
  return x >= y ? tri_index(y, x, m) : tri_index(x, y, m);
}

// Tools.py:99-109  gammaln (6-term Lanczos, Numerical Recipes)
static inline double nr_gammaln(double a) {
  static const double c[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                              -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953E-5};
  double y = a, x = a;
  double tmp = x + 5.5;
  tmp = tmp - (x + 0.5) * std::log(tmp);
  double ser = 1.000000000190015;
  for (int i = 0; i < 6; ++i) { y = y + 1.0; ser = ser + c[i] / y; }
  return -tmp + std::log(2.5066282746310005 * ser / x);
}

static int g_boys_fail = 0;

// Tools.py:111-125  gser: series, eps 3e-12; returns gamma(a,x) (not regularised)
template <class T> static T nr_gser(double a, const T& x) {
  const double eps = 3.0E-12;
  double ap = a;
  T de = T(1.0 / a), summ = T(1.0 / a);
  for (int i = 0; i < 10000; ++i) {
    ap = ap + 1.0;
    de = de * x / ap;
    summ = summ + de;
    if (de < summ * eps) {
      double gln = nr_gammaln(a);
      return summ * (exp(((-x) + a * log(x)) + (-gln)) * std::exp(gln));
    }
  }
  g_boys_fail = 1;
  return T(0.0);
}

// Tools.py:127-151  gcf: modified Lentz, eps 3e-7, fpmin 1e-30
template <class T> static T nr_gcf(double a, const T& x) {
  const double eps = 3.0E-7, fpmin = 1.0E-30;
  T b = x + 1.0 - a;
  T c = T(1.0 / fpmin);
  T d = 1.0 / b;
  T h = d;
  for (int i = 1; i < 10000; ++i) {
    double an = -i * (i - a);
    b = 2.0 + b;
    d = an * d + b;
    if (fabs(d) < fpmin) d = T(fpmin);
    c = b + an / c;
    if (fabs(c) < fpmin) c = T(fpmin);
    d = 1.0 / d;
    T de = d * c;
    h = h * de;
    if (fabs(de - 1.0) < eps) {
      double gln = nr_gammaln(a);
      return (1.0 - exp((-x) + a * log(x) - gln) * h) * std::exp(gln);
    }
  }
  g_boys_fail = 1;
  return T(0.0);
}

// Tools.py:87-97  getgammp
template <class T> static T getgammp(double a, const T& x) {
  if (x < a + 1.0) return nr_gser(a, x);
  return nr_gcf(a, x);
}

// Tools.py:78-83  incompletegammaf(t,n) = F_n(t); builtin max(t,1e-12): a clamped t loses its tangent
template <class T> static T boys(const T& t_in, int n) {
  T t = (1e-12 > val(t_in)) ? T(1e-12) : t_in;
  return getgammp(n + 0.5, t) * pow(t, -n - 0.5) * 0.5;
}

// ------------------------------------------------------------------------------------
// system description shared by all entry points
// ------------------------------------------------------------------------------------
struct Sys {
  int natoms; const int* charges; const double* atoms;   // [natoms][3]
  int nbasis; const int* contr; const int* ltype;        // ltype: 0 s, 1 px, 2 py, 3 pz
  int nprim;
  int ne;                                                 // doubly-occupied orbitals (electrons // 2)
  std::vector<int> off;                                   // first primitive of each function
  void finish() { off.resize(nbasis + 1); off[0] = 0; for (int i = 0; i < nbasis; ++i) off[i + 1] = off[i] + contr[i]; }
};

template <class T> struct V3 { T x[3]; T& operator[](int i) { return x[i]; } const T& operator[](int i) const { return x[i]; } };

template <class T> static T norm2(const V3<T>& a, const V3<T>& b) {  // Tools.py:53-54 on np.subtract(a,b)
  T s = (a[0] - b[0]) * (a[0] - b[0]);
  s = s + (a[1] - b[1]) * (a[1] - b[1]);
  s = s + (a[2] - b[2]) * (a[2] - b[2]);
  return s;
}

// ------------------------------------------------------------------------------------
// Integrals.py:463-501  normalization
// ------------------------------------------------------------------------------------
template <class T> static T norm_primitive(const T& alpha, int lt) {  // Integrals.py:463-474
  // l is a unit vector for p (one entry 1): loop k in range(1,m+1,2) runs once -> div = sqrt(2); s: div = 1
  int l_large = lt > 0 ? 1 : 0;
  double power = l_large / 2.0;
  double div = 1.0;
  if (lt > 0) { div = div * 2.0 * 1; div = std::sqrt(div); }
  double factor = std::pow(2.0 / M_PI, 0.75) * div * std::pow(2.0, power);
  power = power + 0.75;
  return factor * pow(alpha, power);
}

template <class T> static void normalization(const Sys& s, const T* alpha, const T* c, T* coef) {  // Integrals.py:477-501
  int contr = 0;
  for (int ib = 0; ib < s.nbasis; ++ib) {
    int ci = s.contr[ib];
    double div = 1.0;
    int l_large = 0;
    if (s.ltype[ib] > 0) { l_large += 1; div = div * 1; }  // range(1,2m,2) with m=1 -> k=1
    div = div / std::pow(2.0, l_large);
    div = div * std::pow(M_PI, 1.5);
    for (int i = 0; i < ci; ++i) coef[contr + i] = norm_primitive(alpha[contr + i], s.ltype[ib]) * c[contr + i];
    T tmp = T(0.0);
    for (int i = 0; i < ci; ++i)
      for (int j = 0; j < ci; ++j)
        tmp = tmp + coef[contr + j] * coef[contr + i] / pow(alpha[contr + i] + alpha[contr + j], l_large + 1.5);
    tmp = sqrt(tmp * div);
    for (int i = contr; i < contr + ci; ++i) coef[i] = coef[i] / tmp;
    contr += ci;
  }
}

// ------------------------------------------------------------------------------------
// one-electron primitives
// ------------------------------------------------------------------------------------
// Integrals.py:30-48
template <class T> static T overlap_primitive(const T& alpha, const T& ca, const V3<T>& A, int la,
                                              const T& beta, const T& cb, const V3<T>& B, int lb) {
  T gamma = 1.0 / (alpha + beta);
  T inti = T(1.0);
  for (int x = 0; x < 3; ++x) {
    int l1 = (la == x + 1), l2 = (lb == x + 1);
    T tmp = T(1.0);
    if (l1 > 0) tmp = tmp * (B[x] - A[x]) * beta * gamma;
    if (l2 > 0) tmp = tmp * (A[x] - B[x]) * alpha * gamma;
    if (l1 + l2 == 2) tmp = tmp + 0.5 * gamma;
    inti = tmp * inti;
  }
  T ab = -1.0 * norm2(A, B);
  T ex = exp((alpha * beta * ab) * gamma);
  inti = inti * pow(M_PI * gamma, 1.5);
  return ex * inti * ca * cb;
}

// Integrals.py:155-188
template <class T> static T kinetic_primitive(const T& alpha, const T& ca, const V3<T>& A, int la,
                                              const T& beta, const T& cb, const V3<T>& B, int lb) {
  T gamma = 1.0 / (alpha + beta);
  T eta = alpha * beta * gamma;
  T ab = -1.0 * norm2(A, B);
  T S00 = ca * cb * exp(ab * eta) * pow(M_PI * gamma, 1.5);
  T K00 = 3.0 * eta + 2.0 * eta * eta * ab;
  int n = la > 0, m = lb > 0;
  if (n + m == 0) return S00 * K00;
  if (n + m == 2) {
    int i = la - 1, j = lb - 1;
    T Ki0 = 2.0 * eta * gamma * beta * (B[i] - A[i]);
    T K0j = 2.0 * eta * gamma * alpha * (A[j] - B[j]);
    T Si0 = (B[i] - A[i]) * beta * gamma;
    T S0j = (A[j] - B[j]) * alpha * gamma;
    T Sij = S0j * Si0;
    T tot = S0j * Ki0 + Si0 * K0j;
    if (i == j) {
      T Kij = eta * gamma;
      Sij = Sij + 0.5 * gamma;
      return S00 * (tot + Sij * K00 + Kij);
    }
    return S00 * (tot + Sij * K00);
  }
  if (n == 1) {
    int i = la - 1;
    T Ki0 = 2.0 * eta * gamma * beta * (B[i] - A[i]);
    T Si0 = (B[i] - A[i]) * beta * gamma;
    return (Si0 * K00 + Ki0) * S00;
  }
  int j = lb - 1;
  T K0j = 2.0 * eta * gamma * alpha * (A[j] - B[j]);
  T S0j = (A[j] - B[j]) * alpha * gamma;
  return (S0j * K00 + K0j) * S00;
}

// Integrals.py:75-133  (C = nuclear positions, plain doubles: the Tasks path never differentiates them)
template <class T> static T nuclear_primitive(const T& alpha, const T& ca, const V3<T>& A, int la,
                                              const T& beta, const T& cb, const V3<T>& B, int lb,
                                              const double* C, const int* charge, int numatoms) {
  T gamma = alpha + beta;
  T gammainv = 1.0 / gamma;
  T ab = -1.0 * norm2(A, B);
  T S00 = ca * cb * exp(ab * alpha * beta * gammainv) * pow(M_PI * gammainv, 1.5);
  V3<T> P;
  for (int x = 0; x < 3; ++x) P[x] = (alpha * A[x] + beta * B[x]) * gammainv;
  int n = la > 0, m = lb > 0;
  auto pc_of = [&](int k) {
    T s = (P[0] - C[3 * k + 0]) * (P[0] - C[3 * k + 0]);
    s = s + (P[1] - C[3 * k + 1]) * (P[1] - C[3 * k + 1]);
    s = s + (P[2] - C[3 * k + 2]) * (P[2] - C[3 * k + 2]);
    return gamma * s;
  };
  if (n + m == 0) {
    T nuclear = T(0.0);
    for (int i = 0; i < numatoms; ++i) {
      T pc = pc_of(i);
      nuclear = nuclear - (double)charge[i] * boys(pc, 0);
    }
    return nuclear * S00 * 2.0 * pow(gamma / M_PI, 0.5);
  }
  if (n + m == 2) {
    int i = la - 1, j = lb - 1;
    T Si0 = (B[i] - A[i]) * beta * gammainv;
    T S0j = (A[j] - B[j]) * alpha * gammainv;
    T Sij = S0j * Si0;
    if (i == j) Sij = Sij + 0.5 * gammainv;
    T nuclear = T(0.0);
    for (int k = 0; k < numatoms; ++k) {
      T pc = pc_of(k);
      T L00 = boys(pc, 0);
      T L0j = C[3 * k + j] - P[j];
      T Li0 = C[3 * k + i] - P[i];
      T Lij = L0j * Li0 * boys(pc, 2);
      L0j = L0j * boys(pc, 1);
      Li0 = Li0 * boys(pc, 1);
      if (i == j) Lij = Lij - 0.5 * gammainv * boys(pc, 1);
      nuclear = nuclear - (double)charge[k] * (Sij * L00 + Si0 * L0j + S0j * Li0 + Lij);
    }
    return nuclear * S00 * 2.0 * pow(gamma / M_PI, 0.5);
  }
  if (n == 1) {
    int i = la - 1;
    T Si0 = (B[i] - A[i]) * beta * gammainv;
    T nuclear = T(0.0);
    for (int k = 0; k < numatoms; ++k) {
      T pc = pc_of(k);
      T L00 = boys(pc, 0);
      T Li0 = (C[3 * k + i] - P[i]) * boys(pc, 1);
      nuclear = nuclear - (double)charge[k] * (Li0 + L00 * Si0);
    }
    return nuclear * 2.0 * sqrt(gamma / M_PI) * S00;
  }
  int j = lb - 1;
  T S0j = (A[j] - B[j]) * alpha * gammainv;
  T nuclear = T(0.0);
  for (int k = 0; k < numatoms; ++k) {
    T pc = pc_of(k);
    T L00 = boys(pc, 0);
    T L0j = (C[3 * k + j] - P[j]) * boys(pc, 1);
    nuclear = nuclear - (double)charge[k] * (S0j * L00 + L0j);
  }
  return nuclear * S00 * 2.0 * pow(gamma / M_PI, 0.5);
}

// Integrals.py:10-20, 51-62, 135-145 + the *_contracted double loops (:22-28, :64-71, :147-153)
template <class T> static void one_electron(const Sys& s, const T* alpha, const T* coef, const V3<T>* xyz,
                                            T* S, T* Tm, T* Vm) {
  int n = s.nbasis;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      T sv = T(0.0), tv = T(0.0), vv = T(0.0);
      for (int p = s.off[i]; p < s.off[i + 1]; ++p)
        for (int q = s.off[j]; q < s.off[j + 1]; ++q) {
          sv = sv + overlap_primitive(alpha[p], coef[p], xyz[i], s.ltype[i], alpha[q], coef[q], xyz[j], s.ltype[j]);
          tv = tv + kinetic_primitive(alpha[p], coef[p], xyz[i], s.ltype[i], alpha[q], coef[q], xyz[j], s.ltype[j]);
          vv = vv + nuclear_primitive(alpha[p], coef[p], xyz[i], s.ltype[i], alpha[q], coef[q], xyz[j], s.ltype[j],
                                      s.atoms, s.charges, s.natoms);
        }
      S[i * n + j] = S[j * n + i] = sv;
      Tm[i * n + j] = Tm[j * n + i] = tv;
      Vm[i * n + j] = Vm[j * n + i] = vv;
    }
}

// ------------------------------------------------------------------------------------
// Integrals.py:234-401  eris_primitive
// ------------------------------------------------------------------------------------
template <class T> struct EriCtx {
  T Fs[5]; T rho; T nugammainv; V3<T> wpq;
  // Integrals.py:276-284
  void psss(int i, const V3<T>& P, const V3<T>& Ax, int m, T* res) const {
    T PA = P[i] - Ax[i];
    T WP = wpq[i] - P[i];
    for (int n = 0; n <= m; ++n) res[n] = PA * Fs[0 + n] + WP * Fs[1 + n];
  }
  // Integrals.py:286-298
  void ppss(int i, int j, const V3<T>& P, const V3<T>& Ax, const V3<T>& Bx, const T& ginv, int m, T* res) const {
    T aux[5];
    psss(i, P, Ax, m + 1, aux);
    T PB = P[j] - Bx[j];
    T WP = wpq[j] - P[j];
    for (int n = 0; n <= m; ++n) {
      T tmp = PB * aux[n] + WP * aux[n + 1];
      if (i == j) tmp = tmp + 0.5 * ginv * (Fs[0 + n] - rho * ginv * Fs[n + 1]);
      res[n] = tmp;
    }
  }
  // Integrals.py:300-312
  void psps(int i, int k, const V3<T>& P, const V3<T>& Q, const V3<T>& Ax, const V3<T>& Cx, const T& nginv, int m,
            T* res) const {
    T aux[5];
    psss(i, P, Ax, m + 1, aux);
    T QC = Q[k] - Cx[k];
    T WQ = wpq[k] - Q[k];
    for (int n = 0; n <= m; ++n) {
      T tmp = QC * aux[n] + WQ * aux[n + 1];
      if (i == k) tmp = tmp + 0.5 * nginv * Fs[1 + n];
      res[n] = tmp;
    }
  }
  // Integrals.py:314-330
  void ppps(int i, int j, int k, const V3<T>& P, const V3<T>& Q, const V3<T>& Ax, const V3<T>& Bx, const V3<T>& Cx,
            const T& ginv, int m, T* res) const {
    T aux[5], sp[5], ps[5];
    ppss(i, j, P, Ax, Bx, ginv, m + 1, aux);
    T QC = Q[k] - Cx[k];
    T WQ = wpq[k] - Q[k];
    if (i == k) psss(j, P, Bx, m + 1, sp);
    if (j == k) psss(i, P, Ax, m + 1, ps);
    for (int n = 0; n <= m; ++n) {
      T tmp = QC * aux[n] + WQ * aux[n + 1];
      if (i == k) tmp = tmp + 0.5 * nugammainv * sp[n + 1];
      if (j == k) tmp = tmp + 0.5 * nugammainv * ps[n + 1];
      res[n] = tmp;
    }
  }
};

template <class T> static T eris_primitive(const T& a, const T& coefa, const V3<T>& A, int la,
                                           const T& b, const T& coefb, const V3<T>& B, int lb,
                                           const T& c, const T& coefc, const V3<T>& C, int lc,
                                           const T& d, const T& coefd, const V3<T>& D, int ld) {
  T gamma = a + b;
  T gammainv = 1.0 / gamma;
  V3<T> pab, qcd;
  for (int x = 0; x < 3; ++x) pab[x] = (a * A[x] + b * B[x]) * gammainv;
  T fab = coefa * coefb;
  T ab = norm2(A, B) * gammainv;
  T kab = fab * exp(-1.0 * a * b * ab) * gammainv;

  T nu = c + d;
  T nuinv = 1.0 / nu;
  for (int x = 0; x < 3; ++x) qcd[x] = (c * C[x] + d * D[x]) * nuinv;
  T fcd = coefc * coefd;
  T cd = norm2(C, D) * nuinv;
  T kcd = fcd * exp(-1.0 * c * d * cd) * nuinv;

  EriCtx<T> w;
  w.nugammainv = 1.0 / (nu + gamma);
  w.rho = nu * gamma * w.nugammainv;
  T t = w.rho * norm2(pab, qcd);

  int na = la > 0, nb = lb > 0, nc = lc > 0, nd = ld > 0;
  int nF = na + nb + nc + nd + 1;
  for (int i = 0; i < nF; ++i) w.Fs[i] = boys(t, i);  // Integrals.py:261-267, each order independently

  if (nF == 1) return kcd * kab * 2.0 * std::pow(M_PI, 2.5) * w.Fs[0] * sqrt(w.nugammainv);

  T prefactor = kcd * kab * 2.0 * std::pow(M_PI, 2.5) * sqrt(w.nugammainv);
  for (int x = 0; x < 3; ++x) w.wpq[x] = (gamma * pab[x] + nu * qcd[x]) * w.nugammainv;
  int ia = la - 1, ib = lb - 1, ic = lc - 1, id = ld - 1;
  T res[5];

  if (nF == 2) {  // Integrals.py:365-373
    if (nd == 1) { w.psss(id, qcd, D, 0, res); return prefactor * res[0]; }
    if (nc == 1) { w.psss(ic, qcd, C, 0, res); return prefactor * res[0]; }
    if (nb == 1) { w.psss(ib, pab, B, 0, res); return prefactor * res[0]; }
    w.psss(ia, pab, A, 0, res); return prefactor * res[0];
  }
  if (nF == 3) {  // Integrals.py:375-386
    if (nd + nc == 2) { w.ppss(ic, id, qcd, C, D, nuinv, 0, res); return prefactor * res[0]; }
    if (na + nb == 2) { w.ppss(ia, ib, pab, A, B, gammainv, 0, res); return prefactor * res[0]; }
    if (nb + nd == 2) { w.psps(ib, id, pab, qcd, B, D, w.nugammainv, 0, res); return prefactor * res[0]; }
    if (nb + nc == 2) { w.psps(ib, ic, pab, qcd, B, C, w.nugammainv, 0, res); return prefactor * res[0]; }
    if (na + nc == 2) { w.psps(ia, ic, pab, qcd, A, C, w.nugammainv, 0, res); return prefactor * res[0]; }
    w.psps(ia, id, pab, qcd, A, D, w.nugammainv, 0, res); return prefactor * res[0];
  }
  if (nF == 5) {  // Integrals.py:334-363, 388-389  pppp(0)
    int i = ia, j = ib, k = ic, l = id;
    T aux[5];
    w.ppps(i, j, k, pab, qcd, A, B, C, gammainv, 1, aux);
    T QD = qcd[l] - D[l];
    T WQ = w.wpq[l] - qcd[l];
    T tmp = QD * aux[0] + WQ * aux[1];
    if (i == l) { T sp[5]; w.psps(j, k, pab, qcd, B, C, w.nugammainv, 1, sp); tmp = tmp + 0.5 * w.nugammainv * sp[1]; }
    if (j == l) { T ps[5]; w.psps(i, k, pab, qcd, A, C, w.nugammainv, 1, ps); tmp = tmp + 0.5 * w.nugammainv * ps[1]; }
    if (k == l) { T pp[5]; w.ppss(i, j, pab, A, B, gammainv, 1, pp); tmp = tmp + 0.5 * nuinv * (pp[0] - nuinv * w.rho * pp[1]); }
    return tmp * prefactor;
  }
  // nF == 4, Integrals.py:391-399
  if (nd + nc == 2) {
    if (nb == 1) { w.ppps(ic, id, ib, qcd, pab, C, D, B, nuinv, 0, res); return prefactor * res[0]; }
    w.ppps(ic, id, ia, qcd, pab, C, D, A, nuinv, 0, res); return prefactor * res[0];
  }
  if (nc == 1) { w.ppps(ia, ib, ic, pab, qcd, A, B, C, gammainv, 0, res); return prefactor * res[0]; }
  w.ppps(ia, ib, id, pab, qcd, A, B, D, gammainv, 0, res); return prefactor * res[0];
}

// Integrals.py:218-230
template <class T> static T eri_contracted(const Sys& s, const T* alpha, const T* coef, const V3<T>* xyz,
                                           int i, int j, int k, int m) {
  T eris = T(0.0);
  for (int p = s.off[i]; p < s.off[i + 1]; ++p)
    for (int q = s.off[j]; q < s.off[j + 1]; ++q)
      for (int r = s.off[k]; r < s.off[k + 1]; ++r)
        for (int u = s.off[m]; u < s.off[m + 1]; ++u)
          eris = eris + eris_primitive(alpha[p], coef[p], xyz[i], s.ltype[i], alpha[q], coef[q], xyz[j], s.ltype[j],
                                       alpha[r], coef[r], xyz[k], s.ltype[k], alpha[u], coef[u], xyz[m], s.ltype[m]);
  return eris;
}

// Integrals.py:190-216
template <class T> static void erivector(const Sys& s, const T* alpha, const T* coef, const V3<T>* xyz, T* eri) {
  long n = s.nbasis, len_vec = n * (n + 1) / 2;
#pragma omp parallel for schedule(dynamic, 1)
  for (long x = 0; x < len_vec; ++x) {
    int i, j; tri_unindex(x, n, &i, &j);
    for (long y = x; y < len_vec; ++y) {
      int k, m; tri_unindex(y, n, &k, &m);
      eri[tri_index(x, y, len_vec)] = eri_contracted(s, alpha, coef, xyz, i, j, k, m);
    }
  }
}

// ------------------------------------------------------------------------------------
// Energy.py
// ------------------------------------------------------------------------------------
// Energy.py:21-27 (+ Tools.py:56-57)
static double nuclearrepulsion(const Sys& s) {
  double tmp = 0.0;
  for (int i = 0; i < s.natoms; ++i)
    for (int j = i + 1; j < s.natoms; ++j) {
      double r2 = 0; for (int x = 0; x < 3; ++x) { double d = s.atoms[3 * j + x] - s.atoms[3 * i + x]; r2 += d * d; }
      tmp = tmp + s.charges[i] * s.charges[j] / std::sqrt(r2);
    }
  return tmp;
}

// Energy.py:29-38
template <class T> static void fockmatrix(int n, const T* H, const T* eri, const T* D, T* F) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      T tmp = H[i * n + j];
      for (int k = 0; k < n; ++k)
        for (int l = 0; l < n; ++l)
          tmp = D[k * n + l] * (2.0 * eri[eri_index(i, j, k, l, n)] - eri[eri_index(i, k, j, l, n)]) + tmp;
      F[i * n + j] = F[j * n + i] = tmp;
    }
}

// symmetric eigensolver for plain doubles: cyclic Jacobi, ascending eigenvalues (what LAPACK dsyevd returns
// inside algopy.eigh, Eigen.py:69-74, up to eigenvector sign / rotation inside degenerate blocks)
static void jacobi_eigh(int n, const double* Ain, double* w, double* Q) {
  std::vector<double> A(Ain, Ain + (size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Q[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0; for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = A[p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < n; ++k) {
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - sn * akq; A[k * n + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - sn * aqk; A[q * n + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          double qkp = Q[k * n + p], qkq = Q[k * n + q];
          Q[k * n + p] = c * qkp - sn * qkq; Q[k * n + q] = sn * qkp + c * qkq;
        }
      }
  }
  std::vector<int> idx(n); for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
  std::vector<double> Qs((size_t)n * n);
  for (int c = 0; c < n; ++c) { w[c] = A[idx[c] * n + idx[c]]; for (int r = 0; r < n; ++r) Qs[r * n + c] = Q[r * n + idx[c]]; }
  std::memcpy(Q, Qs.data(), sizeof(double) * n * n);
}

static void eigensolver(int n, const double* A, double* w, double* Q) { jacobi_eigh(n, A, w, Q); }

static int g_degenerate_coupling = 0;
// Eigen.py:69-74 on UTPM: value by eigh, tangents by the first-order push-forward
//   dw = diag(Q^T dA Q), dQ = Q (H o Q^T dA Q), H_ij = 1/(w_j - w_i)   (oracle/fwdmode.py:eigh)
static void eigensolver(int n, const Fwd* A, Fwd* w, Fwd* Q) {
  std::vector<double> a((size_t)n * n), wv(n), q((size_t)n * n), m((size_t)n * n), t((size_t)n * n);
  for (int i = 0; i < n * n; ++i) a[i] = A[i].v;
  jacobi_eigh(n, a.data(), wv.data(), q.data());
  for (int i = 0; i < n; ++i) w[i] = Fwd(wv[i]);
  for (int i = 0; i < n * n; ++i) Q[i] = Fwd(q[i]);
  for (int p = 0; p < g_ndir; ++p) {
    // t = dA q ; m = q^T t
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += A[i * n + k].d[p] * q[k * n + j]; t[i * n + j] = s; }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += q[k * n + i] * t[k * n + j]; m[i * n + j] = s; }
    for (int i = 0; i < n; ++i) w[i].d[p] = m[i * n + i];
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
      double gap = wv[j] - wv[i];
      if (i == j || std::fabs(gap) < 1e-9) { if (i != j && std::fabs(m[i * n + j]) > 1e-7) g_degenerate_coupling = 1; m[i * n + j] = 0.0; }
      else m[i * n + j] /= gap;
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += q[i * n + k] * m[k * n + j]; Q[i * n + j].d[p] = s; }
  }
}

template <class T> static void matmul(int n, const T* A, const T* B, T* C) {  // algopy.dot
  std::vector<T> out((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { T s = T(0.0); for (int k = 0; k < n; ++k) s = s + A[i * n + k] * B[k * n + j]; out[i * n + j] = s; }
  for (int i = 0; i < n * n; ++i) C[i] = out[i];
}
template <class T> static void transpose(int n, const T* A, T* At) {
  std::vector<T> out((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out[j * n + i] = A[i * n + j];
  for (int i = 0; i < n * n; ++i) At[i] = out[i];
}

struct ScfOut { int status; int niter; double esteps[1024]; double e_nuc; };
static const double* g_guess_C = nullptr;   // readguess: AO x MO coefficients of a previous run (Energy.py:148-157), or null

// Energy.py:75-199, 311-324 with eigen=True, readguess=None (the only branch Tasks reaches)
template <class T> static T rhfenergy(const Sys& s, const T* alpha_in, const T* coef2, const V3<T>* xyz, int max_scf,
                                      bool logalpha, ScfOut* so, T* Cout, T* eps_out, std::vector<T>* eri_keep = nullptr) {
  const double tool = 1e-8;
  int n = s.nbasis, np = s.nprim;
  std::vector<T> alpha(np), coef(np);
  for (int p = 0; p < np; ++p) alpha[p] = logalpha ? exp(alpha_in[p]) : alpha_in[p];  // Energy.py:120-123
  normalization(s, alpha.data(), coef2, coef.data());                                   // Energy.py:132
  std::vector<T> S((size_t)n * n), Tm((size_t)n * n), Vm((size_t)n * n), H((size_t)n * n);
  one_electron(s, alpha.data(), coef.data(), xyz, S.data(), Tm.data(), Vm.data());        // Energy.py:133-135
  long mm = (long)n * (n + 1) / 2;
  std::vector<T> eri((size_t)(mm * (mm + 1) / 2));
  erivector(s, alpha.data(), coef.data(), xyz, eri.data());                               // Energy.py:136
  for (int i = 0; i < n * n; ++i) H[i] = Tm[i] + Vm[i];                                   // Energy.py:137
  // Energy.py:138-144  S^-1/2 = L diag(1/sqrt(w)) L^T
  std::vector<T> w(n), L((size_t)n * n), LT((size_t)n * n), SqL((size_t)n * n), X((size_t)n * n), XT((size_t)n * n);
  eigensolver(n, S.data(), w.data(), L.data());
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) SqL[i * n + j] = (i == j) ? 1.0 / sqrt(w[i]) : T(0.0);
  transpose(n, L.data(), LT.data());
  matmul(n, L.data(), SqL.data(), X.data());
  matmul(n, X.data(), LT.data(), X.data());
  transpose(n, X.data(), XT.data());
  std::vector<T> F(H), Fp((size_t)n * n), Cp((size_t)n * n), C((size_t)n * n), D((size_t)n * n), ev(n);
  if (g_guess_C) {                                                                      // Energy.py:148-157
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) { double tmp = 0.0; for (int k = 0; k < s.ne; ++k) tmp = tmp + g_guess_C[i * n + k] * g_guess_C[j * n + k]; D[i * n + j] = T(tmp); }
    fockmatrix(n, H.data(), eri.data(), D.data(), F.data());
  }
  T OldE = T(1e8), E_elec = T(0.0);
  so->status = 0; so->niter = 0;
  for (int it = 0; it < max_scf; ++it) {
    // Energy.py:166-173 (the block is duplicated verbatim in the reference; the second pass recomputes the same C)
    matmul(n, XT.data(), F.data(), Fp.data());
    matmul(n, Fp.data(), X.data(), Fp.data());
    eigensolver(n, Fp.data(), ev.data(), Cp.data());
    matmul(n, X.data(), Cp.data(), C.data());
    for (int i = 0; i < n; ++i)                                                          // Energy.py:174-180
      for (int j = 0; j < n; ++j) { T tmp = T(0.0); for (int k = 0; k < s.ne; ++k) tmp = tmp + C[i * n + k] * C[j * n + k]; D[i * n + j] = tmp; }
    fockmatrix(n, H.data(), eri.data(), D.data(), F.data());                             // Energy.py:188
    E_elec = T(0.0);
    for (int i = 0; i < n * n; ++i) E_elec = E_elec + D[i] * (H[i] + F[i]);               // Energy.py:189
    if (it < 1024) so->esteps[it] = val(E_elec);
    so->niter = it + 1;
    if (std::fabs(val(E_elec) - val(OldE)) < tool) { so->status = 1; break; }             // Energy.py:192-194
    OldE = E_elec;
  }
  so->e_nuc = nuclearrepulsion(s);
  if (Cout) for (int i = 0; i < n * n; ++i) Cout[i] = C[i];
  if (eps_out) for (int i = 0; i < n; ++i) eps_out[i] = ev[i];
  if (eri_keep) eri_keep->swap(eri);
  return E_elec + so->e_nuc;  // the caller maps status==0 to the reference's 99999 sentinel (Energy.py:317-322)
}

// ------------------------------------------------------------------------------------
// C entry points (ctypes)
// ------------------------------------------------------------------------------------
static Sys make_sys(int natoms, const int* charges, const double* atoms, int nbasis, const int* contr, const int* ltype,
                    int ne) {
  Sys s; s.natoms = natoms; s.charges = charges; s.atoms = atoms; s.nbasis = nbasis; s.contr = contr; s.ltype = ltype;
  s.ne = ne; s.finish(); s.nprim = s.off[nbasis]; return s;
}

extern "C" {

int orc_pmax(void) { return ORC_PMAX; }
void orc_set_guess(const double* C) { g_guess_C = C; }
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

long orc_eri_index(int i, int j, int k, int l, int n) { return eri_index(i, j, k, l, n); }

// F_m(t) and d/dt through the same loops
int orc_boys(int n, const double* t, int m, double* F, double* dF) {
  g_boys_fail = 0;
  int saved = g_ndir; g_ndir = 1;
  for (int i = 0; i < n; ++i) {
    F[i] = boys(t[i], m);
    Fwd x(t[i]); x.d[0] = 1.0;
    Fwd r = boys(x, m);
    dF[i] = r.d[0];
  }
  g_ndir = saved;
  return g_boys_fail;
}

int orc_normalization(int nbasis, const int* contr, const int* ltype, const double* alpha, const double* c, double* out) {
  Sys s = make_sys(0, nullptr, nullptr, nbasis, contr, ltype, 0);
  normalization(s, alpha, c, out);
  return 0;
}

// S, T, V with ALREADY-NORMALISED coefficients (as the reference's unit tests call them, tests/Test.py:176)
int orc_one_electron(int natoms, const int* charges, const double* atoms, int nbasis, const int* contr, const int* ltype,
                     const double* alpha, const double* coefn, const double* xyz, double* S, double* T, double* V) {
  Sys s = make_sys(natoms, charges, atoms, nbasis, contr, ltype, 0);
  g_boys_fail = 0;
  one_electron(s, alpha, coefn, reinterpret_cast<const V3<double>*>(xyz), S, T, V);
  return g_boys_fail;
}

int orc_eri(int nbasis, const int* contr, const int* ltype, const double* alpha, const double* coefn, const double* xyz,
            double* eri) {
  Sys s = make_sys(0, nullptr, nullptr, nbasis, contr, ltype, 0);
  g_boys_fail = 0;
  erivector(s, alpha, coefn, reinterpret_cast<const V3<double>*>(xyz), eri);
  return g_boys_fail;
}

// a bounded sample of the ERI hot loop for CPU-baseline timing: contracted quartets [q0, q1) of the packed vector
int orc_eri_range(int nbasis, const int* contr, const int* ltype, const double* alpha, const double* coefn,
                  const double* xyz, long q0, long q1, double* out) {
  Sys s = make_sys(0, nullptr, nullptr, nbasis, contr, ltype, 0);
  long n = nbasis, len_vec = n * (n + 1) / 2;
  const V3<double>* r = reinterpret_cast<const V3<double>*>(xyz);
#pragma omp parallel for schedule(dynamic, 64)
  for (long q = q0; q < q1; ++q) {
    // invert q = x*len - x(x+1)/2 + y
    int x, y; tri_unindex(q, len_vec, &x, &y);
    int i, j, k, m; tri_unindex(x, n, &i, &j); tri_unindex(y, n, &k, &m);
    out[q - q0] = eri_contracted(s, alpha, coefn, r, i, j, k, m);
  }
  return 0;
}

// Energy.py:29-38 on its own (for Fock-build parity tests)
int orc_fock(int n, const double* H, const double* eri, const double* D, double* F) { fockmatrix(n, H, eri, D, F); return 0; }

// full RHF: alpha given as ln(alpha) (Task.py:135-136), raw coefficients; returns 0 = converged, 1 = SCF not converged
int orc_rhf(int natoms, const int* charges, const double* atoms, int nbasis, const int* contr, const int* ltype, int ne,
            const double* lnalpha, const double* coef, const double* xyz, int max_scf,
            double* E, int* niter, double* esteps, double* C, double* eps) {
  Sys s = make_sys(natoms, charges, atoms, nbasis, contr, ltype, ne);
  ScfOut so; g_boys_fail = 0;
  double e = rhfenergy<double>(s, lnalpha, coef, reinterpret_cast<const V3<double>*>(xyz), max_scf, true, &so, C, eps);
  *E = e; *niter = so.niter;
  for (int i = 0; i < so.niter && i < 1024; ++i) esteps[i] = so.esteps[i];
  if (g_boys_fail) return 2;
  return so.status ? 0 : 1;
}

// forward-mode gradient of the full RHF energy w.r.t. one argument group (Task.py:25-51):
//   argnum 0: ln(alpha) [nprim], 1: raw coef [nprim], 2: xyz [nbasis*3]
// returns 0 ok, 1 SCF not converged, 2 Boys failure, 3 too many directions, 4 degenerate eigen-coupling
int orc_rhf_grad(int natoms, const int* charges, const double* atoms, int nbasis, const int* contr, const int* ltype,
                 int ne, const double* lnalpha, const double* coef, const double* xyz, int max_scf, int argnum,
                 double* E, int* niter, double* grad) {
  Sys s = make_sys(natoms, charges, atoms, nbasis, contr, ltype, ne);
  int P = argnum == 2 ? 3 * nbasis : s.nprim;
  if (P > ORC_PMAX) return 3;
  g_ndir = P; g_boys_fail = 0; g_degenerate_coupling = 0;
  int rc = 0;
  {
    std::vector<Fwd> a(s.nprim), c(s.nprim);
    std::vector<V3<Fwd>> r(nbasis);
    for (int p = 0; p < s.nprim; ++p) { a[p] = Fwd(lnalpha[p]); c[p] = Fwd(coef[p]); }
    for (int i = 0; i < nbasis; ++i) for (int x = 0; x < 3; ++x) r[i][x] = Fwd(xyz[3 * i + x]);
    if (argnum == 0) for (int p = 0; p < P; ++p) a[p].d[p] = 1.0;
    if (argnum == 1) for (int p = 0; p < P; ++p) c[p].d[p] = 1.0;
    if (argnum == 2) for (int p = 0; p < P; ++p) r[p / 3][p % 3].d[p] = 1.0;
    ScfOut so;
    Fwd e = rhfenergy<Fwd>(s, a.data(), c.data(), r.data(), max_scf, true, &so, nullptr, nullptr);
    *E = e.v; *niter = so.niter;
    for (int p = 0; p < P; ++p) grad[p] = e.d[p];
    rc = so.status ? 0 : 1;
  }
  g_ndir = 0;
  if (g_boys_fail) return 2;
  if (g_degenerate_coupling) return 4;
  return rc;
}

// primitive-level probes (value + forward-mode partials) used to check the kernels' scalar math on the host
// grad = [d/d exps (4) | d/d coefs (4) | d/d centres (12)]
void orc_eri_primitive_grad(const double* ex, const double* co, const double* R, const int* l, double* value, double* grad) {
  int saved = g_ndir; g_ndir = 20;
  {
    Fwd e[4], c[4]; V3<Fwd> r[4];
    for (int i = 0; i < 4; ++i) {
      e[i] = Fwd(ex[i]); e[i].d[i] = 1.0;
      c[i] = Fwd(co[i]); c[i].d[4 + i] = 1.0;
      for (int x = 0; x < 3; ++x) { r[i][x] = Fwd(R[3 * i + x]); r[i][x].d[8 + 3 * i + x] = 1.0; }
    }
    Fwd v = eris_primitive(e[0], c[0], r[0], l[0], e[1], c[1], r[1], l[1], e[2], c[2], r[2], l[2], e[3], c[3], r[3], l[3]);
    *value = v.v;
    for (int p = 0; p < 20; ++p) grad[p] = v.d[p];
  }
  g_ndir = saved;
}

// out[3][11]: S,T,V value + partials wrt (alpha, beta, ca, cb, A[3], B[3])
void orc_one_electron_primitive_grad(double al, double ca, const double* A, int la, double be, double cb, const double* B,
                                     int lb, int natoms, const int* Z, const double* C, double* out) {
  int saved = g_ndir; g_ndir = 10;
  {
    Fwd a(al), b(be), c1(ca), c2(cb); V3<Fwd> Ad, Bd;
    a.d[0] = 1; b.d[1] = 1; c1.d[2] = 1; c2.d[3] = 1;
    for (int x = 0; x < 3; ++x) { Ad[x] = Fwd(A[x]); Ad[x].d[4 + x] = 1; Bd[x] = Fwd(B[x]); Bd[x].d[7 + x] = 1; }
    Fwd r[3];
    r[0] = overlap_primitive(a, c1, Ad, la, b, c2, Bd, lb);
    r[1] = kinetic_primitive(a, c1, Ad, la, b, c2, Bd, lb);
    r[2] = nuclear_primitive(a, c1, Ad, la, b, c2, Bd, lb, C, Z, natoms);
    for (int k = 0; k < 3; ++k) { out[11 * k] = r[k].v; for (int i = 0; i < 10; ++i) out[11 * k + 1 + i] = r[k].d[i]; }
  }
  g_ndir = saved;
}

}  // extern "C"
