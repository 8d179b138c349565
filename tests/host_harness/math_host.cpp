// TEST HARNESS (not shipped, not used by the product): compiles the kernels' scalar math header
// diffiqult_b200/csrc/dq_math.cuh for the HOST so its formulas can be checked against the oracle in the
// build container, which has no GPU.  The product only ever runs this header inside CUDA kernels.
#include "../../diffiqult_b200/csrc/dq_math.cuh"
#include <cstring>
using namespace dq;

static BoysConst g_bc;
static bool g_init = false;
static void init() { if (!g_init) { boys_const_init(&g_bc); g_init = true; } }

template <bool LA, bool LB, bool LC, bool LD>
static double run(const PairPrim& bra, const PairPrim& ket, const int* l, double w, PairAdj* ba, PairAdj* ka, double* e) {
  return prim_quartet_axes<LA, LB, LC, LD, true>(bra, ket, l[0] - 1, l[1] - 1, l[2] - 1, l[3] - 1, g_bc, w, ba, ka, e);
}
template <bool LA, bool LB, bool LC, bool LD>
static double runv(const PairPrim& bra, const PairPrim& ket, const int* l) {
  return prim_quartet_axes<LA, LB, LC, LD, false>(bra, ket, l[0] - 1, l[1] - 1, l[2] - 1, l[3] - 1, g_bc, 0.0, nullptr, nullptr, nullptr);
}

// the same for a kernel of angular class L (orders 0..L; the regime boundaries of boys_ref depend on L): F,dF: [n][5]
template <int L> static void boys_class(int n, const double* t, double* F, double* dF) {
  for (int i = 0; i < n; ++i) boys_ref<L, true>(t[i], g_bc, F + 5 * i, dF + 5 * i);
}
extern "C" {

void hh_boys(int n, const double* t, double* F, double* dF) {  // F,dF: [n][5]
  init();
  for (int i = 0; i < n; ++i) boys_ref<4, true>(t[i], g_bc, F + 5 * i, dF + 5 * i);
}
void hh_boys_class(int L, int n, const double* t, double* F, double* dF) {
  init();
  switch (L) {
    case 0: boys_class<0>(n, t, F, dF); break;
    case 1: boys_class<1>(n, t, F, dF); break;
    case 2: boys_class<2>(n, t, F, dF); break;
    case 3: boys_class<3>(n, t, F, dF); break;
    default: boys_class<4>(n, t, F, dF); break;
  }
}
void hh_boys_value(int n, const double* t, double* F) {
  init(); double d[5];
  for (int i = 0; i < n; ++i) boys_ref<4, false>(t[i], g_bc, F + 5 * i, d);
}

// exps[4], coefs[4], centres[4][3], l[4] (0 s, 1..3 p) -> value (two ways) and 20 partials:
// grad = [d/d exps (4) | d/d coefs (4) | d/d centres (12)]
void hh_prim_quartet(const double* ex, const double* co, const double* R, const int* l, double* value, double* value_nograd,
                     double* grad) {
  init();
  PairPrim bra = make_pair(ex[0], co[0], R + 0, l[0], ex[1], co[1], R + 3, l[1]);
  PairPrim ket = make_pair(ex[2], co[2], R + 6, l[2], ex[3], co[3], R + 9, l[3]);
  PairAdj ba, ka; std::memset(&ba, 0, sizeof ba); std::memset(&ka, 0, sizeof ka);
  double e[4] = {0, 0, 0, 0};
  int code = (l[0] > 0) * 8 + (l[1] > 0) * 4 + (l[2] > 0) * 2 + (l[3] > 0);
  double v = 0, v2 = 0;
  switch (code) {
#define CASE(c, A, B, C, D) case c: v = run<A, B, C, D>(bra, ket, l, 1.0, &ba, &ka, e); v2 = runv<A, B, C, D>(bra, ket, l); break;
    CASE(0, false, false, false, false) CASE(1, false, false, false, true) CASE(2, false, false, true, false) CASE(3, false, false, true, true)
    CASE(4, false, true, false, false) CASE(5, false, true, false, true) CASE(6, false, true, true, false) CASE(7, false, true, true, true)
    CASE(8, true, false, false, false) CASE(9, true, false, false, true) CASE(10, true, false, true, false) CASE(11, true, false, true, true)
    CASE(12, true, true, false, false) CASE(13, true, true, false, true) CASE(14, true, true, true, false) CASE(15, true, true, true, true)
#undef CASE
  }
  *value = v; *value_nograd = v2;
  for (int i = 0; i < 20; ++i) grad[i] = 0.0;
  pair_backward(ex[0], co[0], R + 0, l[0], ex[1], co[1], R + 3, l[1], ba, e[0], e[1], grad + 0, grad + 1, grad + 4, grad + 5, grad + 8, grad + 11);
  pair_backward(ex[2], co[2], R + 6, l[2], ex[3], co[3], R + 9, l[3], ka, e[2], e[3], grad + 2, grad + 3, grad + 6, grad + 7, grad + 14, grad + 17);
}

// pair prefactor K of make_pair and the exponential-underflow marker the kernels test
void hh_pair_prefactor(double a, double ca, const double* A, double b, double cb, const double* B, double* K, int* underflowed) {
  const PairPrim p = make_pair(a, ca, A, 0, b, cb, B, 0);
  *K = p.K;
  *underflowed = pair_underflowed(p.K) ? 1 : 0;
}

// one primitive pair: S,T,V and their partials wrt (alpha, beta, ca, cb, A[3], B[3]) -> out[3][11]
void hh_one_electron_primitive(double al, double ca, const double* A, int la, double be, double cb, const double* B, int lb,
                               int natoms, const int* Z, const double* C, double* out) {
  init();
  typedef Dual<10> D;
  D a = seed<10>(al, 0), b = seed<10>(be, 1), c1 = seed<10>(ca, 2), c2 = seed<10>(cb, 3);
  D Ad[3], Bd[3];
  for (int x = 0; x < 3; ++x) { Ad[x] = seed<10>(A[x], 4 + x); Bd[x] = seed<10>(B[x], 7 + x); }
  D S, T, V;
  one_electron_primitive<D>(a, c1, Ad, la, b, c2, Bd, lb, natoms, Z, C, g_bc, &S, &T, &V);
  double s, t, v;
  one_electron_primitive<double>(al, ca, A, la, be, cb, B, lb, natoms, Z, C, g_bc, &s, &t, &v);
  const D* r[3] = {&S, &T, &V};
  const double pv[3] = {s, t, v};
  for (int k = 0; k < 3; ++k) { out[11 * k] = pv[k]; for (int i = 0; i < 10; ++i) out[11 * k + 1 + i] = r[k]->d[i]; if (r[k]->v != pv[k]) out[11 * k] = NAN; }
}
}
