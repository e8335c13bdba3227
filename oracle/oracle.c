/*
 * oracle.c -- plain C restatement of the reference's algorithms for the GPU hot path.
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product
 * (stencil_benchmarks_b200/).  Used by tests/ for parity checks at sizes where
 * the NumPy oracle (oracle/stencils.py) would need tens of GB of temporaries,
 * and by bench.py as the CPU baseline of kind "port" when oracle/_ref (the
 * reference's own OpenMP kernels) is unavailable.
 *
 * Each function follows one verify_stencil of
 * stencil_benchmarks/benchmarks_collection/stencils/base.py (cited below) with
 * the same operation order; compiled with -ffp-contract=off -fno-fast-math so
 * float64 results are bit-identical to NumPy's.  Parity is pinned through
 * tests/test_oracle.py, which checks these functions against the golden vectors
 * captured from the reference (tests/golden/).
 *
 * Conventions match include/sbench_b200.h: pointers address the FIRST INTERIOR
 * element, strides are in elements, i is unit stride.  OpenMP over (k, j) or
 * (j) -- the loops are embarrassingly parallel.
 */
#include <stddef.h>
#include <stdint.h>

#define IDX(i, j, k) ((int64_t)(i) + (int64_t)(j) * sy + (int64_t)(k) * sz)

#define DEFINE_ORACLE(T, SUFFIX)                                                                   \
  /* base.py:181-188 */                                                                            \
  void oracle_copy_##SUFFIX(const T* inp, T* out, int64_t nx, int64_t ny, int64_t nz, int64_t sy,  \
                            int64_t sz) {                                                          \
    _Pragma("omp parallel for collapse(2)") for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0;   \
                                                                                  j < ny; ++j) for \
        (int64_t i = 0; i < nx; ++i) out[IDX(i, j, k)] = inp[IDX(i, j, k)];                        \
  }                                                                                                \
  /* base.py:194-205 (sign = 0) and base.py:211-222 (sign = 1, symmetric) */                       \
  void oracle_average_##SUFFIX(const T* inp, T* out, int64_t nx, int64_t ny, int64_t nz,           \
                               int64_t sy, int64_t sz, int axis, int symmetric) {                  \
    const int64_t step = axis == 0 ? 1 : (axis == 1 ? sy : sz);                                    \
    _Pragma("omp parallel for collapse(2)") for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0;   \
                                                                                  j < ny; ++j) for \
        (int64_t i = 0; i < nx; ++i) {                                                             \
      const int64_t c = IDX(i, j, k);                                                              \
      out[c] = (inp[c + step] + (symmetric ? inp[c - step] : inp[c])) / 2;                         \
    }                                                                                              \
  }                                                                                                \
  /* base.py:238-254 */                                                                            \
  void oracle_laplacian_##SUFFIX(const T* inp, T* out, int64_t nx, int64_t ny, int64_t nz,         \
                                 int64_t sy, int64_t sz, int along) {                              \
    const int64_t steps[3] = {1, sy, sz};                                                          \
    _Pragma("omp parallel for collapse(2)") for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0;   \
                                                                                  j < ny; ++j) for \
        (int64_t i = 0; i < nx; ++i) {                                                             \
      const int64_t c = IDX(i, j, k);                                                              \
      T acc = 0;                                                                                   \
      for (int axis = 0; axis < 3; ++axis)                                                         \
        if (along & (1 << axis)) acc += 2 * inp[c] - inp[c + steps[axis]] - inp[c - steps[axis]];  \
      out[c] = acc;                                                                                \
    }                                                                                              \
  }                                                                                                \
  static inline T lap_##SUFFIX(const T* inp, int64_t c, int64_t sy) {                              \
    return 4 * inp[c] - (inp[c + 1] + inp[c - 1] + inp[c + sy] + inp[c - sy]);                     \
  }                                                                                                \
  static inline T lim_##SUFFIX(T flux, T delta) { return flux * delta > 0 ? 0 : flux; }            \
  /* base.py:284-307 */                                                                            \
  void oracle_hdiff_##SUFFIX(const T* inp, const T* coeff, T* out, int64_t nx, int64_t ny,         \
                             int64_t nz, int64_t sy, int64_t sz) {                                 \
    _Pragma("omp parallel for collapse(2)") for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0;   \
                                                                                  j < ny; ++j) for \
        (int64_t i = 0; i < nx; ++i) {                                                             \
      const int64_t c = IDX(i, j, k);                                                              \
      const T lap_c = lap_##SUFFIX(inp, c, sy);                                                    \
      const T flx = lim_##SUFFIX(lap_##SUFFIX(inp, c + 1, sy) - lap_c, inp[c + 1] - inp[c]);       \
      const T flx_m = lim_##SUFFIX(lap_c - lap_##SUFFIX(inp, c - 1, sy), inp[c] - inp[c - 1]);     \
      const T fly = lim_##SUFFIX(lap_##SUFFIX(inp, c + sy, sy) - lap_c, inp[c + sy] - inp[c]);     \
      const T fly_m = lim_##SUFFIX(lap_c - lap_##SUFFIX(inp, c - sy, sy), inp[c] - inp[c - sy]);   \
      out[c] = inp[c] - coeff[c] * (flx - flx_m + fly - fly_m);                                    \
    }                                                                                              \
  }                                                                                                \
  /* base.py:415-473; ccol/dcol: scratch with the fields' strides */                               \
  void oracle_vadv_##SUFFIX(const T* stage, const T* pos, const T* tens, T* tensstage,             \
                            const T* wcon, T* ccol, T* dcol, int64_t nx, int64_t ny, int64_t nz,   \
                            int64_t sy, int64_t sz, int ishift, int jshift) {                      \
    const T dtr_stage = (T)(3.0 / 20.0);                                                           \
    const T bet_m = (T)0.5, bet_p = (T)0.5;                                                        \
    const int64_t ws = ishift + jshift * sy;                                                       \
    _Pragma("omp parallel for") for (int64_t j = 0; j < ny; ++j) for (int64_t i = 0; i < nx;       \
                                                                      ++i) {                       \
      int64_t c = IDX(i, j, 0);                                                                    \
      {                                                                                            \
        const T gcv = (T)0.25 * (wcon[c + ws + sz] + wcon[c + sz]);                                \
        const T cs = gcv * bet_m;                                                                  \
        T cc = gcv * bet_p;                                                                        \
        const T b = dtr_stage - cc;                                                                \
        const T corr = -cs * (stage[c + sz] - stage[c]);                                           \
        T d = dtr_stage * pos[c] + tens[c] + tensstage[c] + corr;                                  \
        ccol[c] = cc / b;                                                                          \
        dcol[c] = d / b;                                                                           \
      }                                                                                            \
      for (int64_t k = 1; k < nz - 1; ++k) {                                                       \
        c = IDX(i, j, k);                                                                          \
        const T gav = (T)-0.25 * (wcon[c + ws] + wcon[c]);                                         \
        const T gcv = (T)0.25 * (wcon[c + ws + sz] + wcon[c + sz]);                                \
        const T as = gav * bet_m, cs = gcv * bet_m;                                                \
        const T a = gav * bet_p;                                                                   \
        T cc = gcv * bet_p;                                                                        \
        const T b = dtr_stage - a - cc;                                                            \
        const T corr = -as * (stage[c - sz] - stage[c]) - cs * (stage[c + sz] - stage[c]);         \
        T d = dtr_stage * pos[c] + tens[c] + tensstage[c] + corr;                                  \
        const T divided = (T)1.0 / (b - ccol[c - sz] * a);                                         \
        ccol[c] = cc * divided;                                                                    \
        dcol[c] = (d - dcol[c - sz] * a) * divided;                                                \
      }                                                                                            \
      c = IDX(i, j, nz - 1);                                                                       \
      {                                                                                            \
        const T gav = (T)-0.25 * (wcon[c + ws] + wcon[c]);                                         \
        const T as = gav * bet_m;                                                                  \
        const T a = gav * bet_p;                                                                   \
        const T b = dtr_stage - a;                                                                 \
        const T corr = -as * (stage[c - sz] - stage[c]);                                           \
        T d = dtr_stage * pos[c] + tens[c] + tensstage[c] + corr;                                  \
        dcol[c] = (d - dcol[c - sz] * a) / (b - ccol[c - sz] * a);                                 \
      }                                                                                            \
      T x = dcol[c];                                                                               \
      tensstage[c] = dtr_stage * (x - pos[c]);                                                     \
      for (int64_t k = nz - 2; k >= 0; --k) {                                                      \
        c = IDX(i, j, k);                                                                          \
        x = dcol[c] - ccol[c] * x;                                                                 \
        tensstage[c] = dtr_stage * (x - pos[c]);                                                   \
      }                                                                                            \
    }                                                                                              \
  }                                                                                                \
  /* stream/cuda_hip.j2:132-173: one round of copy, scale, add, triad */                           \
  void oracle_stream_round_##SUFFIX(T* a, T* b, T* c, uint64_t n, T scalar) {                      \
    _Pragma("omp parallel for") for (uint64_t i = 0; i < n; ++i) c[i] = a[i];                      \
    _Pragma("omp parallel for") for (uint64_t i = 0; i < n; ++i) b[i] = scalar * c[i];             \
    _Pragma("omp parallel for") for (uint64_t i = 0; i < n; ++i) c[i] = a[i] + b[i];               \
    _Pragma("omp parallel for") for (uint64_t i = 0; i < n; ++i) a[i] = b[i] + scalar * c[i];      \
  }                                                                                                \
  void oracle_stream_triad_##SUFFIX(T* a, const T* b, const T* c, uint64_t n, T scalar) {          \
    _Pragma("omp parallel for") for (uint64_t i = 0; i < n; ++i) a[i] = b[i] + scalar * c[i];      \
  }

DEFINE_ORACLE(double, f64)
DEFINE_ORACLE(float, f32)
