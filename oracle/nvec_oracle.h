/* nvec_oracle.h -- TEST INFRASTRUCTURE ONLY (parity oracle), never shipped.
 *
 * Plain-C, single-threaded CPU restatement of the arithmetic of the reference's
 * nvector_serial (SUNDIALS 7.5.0, src/nvector/serial/nvector_serial.c) on raw
 * double arrays.  Every function cites the reference lines it follows.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (sundials_b200/) never does.
 *
 * PARITY PINNED: tests/test_oracle_cpu.py checks every function here bit-for-bit
 * against (a) the unmodified reference nvector_serial compiled from
 * /root/reference into oracle/_ref/lib/libsundials_ref.so (when present) and
 * (b) the committed golden vectors tests/golden/nvec_golden.npz that were
 * generated from that reference build by tests/golden/make_golden.py, plus the
 * known answers of the reference's own test/unit_tests/nvector/test_nvector.c.
 *
 * Conventions: sunrealtype = double, sunindextype = int64_t (reference defaults).
 * Vector arrays are `double* const*` (array of data pointers); "array identity"
 * aliasing of the reference (Y == Z as N_Vector* arrays) is expressed by passing
 * the same pointer-array address.  2-D arrays Y[j][i] (j = sum index, i = vector
 * index; nvector_serial.c:1387,1504) are passed flattened as P[j*nvec + i] with a
 * flag telling whether the reference call had Y == Z / X[0] == Z.
 */
#ifndef NVEC_ORACLE_H
#define NVEC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t orc_index;

/* streaming ops */
void orc_linear_sum(double a, const double* x, double b, const double* y, double* z, orc_index n);
void orc_const(double c, double* z, orc_index n);
void orc_prod(const double* x, const double* y, double* z, orc_index n);
void orc_div(const double* x, const double* y, double* z, orc_index n);
void orc_scale(double c, const double* x, double* z, orc_index n);
void orc_abs(const double* x, double* z, orc_index n);
void orc_inv(const double* x, double* z, orc_index n);
void orc_add_const(const double* x, double b, double* z, orc_index n);
void orc_compare(double c, const double* x, double* z, orc_index n);

/* reductions */
double orc_dot_prod(const double* x, const double* y, orc_index n);
double orc_max_norm(const double* x, orc_index n);
double orc_wsqr_sum(const double* x, const double* w, orc_index n);
double orc_wsqr_sum_mask(const double* x, const double* w, const double* id, orc_index n);
double orc_wrms_norm(const double* x, const double* w, orc_index n);
double orc_wrms_norm_mask(const double* x, const double* w, const double* id, orc_index n);
double orc_min(const double* x, orc_index n);
double orc_wl2_norm(const double* x, const double* w, orc_index n);
double orc_l1_norm(const double* x, orc_index n);
int    orc_inv_test(const double* x, double* z, orc_index n);
int    orc_constr_mask(const double* c, const double* x, double* m, orc_index n);
double orc_min_quotient(const double* num, const double* denom, orc_index n);

/* fused ops */
int orc_linear_combination(int nvec, const double* c, double* const* X, double* z, orc_index n);
int orc_scale_add_multi(int nvec, const double* a, const double* x, double* const* Y,
                        double* const* Z, orc_index n);
int orc_dot_prod_multi(int nvec, const double* x, double* const* Y, double* dotprods, orc_index n);

/* vector-array ops */
int orc_linear_sum_vector_array(int nvec, double a, double* const* X, double b, double* const* Y,
                                double* const* Z, orc_index n);
int orc_scale_vector_array(int nvec, const double* c, double* const* X, double* const* Z, orc_index n);
int orc_const_vector_array(int nvec, double c, double* const* Z, orc_index n);
int orc_wrms_norm_vector_array(int nvec, double* const* X, double* const* W, double* nrm, orc_index n);
int orc_wrms_norm_mask_vector_array(int nvec, double* const* X, double* const* W, const double* id,
                                    double* nrm, orc_index n);
int orc_scale_add_multi_vector_array(int nvec, int nsum, const double* a, double* const* X,
                                     double* const* Y, double* const* Z, orc_index n);
int orc_linear_combination_vector_array(int nvec, int nsum, const double* c, double* const* X,
                                        double* const* Z, orc_index n);

/* multi-rank (MPIPlusX / MPIManyVector) reduction semantics on host scalars:
 * combine per-rank local results exactly as nvector_manyvector.c does. */
double orc_mpi_wrms_from_local(const double* local_sqrsums, int nranks, orc_index global_n);

/* benchmark input generator (mirrors the reference benchmark's LCG, fixed seed) */
void orc_fill_uniform(double* x, orc_index n, uint32_t seed, double lo, double hi);

#ifdef __cplusplus
}
#endif
#endif
