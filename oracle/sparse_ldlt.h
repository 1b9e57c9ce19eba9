/*
 * oracle/sparse_ldlt.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped in the product path).
 *
 * A from-scratch simplicial sparse LDL^T (no pivoting) used by the oracle where the reference
 * calls Eigen::SimplicialLDLT<SparseMatrix<Scalar>> (reference inc/deform/arap.h:336-339 compute,
 * :420 solve, :462 member). Eigen is an un-vendored dependency (version unpinned by
 * cmake/FindEigen3.cmake:4-10); its published algorithm is: fill-reducing symmetric permutation,
 * elimination-tree symbolic analysis, up-looking numeric LDL^T on the lower triangle, and
 * solve = P, L^-1, D^-1, L^-T, P^T. Mathematically this is an exact SPD solve; the ordering only
 * changes fill and rounding (~1e-13 relative), so a nested-dissection ordering stands in for AMD.
 *
 * The numeric type is LDLT_REAL (double by default). The file is compiled a second time with
 * -DLDLT_REAL=float -DLDLT_F32 so that the float oracle solves in float like
 * SimplicialLDLT<SparseMatrix<float>> does; the float build's symbols carry an _f32 suffix.
 */
#ifndef ORACLE_SPARSE_LDLT_H
#define ORACLE_SPARSE_LDLT_H

#include <stdint.h>

#ifndef LDLT_REAL
#define LDLT_REAL double
#endif
#ifdef LDLT_F32
#define SparseLDLT SparseLDLT_f32
#define ldlt_factor ldlt_factor_f32
#define ldlt_solve ldlt_solve_f32
#define ldlt_free ldlt_free_f32
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SparseLDLT {
    int n;
    int64_t lnz;      /* strictly-lower nonzeros of L */
    int *perm;        /* perm[new] = old */
    int *iperm;       /* iperm[old] = new */
    int *parent;      /* elimination tree */
    int64_t *Lp;      /* n+1 column pointers of L (CSC, strictly lower, unit diagonal implied) */
    int *Li;
    LDLT_REAL *Lx;
    LDLT_REAL *D;
    LDLT_REAL *work;  /* n values */
    int ok;           /* 1 if the factorisation succeeded (no zero pivot) */
} SparseLDLT;

/*
 * Factor the symmetric matrix given in CSR (or CSC; it is symmetric) with BOTH triangles stored:
 *   rowptr[n+1], colidx[nnz], val[nnz].  coords (n x 3, may be NULL) drive the nested-dissection
 * ordering; with NULL a plain natural ordering is used (fine for tiny systems).
 * Returns NULL on allocation failure; check ->ok for numerical success
 * (Eigen's info()==Success, arap.h:339).
 */
SparseLDLT *ldlt_factor(int n, const int *rowptr, const int *colidx, const LDLT_REAL *val,
                        const double *coords);

/* Solve A x = b in place for one right-hand side (arap.h:420 solves one coordinate at a time). */
void ldlt_solve(SparseLDLT *F, LDLT_REAL *x);

void ldlt_free(SparseLDLT *F);

#ifdef __cplusplus
}
#endif
#endif
