/*
 * oracle/arap_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's ARAP solve path, reference inc/deform/arap.h (cheind/mesh-deform),
 * followed function by function; every function below cites the reference lines it restates.
 * It exists to CHECK the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs). Nothing in the product path (mesh_deform_b200/, inc/, include/) may
 * include, link or call it.
 *
 * Why a restatement and not the reference itself: arap.h needs Eigen (arap.h:14-17), which is not
 * installed in this image and cannot be fetched (no network), so the reference cannot be compiled.
 * Third-party arithmetic restated here from the published algorithms (Eigen, version unpinned by
 * the reference's cmake/FindEigen3.cmake:4-10):
 *   - SparseMatrix::setFromTriplets  (arap.h:238,336): compressed, inner indices ascending,
 *     duplicates summed in insertion order, tiny values kept.
 *   - JacobiSVD<Matrix3>             (arap.h:376): two-sided Jacobi, singular values sorted descending.
 *   - SimplicialLDLT                 (arap.h:336-339,420): see sparse_ldlt.c.
 *
 * PARITY PINS: the reference's own tests pin only computeCotanWeights / CSR assembly
 * (tests/test_cotan.cpp:26-53 -> tests/test_oracle_pins.py). No reference test calls deform(n>0),
 * so the local step, RHS, solve and final positions are "parity unpinned" by the reference; they
 * are cross-checked against an independent numpy/scipy restatement (oracle/numpy_ref.py) instead.
 *
 * Precision: REAL is the reference's PrecisionType (arap.h:49,53): double by default, float when
 * compiled with -DORACLE_F32 (symbols then carry an _f32 suffix).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <time.h>

#ifdef ORACLE_F32
#define REAL float
#define NAME(x) x##_f32
#define LDLT_REAL float
#define LDLT_F32
#define REAL_EPS FLT_EPSILON
#define REAL_MIN FLT_MIN
#else
#define REAL double
#define NAME(x) x##_f64
#define REAL_EPS DBL_EPSILON
#define REAL_MIN DBL_MIN
#endif
#include "sparse_ldlt.h"

enum { T_GEOMETRY, T_WEIGHTS, T_ASSEMBLY, T_FACTOR, T_LOCAL, T_RHS, T_SOLVE, T_WRITEBACK, T_COUNT };

typedef struct {
    /* arap.h:450-464, member for member */
    int nV, nF;
    int *faces;              /* _faces: 3 x F, [v0 v1 v2] per face (arap.h:452) */
    REAL *p, *pprime;        /* _p, _pprime: 3 x V, xyz per vertex (arap.h:451) */
    int *w_rowptr, *w_colidx;/* _edgeWeights: row-major sparse V x V (arap.h:453) */
    REAL *w_val;
    REAL *rot;               /* _rotations: V x (3x3), stored row-major r[3*a+b] = R(a,b) (arap.h:454) */
    int nFree;               /* _numberOfFreeVariables (arap.h:456) */
    int *freeIdx;            /* _freeIdxMap (arap.h:457) */
    unsigned char *isCon;    /* _constrainedLocations: key set (arap.h:458) */
    REAL *conLoc;            /*                         values */
    int *L_rowptr, *L_colidx;/* _L (arap.h:460), both triangles, columns ascending */
    REAL *L_val;
    REAL *bFixed, *b;        /* _bFixed, _b: 3 x nFree, xyz per free vertex (arap.h:461) */
    SparseLDLT *solver;      /* _solver (arap.h:462) */
    int dirty;               /* _dirty (arap.h:464) */
    double timers[T_COUNT];
} Oracle;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* std::max(a, b) semantics: returns a unless a < b (so max(c, NaN) == c, as at arap.h:208) */
static REAL rmax(REAL a, REAL b) { return (a < b) ? b : a; }

/* ------------------------------------------------------------------------------------------ */
/* setFromTriplets (Eigen; arap.h:238,336): CSR, ascending columns, duplicates summed in order */
/* ------------------------------------------------------------------------------------------ */
static void csr_from_triplets(int nRows, long nT, const int *ti, const int *tj, const REAL *tv,
                              int **rowptr_out, int **colidx_out, REAL **val_out) {
    int *cnt = (int *)calloc((size_t)nRows + 1, sizeof(int));
    for (long t = 0; t < nT; ++t) cnt[ti[t] + 1]++;
    for (int r = 0; r < nRows; ++r) cnt[r + 1] += cnt[r];
    int *bj = (int *)malloc(sizeof(int) * (size_t)(nT > 0 ? nT : 1));
    REAL *bv = (REAL *)malloc(sizeof(REAL) * (size_t)(nT > 0 ? nT : 1));
    int *cur = (int *)malloc(sizeof(int) * ((size_t)nRows + 1));
    memcpy(cur, cnt, sizeof(int) * ((size_t)nRows + 1));
    for (long t = 0; t < nT; ++t) { int q = cur[ti[t]]++; bj[q] = tj[t]; bv[q] = tv[t]; }
    free(cur);
    int *rowptr = (int *)calloc((size_t)nRows + 1, sizeof(int));
    int out = 0;
    for (int r = 0; r < nRows; ++r) {
        int lo = cnt[r], hi = cnt[r + 1];
        /* stable insertion sort by column keeps insertion order among duplicates */
        for (int a = lo + 1; a < hi; ++a) {
            int cj = bj[a]; REAL cv = bv[a]; int b = a - 1;
            while (b >= lo && bj[b] > cj) { bj[b + 1] = bj[b]; bv[b + 1] = bv[b]; --b; }
            bj[b + 1] = cj; bv[b + 1] = cv;
        }
        rowptr[r] = out;
        for (int a = lo; a < hi;) {
            int cj = bj[a]; REAL s = bv[a]; ++a;
            while (a < hi && bj[a] == cj) { s += bv[a]; ++a; }
            bj[out] = cj; bv[out] = s; ++out;          /* out <= lo always: in-place compaction */
        }
    }
    rowptr[nRows] = out;
    free(cnt);
    *rowptr_out = rowptr;
    *colidx_out = (int *)realloc(bj, sizeof(int) * (size_t)(out > 0 ? out : 1));
    *val_out = (REAL *)realloc(bv, sizeof(REAL) * (size_t)(out > 0 ? out : 1));
}

/* ------------------------------------------------------------------------------------------ */
/* JacobiSVD<Matrix3>, full U and V (Eigen; arap.h:376)                                       */
/* m, u, v are row-major 3x3; s[3] descending. m = u * diag(s) * v^T                          */
/* ------------------------------------------------------------------------------------------ */
static void rows_apply(REAL *a, int p, int q, REAL g00, REAL g01, REAL g10, REAL g11) {
    /* rows p,q of a <- G * [row p; row q] */
    for (int c = 0; c < 3; ++c) {
        REAL x = a[3 * p + c], y = a[3 * q + c];
        a[3 * p + c] = g00 * x + g01 * y;
        a[3 * q + c] = g10 * x + g11 * y;
    }
}
static void cols_apply(REAL *a, int p, int q, REAL g00, REAL g01, REAL g10, REAL g11) {
    /* cols p,q of a <- [col p, col q] * G */
    for (int r = 0; r < 3; ++r) {
        REAL x = a[3 * r + p], y = a[3 * r + q];
        a[3 * r + p] = x * g00 + y * g10;
        a[3 * r + q] = x * g01 + y * g11;
    }
}

static void jacobi_svd3(const REAL *m, REAL *u, REAL *s, REAL *v) {
    REAL work[9];
    REAL scale = 0;
    for (int i = 0; i < 9; ++i) scale = rmax(scale, (REAL)fabs((double)m[i]));
    if (scale == 0) scale = 1;
    for (int i = 0; i < 9; ++i) { work[i] = m[i] / scale; u[i] = v[i] = (i % 4 == 0) ? 1 : 0; }
    const REAL precision = 2 * REAL_EPS, considerAsZero = 2 * REAL_MIN;
    REAL maxDiag = rmax(rmax((REAL)fabs((double)work[0]), (REAL)fabs((double)work[4])), (REAL)fabs((double)work[8]));
    for (int sweep = 0; sweep < 100; ++sweep) {
        int finished = 1;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                REAL thr = rmax(considerAsZero, precision * maxDiag);
                if (fabs((double)work[3 * p + q]) <= thr && fabs((double)work[3 * q + p]) <= thr) continue;
                finished = 0;
                /* 2x2 block B = [[a,b],[c,d]] on rows/cols (p,q) */
                REAL a = work[3 * p + p], b = work[3 * p + q], c = work[3 * q + p], d = work[3 * q + q];
                /* left rotation Q = [[cs,sn],[-sn,cs]] with tan = (c-b)/(a+d) makes Q*B symmetric */
                REAL t = a + d, del = c - b, cs = 1, sn = 0;
                REAL h = (REAL)hypot((double)t, (double)del);
                if (h > 0 && fabs((double)del) > REAL_MIN) { cs = t / h; sn = del / h; }
                REAL x = cs * a + sn * c, y = cs * b + sn * d, z = -sn * b + cs * d;
                /* symmetric Schur: J = [[jc,js],[-js,jc]] diagonalises [[x,y],[y,z]] */
                REAL jc = 1, js = 0;
                if (fabs((double)y) > REAL_MIN) {
                    REAL tau = (z - x) / (2 * y);
                    REAL w = (REAL)sqrt(1.0 + (double)tau * (double)tau);
                    REAL tt = (tau >= 0) ? 1 / (tau + w) : 1 / (tau - w);
                    jc = 1 / (REAL)sqrt(1.0 + (double)tt * (double)tt);
                    js = tt * jc;
                }
                /* left factor G = J^T * Q */
                REAL g00 = jc * cs + js * sn, g01 = jc * sn - js * cs;
                REAL g10 = js * cs - jc * sn, g11 = js * sn + jc * cs;
                rows_apply(work, p, q, g00, g01, g10, g11);          /* work <- G * work     */
                cols_apply(u, p, q, g00, g10, g01, g11);             /* U    <- U * G^T      */
                cols_apply(work, p, q, jc, js, -js, jc);             /* work <- work * J     */
                cols_apply(v, p, q, jc, js, -js, jc);                /* V    <- V * J        */
                maxDiag = rmax(maxDiag, rmax((REAL)fabs((double)work[3 * p + p]), (REAL)fabs((double)work[3 * q + q])));
            }
        if (finished) break;
    }
    for (int i = 0; i < 3; ++i) {
        REAL a = work[4 * i];
        if (a < 0) { for (int r = 0; r < 3; ++r) u[3 * r + i] = -u[3 * r + i]; a = -a; }
        s[i] = a * scale;
    }
    for (int i = 0; i < 3; ++i) {             /* sort descending, swapping columns of U and V */
        int best = i;
        for (int k = i + 1; k < 3; ++k) if (s[k] > s[best]) best = k;
        if (best != i) {
            REAL ts = s[i]; s[i] = s[best]; s[best] = ts;
            for (int r = 0; r < 3; ++r) {
                REAL tu = u[3 * r + i]; u[3 * r + i] = u[3 * r + best]; u[3 * r + best] = tu;
                REAL tv = v[3 * r + i]; v[3 * r + i] = v[3 * r + best]; v[3 * r + best] = tv;
            }
        }
    }
}

static REAL det3(const REAL *a) {
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}

/* ------------------------------------------------------------------------------------------ */
/* the class                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* arap.h:66-70 ctor + :149-155 initializeMeshTopology */
Oracle *NAME(oracle_create)(int nV, int nF, const int *faces) {
    Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
    o->nV = nV; o->nF = nF;
    o->faces = (int *)malloc(sizeof(int) * 3 * (size_t)(nF > 0 ? nF : 1));
    memcpy(o->faces, faces, sizeof(int) * 3 * (size_t)nF);
    o->isCon = (unsigned char *)calloc((size_t)(nV > 0 ? nV : 1), 1);
    o->conLoc = (REAL *)calloc(3 * (size_t)(nV > 0 ? nV : 1), sizeof(REAL));
    o->dirty = 1;
    return o;
}

static void free_system(Oracle *o) {
    free(o->L_rowptr); free(o->L_colidx); free(o->L_val); free(o->bFixed); free(o->b);
    o->L_rowptr = o->L_colidx = NULL; o->L_val = o->bFixed = o->b = NULL;
    if (o->solver) { ldlt_free(o->solver); o->solver = NULL; }
}

void NAME(oracle_destroy)(Oracle *o) {
    if (!o) return;
    free_system(o);
    free(o->faces); free(o->p); free(o->pprime); free(o->w_rowptr); free(o->w_colidx); free(o->w_val);
    free(o->rot); free(o->freeIdx); free(o->isCon); free(o->conLoc);
    free(o);
}

/* arap.h:81-85 setConstraint: overwrite-or-insert, cast to Scalar, mark dirty */
void NAME(oracle_set_constraint)(Oracle *o, int vidx, const double *xyz) {
    o->isCon[vidx] = 1;
    for (int d = 0; d < 3; ++d) o->conLoc[3 * (size_t)vidx + d] = (REAL)xyz[d];
    o->dirty = 1;
}

/* arap.h:162-168 initializeMeshGeometry: rest pose = CURRENT mesh positions, cast to Scalar */
static void initializeMeshGeometry(Oracle *o, const void *mesh, int meshIsFloat) {
    size_t n3 = 3 * (size_t)o->nV;
    o->p = (REAL *)realloc(o->p, sizeof(REAL) * (n3 ? n3 : 1));
    o->pprime = (REAL *)realloc(o->pprime, sizeof(REAL) * (n3 ? n3 : 1));
    for (size_t i = 0; i < n3; ++i)
        o->p[i] = meshIsFloat ? (REAL)((const float *)mesh)[i] : (REAL)((const double *)mesh)[i];
    memcpy(o->pprime, o->p, sizeof(REAL) * n3);
}

/* arap.h:182-239 computeCotanWeights (+ :435-440 undirectedEdge) */
static void computeCotanWeights(Oracle *o) {
    long nT = 6 * (long)o->nF;
    int *ti = (int *)malloc(sizeof(int) * (size_t)(nT ? nT : 1));
    int *tj = (int *)malloc(sizeof(int) * (size_t)(nT ? nT : 1));
    REAL *tv = (REAL *)malloc(sizeof(REAL) * (size_t)(nT ? nT : 1));
    for (int f = 0; f < o->nF; ++f) {
        const int *vids = o->faces + 3 * (size_t)f;
        const REAL *v0 = o->p + 3 * (size_t)vids[0], *v1 = o->p + 3 * (size_t)vids[1], *v2 = o->p + 3 * (size_t)vids[2];
        /* squaredNorm() of a fixed-size 3-vector: Eigen's unrolled reduction is x^2 + (y^2 + z^2) */
        REAL ea[3], eb[3], ec[3];
        for (int d = 0; d < 3; ++d) { ea[d] = v1[d] - v0[d]; eb[d] = v2[d] - v1[d]; ec[d] = v0[d] - v2[d]; }
        REAL l0 = ea[0] * ea[0] + (ea[1] * ea[1] + ea[2] * ea[2]);
        REAL l1 = eb[0] * eb[0] + (eb[1] * eb[1] + eb[2] * eb[2]);
        REAL l2 = ec[0] * ec[0] + (ec[1] * ec[1] + ec[2] * ec[2]);
        l0 = rmax((REAL)1e-8, l0); l1 = rmax((REAL)1e-8, l1); l2 = rmax((REAL)1e-8, l2);      /* :199-201 */
        l0 = (REAL)sqrt((double)l0); l1 = (REAL)sqrt((double)l1); l2 = (REAL)sqrt((double)l2); /* :203-205 */
        const REAL semip = (REAL)0.5 * (l0 + l1 + l2);                                         /* :207 */
        REAL heron = semip * (semip - l0) * (semip - l1) * (semip - l2);
        const REAL area = rmax((REAL)1e-8, (REAL)sqrt((double)heron));                         /* :208 (sqrt(<0) = NaN -> max picks 1e-8, as std::max(1e-8, NaN)) */
        const REAL denom = (REAL)1.0 / ((REAL)4.0 * area);                                     /* :210 */
        REAL cot0 = (-l0 * l0 + l1 * l1 + l2 * l2) * denom;                                    /* :212-214 */
        REAL cot1 = (l0 * l0 - l1 * l1 + l2 * l2) * denom;
        REAL cot2 = (l0 * l0 + l1 * l1 - l2 * l2) * denom;
        cot0 = rmax((REAL)1e-10, cot0); cot1 = rmax((REAL)1e-10, cot1); cot2 = rmax((REAL)1e-10, cot2); /* :216-218 */
        const int va[3] = {vids[0], vids[1], vids[2]}, vb[3] = {vids[1], vids[2], vids[0]};    /* :220-222 */
        const REAL cw[3] = {cot0 * (REAL)0.5, cot1 * (REAL)0.5, cot2 * (REAL)0.5};
        for (int k = 0; k < 3; ++k) {
            int lo = va[k] > vb[k] ? vb[k] : va[k], hi = va[k] > vb[k] ? va[k] : vb[k];
            ti[6 * (long)f + k] = lo; tj[6 * (long)f + k] = hi; tv[6 * (long)f + k] = cw[k];              /* :225-227 */
            ti[6 * (long)f + 3 + k] = hi; tj[6 * (long)f + 3 + k] = lo; tv[6 * (long)f + 3 + k] = cw[k];  /* :230-232 */
        }
    }
    free(o->w_rowptr); free(o->w_colidx); free(o->w_val);
    csr_from_triplets(o->nV, nT, ti, tj, tv, &o->w_rowptr, &o->w_colidx, &o->w_val);              /* :235-238 */
    free(ti); free(tj); free(tv);
}

/* arap.h:246-249 initializeRotations */
static void initializeRotations(Oracle *o) {
    size_t n = (size_t)o->nV;
    o->rot = (REAL *)realloc(o->rot, sizeof(REAL) * 9 * (n ? n : 1));
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 9; ++k) o->rot[9 * i + k] = (k % 4 == 0) ? 1 : 0;
}

/* arap.h:261-272 initializeFreeVariableMapping */
static void initializeFreeVariableMapping(Oracle *o) {
    o->freeIdx = (int *)realloc(o->freeIdx, sizeof(int) * (size_t)(o->nV ? o->nV : 1));
    int freeIdx = 0;
    for (int i = 0; i < o->nV; ++i) o->freeIdx[i] = o->isCon[i] ? -1 : freeIdx++;
    o->nFree = freeIdx;
}

/* arap.h:277-281 initializeConstraints */
static void initializeConstraints(Oracle *o) {
    for (int i = 0; i < o->nV; ++i)
        if (o->isCon[i]) memcpy(o->pprime + 3 * (size_t)i, o->conLoc + 3 * (size_t)i, sizeof(REAL) * 3);
}

/* arap.h:292-340 setupLinearSystem */
static int setupLinearSystem(Oracle *o) {
    double t0 = now_s();
    free_system(o);
    int nf = o->nFree;
    o->bFixed = (REAL *)calloc(3 * (size_t)(nf ? nf : 1), sizeof(REAL));
    o->b = (REAL *)calloc(3 * (size_t)(nf ? nf : 1), sizeof(REAL));
    long cap = (long)o->w_rowptr[o->nV] * 2 + 1;
    int *ti = (int *)malloc(sizeof(int) * (size_t)cap), *tj = (int *)malloc(sizeof(int) * (size_t)cap);
    REAL *tv = (REAL *)malloc(sizeof(REAL) * (size_t)cap);
    long nT = 0;
    for (int vi = 0; vi < o->nV; ++vi) {
        int iidx = o->freeIdx[vi];
        if (iidx == -1) continue;                                                  /* :314-317 */
        for (int k = o->w_rowptr[vi]; k < o->w_rowptr[vi + 1]; ++k) {
            int vj = o->w_colidx[k];
            int jidx = o->freeIdx[vj];
            const REAL wij = o->w_val[k];
            if (jidx == -1) {
                for (int d = 0; d < 3; ++d) o->bFixed[3 * (size_t)iidx + d] += wij * o->conLoc[3 * (size_t)vj + d]; /* :327 */
            } else {
                ti[nT] = iidx; tj[nT] = jidx; tv[nT] = -wij; ++nT;                  /* :329 */
            }
            ti[nT] = iidx; tj[nT] = iidx; tv[nT] = wij; ++nT;                       /* :332 */
        }
    }
    csr_from_triplets(nf, nT, ti, tj, tv, &o->L_rowptr, &o->L_colidx, &o->L_val);   /* :336 */
    free(ti); free(tj); free(tv);
    o->timers[T_ASSEMBLY] += now_s() - t0;

    t0 = now_s();
    double *coords = (double *)malloc(sizeof(double) * 3 * (size_t)(nf ? nf : 1));
    for (int vi = 0; vi < o->nV; ++vi)
        if (o->freeIdx[vi] >= 0)
            for (int d = 0; d < 3; ++d) coords[3 * (size_t)o->freeIdx[vi] + d] = (double)o->p[3 * (size_t)vi + d];
    o->solver = ldlt_factor(nf, o->L_rowptr, o->L_colidx, o->L_val, coords);        /* :337 */
    free(coords);
    o->timers[T_FACTOR] += now_s() - t0;
    return o->solver && o->solver->ok;                                              /* :339 */
}

/* arap.h:354-384 estimateRotations (local step) */
static void estimateRotations(Oracle *o) {
    for (int vi = 0; vi < o->nV; ++vi) {
        const REAL *pi = o->p + 3 * (size_t)vi, *ppi = o->pprime + 3 * (size_t)vi;
        REAL cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = o->w_rowptr[vi]; k < o->w_rowptr[vi + 1]; ++k) {
            const REAL wij = o->w_val[k];
            const REAL *pj = o->p + 3 * (size_t)o->w_colidx[k], *ppj = o->pprime + 3 * (size_t)o->w_colidx[k];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    cov[3 * a + b] += wij * (pi[a] - pj[a]) * (ppi[b] - ppj[b]);   /* :373 */
        }
        REAL u[9], v[9], s[3];
        jacobi_svd3(cov, u, s, v);                                                  /* :376 */
        /* v * ut */
        REAL vut[9];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                REAL acc = 0;
                for (int c = 0; c < 3; ++c) acc += v[3 * a + c] * u[3 * b + c];
                vut[3 * a + b] = acc;
            }
        const REAL flip = det3(vut);                                                /* :381 */
        REAL *r = o->rot + 9 * (size_t)vi;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)                                             /* :382 v * id * ut */
                r[3 * a + b] = v[3 * a + 0] * u[3 * b + 0] + v[3 * a + 1] * u[3 * b + 1] + (v[3 * a + 2] * flip) * u[3 * b + 2];
    }
}

/* arap.h:393-430 estimatePositions (global step) */
static void estimatePositions(Oracle *o) {
    double t0 = now_s();
    int nf = o->nFree;
    memcpy(o->b, o->bFixed, sizeof(REAL) * 3 * (size_t)nf);                         /* :395 */
    for (int vi = 0; vi < o->nV; ++vi) {
        int iidx = o->freeIdx[vi];
        if (iidx == -1) continue;
        const REAL *ri = o->rot + 9 * (size_t)vi;
        for (int k = o->w_rowptr[vi]; k < o->w_rowptr[vi + 1]; ++k) {
            int vj = o->w_colidx[k];
            const REAL wij = o->w_val[k];
            const REAL *rj = o->rot + 9 * (size_t)vj;
            REAL d[3];
            for (int a = 0; a < 3; ++a) d[a] = o->p[3 * (size_t)vi + a] - o->p[3 * (size_t)vj + a];
            for (int a = 0; a < 3; ++a) {                                           /* :410-412 */
                REAL acc = 0;
                for (int c = 0; c < 3; ++c) acc += (ri[3 * a + c] + rj[3 * a + c]) * d[c];
                o->b[3 * (size_t)iidx + a] += acc * wij * (REAL)0.5;
            }
        }
    }
    o->timers[T_RHS] += now_s() - t0;

    t0 = now_s();
    LDLT_REAL *u = (LDLT_REAL *)malloc(sizeof(LDLT_REAL) * (size_t)(nf ? nf : 1));
    for (int d = 0; d < 3; ++d) {                                                   /* :419-429 */
        for (int i = 0; i < nf; ++i) u[i] = o->b[3 * (size_t)i + d];
        ldlt_solve(o->solver, u);
        int idx = 0;
        for (int vi = 0; vi < o->nV; ++vi)
            if (o->freeIdx[vi] != -1) o->pprime[3 * (size_t)vi + d] = (REAL)u[idx++];
    }
    free(u);
    o->timers[T_SOLVE] += now_s() - t0;
}

/* arap.h:101-138 deform. Returns 1 (true) / 0 (false). mesh: V x 3 positions, float or double. */
int NAME(oracle_deform)(Oracle *o, void *mesh, int meshIsFloat, int numberOfIterations) {
    if (o->dirty) {
        double t0 = now_s();
        initializeMeshGeometry(o, mesh, meshIsFloat);
        o->timers[T_GEOMETRY] += now_s() - t0;
        t0 = now_s();
        computeCotanWeights(o);
        o->timers[T_WEIGHTS] += now_s() - t0;
        initializeRotations(o);
        initializeFreeVariableMapping(o);
        initializeConstraints(o);
        if (o->nFree == o->nV) return 1;                                            /* :113-114 */
        if (!setupLinearSystem(o)) return 0;                                        /* :116-117 */
        o->dirty = 0;
    }
    for (int i = 0; i < numberOfIterations; ++i) {
        double t0 = now_s();
        estimateRotations(o);
        o->timers[T_LOCAL] += now_s() - t0;
        estimatePositions(o);
    }
    double t0 = now_s();
    size_t n3 = 3 * (size_t)o->nV;                                                  /* :133-135 */
    if (meshIsFloat) for (size_t i = 0; i < n3; ++i) ((float *)mesh)[i] = (float)o->pprime[i];
    else for (size_t i = 0; i < n3; ++i) ((double *)mesh)[i] = (double)o->pprime[i];
    o->timers[T_WRITEBACK] += now_s() - t0;
    return 1;
}

/*
 * ARAP energy. NOT in the reference (it has no energy function and no convergence test); defined
 * here as Sorkine & Alexa 2007 eq. (3)/(7) with unit cell weights, over all directed CSR entries,
 * with the rotations of the last local step, accumulated in double:
 *   E = sum_i sum_{j in N(i)} w_ij || (p'_i - p'_j) - R_i (p_i - p_j) ||^2
 */
double NAME(oracle_energy)(const Oracle *o) {
    double E = 0;
    if (!o->w_rowptr || !o->rot) return 0;
    for (int vi = 0; vi < o->nV; ++vi) {
        const REAL *ri = o->rot + 9 * (size_t)vi;
        for (int k = o->w_rowptr[vi]; k < o->w_rowptr[vi + 1]; ++k) {
            int vj = o->w_colidx[k];
            double d[3], dp[3];
            for (int a = 0; a < 3; ++a) {
                d[a] = (double)o->p[3 * (size_t)vi + a] - (double)o->p[3 * (size_t)vj + a];
                dp[a] = (double)o->pprime[3 * (size_t)vi + a] - (double)o->pprime[3 * (size_t)vj + a];
            }
            double e2 = 0;
            for (int a = 0; a < 3; ++a) {
                double rd = (double)ri[3 * a] * d[0] + (double)ri[3 * a + 1] * d[1] + (double)ri[3 * a + 2] * d[2];
                e2 += (dp[a] - rd) * (dp[a] - rd);
            }
            E += (double)o->w_val[k] * e2;
        }
    }
    return E;
}

/* ------------------------------------------------------------------------------------------ */
/* inspection (what tests/accessor.h:16-22 PrivateAccessor does, widened to all members)      */
/* ------------------------------------------------------------------------------------------ */
int NAME(oracle_nnz)(const Oracle *o) { return o->w_rowptr ? o->w_rowptr[o->nV] : 0; }
int NAME(oracle_nfree)(const Oracle *o) { return o->nFree; }
int NAME(oracle_dirty)(const Oracle *o) { return o->dirty; }
long long NAME(oracle_factor_nnz)(const Oracle *o) { return o->solver ? (long long)o->solver->lnz : 0; }
int NAME(oracle_L_nnz)(const Oracle *o) { return o->L_rowptr ? o->L_rowptr[o->nFree] : 0; }

void NAME(oracle_get_csr)(const Oracle *o, int *rowptr, int *colidx, REAL *val) {
    memcpy(rowptr, o->w_rowptr, sizeof(int) * ((size_t)o->nV + 1));
    memcpy(colidx, o->w_colidx, sizeof(int) * (size_t)o->w_rowptr[o->nV]);
    memcpy(val, o->w_val, sizeof(REAL) * (size_t)o->w_rowptr[o->nV]);
}
void NAME(oracle_get_L)(const Oracle *o, int *rowptr, int *colidx, REAL *val) {
    memcpy(rowptr, o->L_rowptr, sizeof(int) * ((size_t)o->nFree + 1));
    memcpy(colidx, o->L_colidx, sizeof(int) * (size_t)o->L_rowptr[o->nFree]);
    memcpy(val, o->L_val, sizeof(REAL) * (size_t)o->L_rowptr[o->nFree]);
}
void NAME(oracle_get_free_map)(const Oracle *o, int *freeIdx) { memcpy(freeIdx, o->freeIdx, sizeof(int) * (size_t)o->nV); }
void NAME(oracle_get_rotations)(const Oracle *o, REAL *rot) { memcpy(rot, o->rot, sizeof(REAL) * 9 * (size_t)o->nV); }
void NAME(oracle_get_rest)(const Oracle *o, REAL *p) { memcpy(p, o->p, sizeof(REAL) * 3 * (size_t)o->nV); }
void NAME(oracle_get_positions)(const Oracle *o, REAL *pp) { memcpy(pp, o->pprime, sizeof(REAL) * 3 * (size_t)o->nV); }
void NAME(oracle_get_bfixed)(const Oracle *o, REAL *bf) { memcpy(bf, o->bFixed, sizeof(REAL) * 3 * (size_t)o->nFree); }
void NAME(oracle_get_b)(const Oracle *o, REAL *b) { memcpy(b, o->b, sizeof(REAL) * 3 * (size_t)o->nFree); }
void NAME(oracle_get_timers)(const Oracle *o, double *t) { memcpy(t, o->timers, sizeof(double) * T_COUNT); }
void NAME(oracle_reset_timers)(Oracle *o) { memset(o->timers, 0, sizeof(o->timers)); }

/* one 3x3 SVD-to-rotation, exposed so tests can probe the local step's kernel of arithmetic */
void NAME(oracle_rotation_from_covariance)(const REAL *cov, REAL *r) {
    REAL u[9], v[9], s[3], vut[9];
    jacobi_svd3(cov, u, s, v);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            REAL acc = 0;
            for (int c = 0; c < 3; ++c) acc += v[3 * a + c] * u[3 * b + c];
            vut[3 * a + b] = acc;
        }
    const REAL flip = det3(vut);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            r[3 * a + b] = v[3 * a + 0] * u[3 * b + 0] + v[3 * a + 1] * u[3 * b + 1] + (v[3 * a + 2] * flip) * u[3 * b + 2];
}
void NAME(oracle_svd3)(const REAL *m, REAL *u, REAL *s, REAL *v) { jacobi_svd3(m, u, s, v); }
