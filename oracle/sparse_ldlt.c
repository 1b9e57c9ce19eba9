/*
 * oracle/sparse_ldlt.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). See sparse_ldlt.h.
 *
 * Stands in for Eigen::SimplicialLDLT (reference inc/deform/arap.h:336-339, :420).
 * Ordering: recursive coordinate bisection with one-sided vertex separators (nested dissection).
 * Factorisation: elimination tree + column counts, then the up-looking row-by-row LDL^T
 * (each row of L is the solution of a sparse triangular system whose pattern is read off the
 * elimination tree), as described in Davis, "Direct Methods for Sparse Linear Systems", ch. 4.
 */
#include "sparse_ldlt.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------------ */
/* nested dissection                                                                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct { double key; int v; } KeyedVertex;

typedef struct {
    const int *rp, *ci;
    const double *xyz;
    int *mark;          /* region stamp per vertex */
    int stamp;
    int *order;         /* order[pos] = vertex */
    int pos;
    KeyedVertex *kv;    /* scratch, n entries */
} NdCtx;

static int kv_less(const KeyedVertex *a, const KeyedVertex *b) {
    if (a->key != b->key) return a->key < b->key;
    return a->v < b->v;
}

/* Rearrange kv[0..n) so that kv[0..k) are the k smallest (quickselect, median-of-three). */
static void kv_select(KeyedVertex *kv, int n, int k) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        KeyedVertex a = kv[lo], b = kv[mid], c = kv[hi], piv;
        if (kv_less(&a, &b)) {
            if (kv_less(&b, &c)) piv = b; else piv = kv_less(&a, &c) ? c : a;
        } else {
            if (kv_less(&a, &c)) piv = a; else piv = kv_less(&b, &c) ? c : b;
        }
        int i = lo, j = hi;
        while (i <= j) {
            while (kv_less(&kv[i], &piv)) ++i;
            while (kv_less(&piv, &kv[j])) --j;
            if (i <= j) { KeyedVertex t = kv[i]; kv[i] = kv[j]; kv[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j; else if (k >= i) lo = i; else return;
    }
}

#define ND_LEAF 48

static void nd_recurse(NdCtx *c, int *verts, int n) {
    if (n <= ND_LEAF) {
        for (int i = 0; i < n; ++i) c->order[c->pos++] = verts[i];
        return;
    }
    /* longest bounding-box axis */
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            double x = c->xyz[3 * (size_t)verts[i] + d];
            if (x < lo[d]) lo[d] = x;
            if (x > hi[d]) hi[d] = x;
        }
    int axis = 0;
    if (hi[1] - lo[1] > hi[axis] - lo[axis]) axis = 1;
    if (hi[2] - lo[2] > hi[axis] - lo[axis]) axis = 2;

    KeyedVertex *kv = c->kv;  /* safe to share: children run after we are done with it */
    for (int i = 0; i < n; ++i) { kv[i].key = c->xyz[3 * (size_t)verts[i] + axis]; kv[i].v = verts[i]; }
    int half = n / 2;
    kv_select(kv, n, half);

    int stampA = ++c->stamp, stampB = ++c->stamp;
    for (int i = 0; i < half; ++i) c->mark[kv[i].v] = stampA;
    for (int i = half; i < n; ++i) c->mark[kv[i].v] = stampB;

    /* separator = vertices of A that touch B; write [A \ sep | sep | B] back into verts */
    int nA = 0, nS = 0;
    int *sep = (int *)malloc(sizeof(int) * (size_t)half);
    for (int i = 0; i < half; ++i) {
        int v = kv[i].v, touches = 0;
        for (int p = c->rp[v]; p < c->rp[v + 1]; ++p)
            if (c->mark[c->ci[p]] == stampB) { touches = 1; break; }
        if (touches) sep[nS++] = v; else verts[nA++] = v;
    }
    int nB = n - half;
    for (int i = 0; i < nB; ++i) verts[nA + i] = kv[half + i].v;

    if (nA == 0 || nS == 0) {
        /* no progress possible by bisection (disconnected halves or everything a separator) */
        if (nS == 0) {           /* halves are disconnected: order independently */
            free(sep);
            nd_recurse(c, verts, nA);
            nd_recurse(c, verts + nA, nB);
            return;
        }
        for (int i = 0; i < nS; ++i) c->order[c->pos++] = sep[i];
        free(sep);
        nd_recurse(c, verts + nA, nB);
        return;
    }
    nd_recurse(c, verts, nA);
    nd_recurse(c, verts + nA, nB);
    for (int i = 0; i < nS; ++i) c->order[c->pos++] = sep[i];
    free(sep);
}

static void nd_order(int n, const int *rp, const int *ci, const double *xyz, int *perm) {
    NdCtx c;
    c.rp = rp; c.ci = ci; c.xyz = xyz;
    c.mark = (int *)calloc((size_t)n, sizeof(int));
    c.stamp = 0;
    c.order = perm;
    c.pos = 0;
    c.kv = (KeyedVertex *)malloc(sizeof(KeyedVertex) * (size_t)(n > 0 ? n : 1));
    int *verts = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) verts[i] = i;
    nd_recurse(&c, verts, n);
    free(verts); free(c.kv); free(c.mark);
}

/* ------------------------------------------------------------------------------------------ */
/* factorisation                                                                              */
/* ------------------------------------------------------------------------------------------ */

SparseLDLT *ldlt_factor(int n, const int *rowptr, const int *colidx, const LDLT_REAL *val,
                        const double *coords) {
    SparseLDLT *F = (SparseLDLT *)calloc(1, sizeof(SparseLDLT));
    if (!F) return NULL;
    size_t nn = (size_t)(n > 0 ? n : 1);
    F->n = n;
    F->perm = (int *)malloc(sizeof(int) * nn);
    F->iperm = (int *)malloc(sizeof(int) * nn);
    F->parent = (int *)malloc(sizeof(int) * nn);
    F->Lp = (int64_t *)calloc(nn + 1, sizeof(int64_t));
    F->D = (LDLT_REAL *)malloc(sizeof(LDLT_REAL) * nn);
    F->work = (LDLT_REAL *)calloc(nn, sizeof(LDLT_REAL));

    if (coords && n > ND_LEAF) nd_order(n, rowptr, colidx, coords, F->perm);
    else for (int i = 0; i < n; ++i) F->perm[i] = i;
    for (int i = 0; i < n; ++i) F->iperm[F->perm[i]] = i;

    /* upper triangle (i <= k) of the permuted matrix, column by column */
    int64_t *Up = (int64_t *)calloc(nn + 1, sizeof(int64_t));
    for (int k = 0; k < n; ++k) {
        int old = F->perm[k];
        int64_t cnt = 0;
        for (int p = rowptr[old]; p < rowptr[old + 1]; ++p)
            if (F->iperm[colidx[p]] <= k) ++cnt;
        Up[k + 1] = Up[k] + cnt;
    }
    int *Ui = (int *)malloc(sizeof(int) * (size_t)(Up[n] > 0 ? Up[n] : 1));
    LDLT_REAL *Ux = (LDLT_REAL *)malloc(sizeof(LDLT_REAL) * (size_t)(Up[n] > 0 ? Up[n] : 1));
    for (int k = 0; k < n; ++k) {
        int old = F->perm[k];
        int64_t q = Up[k];
        for (int p = rowptr[old]; p < rowptr[old + 1]; ++p) {
            int i = F->iperm[colidx[p]];
            if (i <= k) { Ui[q] = i; Ux[q] = val[p]; ++q; }
        }
    }

    /* symbolic: elimination tree and column counts */
    int *flag = (int *)malloc(sizeof(int) * nn);
    int *lnz = (int *)calloc(nn, sizeof(int));
    for (int k = 0; k < n; ++k) {
        F->parent[k] = -1;
        flag[k] = k;
        for (int64_t p = Up[k]; p < Up[k + 1]; ++p) {
            int i = Ui[p];
            while (i < k && flag[i] != k) {
                if (F->parent[i] == -1) F->parent[i] = k;
                ++lnz[i];
                flag[i] = k;
                i = F->parent[i];
            }
        }
    }
    for (int k = 0; k < n; ++k) F->Lp[k + 1] = F->Lp[k] + lnz[k];
    F->lnz = F->Lp[n];
    F->Li = (int *)malloc(sizeof(int) * (size_t)(F->lnz > 0 ? F->lnz : 1));
    F->Lx = (LDLT_REAL *)malloc(sizeof(LDLT_REAL) * (size_t)(F->lnz > 0 ? F->lnz : 1));
    if (!F->Li || !F->Lx) { F->ok = 0; free(Up); free(Ui); free(Ux); free(flag); free(lnz); return F; }

    /* numeric: row k of L solves L(0:k,0:k) y = A(0:k,k) */
    int *pattern = (int *)malloc(sizeof(int) * nn);
    LDLT_REAL *Y = F->work;
    F->ok = 1;
    for (int k = 0; k < n; ++k) {
        int top = n;
        Y[k] = 0;
        flag[k] = k;
        lnz[k] = 0;
        for (int64_t p = Up[k]; p < Up[k + 1]; ++p) {
            int i = Ui[p];
            Y[i] += Ux[p];
            int len = 0;
            while (flag[i] != k) { pattern[len++] = i; flag[i] = k; i = F->parent[i]; }
            while (len > 0) pattern[--top] = pattern[--len];
        }
        LDLT_REAL dk = Y[k];
        Y[k] = 0;
        for (; top < n; ++top) {
            int i = pattern[top];
            LDLT_REAL yi = Y[i];
            Y[i] = 0;
            int64_t pend = F->Lp[i] + lnz[i];
            for (int64_t p = F->Lp[i]; p < pend; ++p) Y[F->Li[p]] -= F->Lx[p] * yi;
            LDLT_REAL lki = yi / F->D[i];
            dk -= lki * yi;
            F->Li[pend] = k;
            F->Lx[pend] = lki;
            ++lnz[i];
        }
        F->D[k] = dk;
        if (dk == 0 || dk != dk) { F->ok = 0; break; }   /* zero pivot -> NumericalIssue */
    }
    memset(Y, 0, sizeof(LDLT_REAL) * nn);
    free(pattern); free(flag); free(lnz); free(Up); free(Ui); free(Ux);
    return F;
}

void ldlt_solve(SparseLDLT *F, LDLT_REAL *x) {
    int n = F->n;
    LDLT_REAL *y = F->work;
    for (int k = 0; k < n; ++k) y[k] = x[F->perm[k]];
    for (int j = 0; j < n; ++j) {
        LDLT_REAL yj = y[j];
        for (int64_t p = F->Lp[j]; p < F->Lp[j + 1]; ++p) y[F->Li[p]] -= F->Lx[p] * yj;
    }
    for (int j = 0; j < n; ++j) y[j] /= F->D[j];
    for (int j = n - 1; j >= 0; --j) {
        LDLT_REAL yj = y[j];
        for (int64_t p = F->Lp[j]; p < F->Lp[j + 1]; ++p) yj -= F->Lx[p] * y[F->Li[p]];
        y[j] = yj;
    }
    for (int k = 0; k < n; ++k) x[F->perm[k]] = y[k];
    memset(y, 0, sizeof(LDLT_REAL) * (size_t)n);
}

void ldlt_free(SparseLDLT *F) {
    if (!F) return;
    free(F->perm); free(F->iperm); free(F->parent); free(F->Lp); free(F->Li); free(F->Lx);
    free(F->D); free(F->work); free(F);
}
