/* TEST INFRASTRUCTURE — type-generic body of the CPU oracle, included four times by hb_oracle.c.
 * Parameters: T (scalar), R (its precision type), NAME(x) (suffix macro), CPLX (0/1), CONJ(z), ABS(z), REAL(z).
 * Every function restates, in plain serial C, what the reference's cpu_engine path computes; the
 * reference file:line each one follows is cited on the function. No part of the product links this.
 */

/* sparse/hala_sparse_utils.hpp:103-118  sparse_gemv_array<T,trans>
 * 'N': per row, left-to-right sum, y[i] = alpha*sum + beta*y[i]  (y is read even when beta == 0, as the reference does)
 * 'T'/'C': y *= beta (scal over num_cols), then y[indx[j]] += alpha * x[i] * op(vals[j]) in row order. */
static void NAME(spmv)(char trans, int M, int N, T alpha, const int *pntr, const int *indx, const T *vals,
                       const T *x, T beta, T *y){
    if (trans == 'N' || trans == 'n'){
        for(int i=0; i<M; i++){
            T sum = 0;
            for(int j=pntr[i]; j<pntr[i+1]; j++)
                sum += vals[j] * x[indx[j]];
            y[i] = alpha * sum + beta * y[i];
        }
    }else{
        int cj = (trans == 'C' || trans == 'c');
        for(int i=0; i<N; i++) y[i] = beta * y[i];
        for(int i=0; i<M; i++)
            for(int j=pntr[i]; j<pntr[i+1]; j++)
                y[indx[j]] += alpha * x[i] * (cj ? CONJ(vals[j]) : vals[j]);
    }
}

/* blas/hala_blas_1.hpp:53-61 vcopy -> ?copy_ */
static void NAME(copy)(int n, const T *x, int incx, T *y, int incy){
    for(int i=0; i<n; i++) y[(size_t) i * incy] = x[(size_t) i * incx];
}
/* blas/hala_blas_1.hpp:209-217 axpy -> ?axpy_ : y += alpha x */
static void NAME(axpy)(int n, T alpha, const T *x, int incx, T *y, int incy){
    for(int i=0; i<n; i++) y[(size_t) i * incy] += alpha * x[(size_t) i * incx];
}
/* blas/hala_blas_1.hpp:316-335 scal -> ?scal_ : x *= alpha */
static void NAME(scal)(int n, T alpha, T *x, int incx){
    for(int i=0; i<n; i++) x[(size_t) i * incx] = alpha * x[(size_t) i * incx];
}
/* blas/hala_blas_1.hpp:229-246 dot_array<conjugate> (complex: HALA's own loop; real: ?dot_), :277-290 */
static T NAME(dot)(int cj, int n, const T *x, int incx, const T *y, int incy){
    T sum = 0;
    for(int i=0; i<n; i++)
        sum += (cj ? CONJ(x[(size_t) i * incx]) : x[(size_t) i * incx]) * y[(size_t) i * incy];
    return sum;
}
/* blas/hala_blas_1.hpp:104-121 norm2 -> ?nrm2_ ; restated as the netlib scaled sum-of-squares loop (dnrm2/dznrm2, BLAS 3.8) */
static R NAME(nrm2)(int n, const T *x, int incx){
    R scale = 0, ssq = 1;
    for(int i=0; i<n; i++){
#if CPLX
        R parts[2] = { REAL(x[(size_t) i * incx]), IMAG(x[(size_t) i * incx]) };
        for(int k=0; k<2; k++){
            R a = parts[k] < 0 ? -parts[k] : parts[k];
#else
        {
            R a = x[(size_t) i * incx] < 0 ? -x[(size_t) i * incx] : x[(size_t) i * incx];
#endif
            if (a != 0){
                if (scale < a){ ssq = 1 + ssq * (scale / a) * (scale / a); scale = a; }
                else          { ssq += (a / scale) * (a / scale); }
            }
        }
    }
    return scale * SQRT(ssq);
}

/* blas/hala_blas_2.hpp gemv -> ?gemv_ ; column-major A (lda), unit strides. Used by krylov_project / krylov_combine
 * (hex/solvers/hala_solvers_gmres.hpp:47-62). 'T': y[j] = alpha * sum_i A[i,j] x[i] + beta y[j]; 'C' conjugates A. */
static void NAME(gemv)(char trans, int M, int N, T alpha, const T *A, int lda, const T *x, T beta, T *y){
    if (trans == 'N' || trans == 'n'){
        for(int i=0; i<M; i++) y[i] = (beta == 0) ? 0 : beta * y[i];
        for(int j=0; j<N; j++){
            T t = alpha * x[j];
            const T *a = A + (size_t) j * lda;
            for(int i=0; i<M; i++) y[i] += t * a[i];
        }
    }else{
        int cj = (trans == 'C' || trans == 'c');
        for(int j=0; j<N; j++){
            const T *a = A + (size_t) j * lda;
            T sum = 0;
            for(int i=0; i<M; i++) sum += (cj ? CONJ(a[i]) : a[i]) * x[i];
            y[j] = (beta == 0) ? alpha * sum : alpha * sum + beta * y[j];
        }
    }
}

/* blas/hala_blas_1.hpp:361-365 rotg -> ?rotg_ (netlib BLAS 3.8 drotg / zrotg) */
static void NAME(rotg)(T *a, T *b, R *c, T *s){
#if CPLX
    R absa = ABS(*a);
    if (absa == 0){
        *c = 0; *s = 1; *a = *b;
    }else{
        R scale = absa + ABS(*b);
        R na = ABS(*a / scale), nb = ABS(*b / scale);
        R norm = scale * SQRT(na * na + nb * nb);
        T alpha = *a / absa;
        *c = absa / norm;
        *s = alpha * CONJ(*b) / norm;
        *a = alpha * norm;
    }
#else
    R roe = *b, absa = ABS(*a), absb = ABS(*b);
    if (absa > absb) roe = *a;
    R scale = absa + absb;
    if (scale == 0){
        *c = 1; *s = 0; *a = 0; *b = 0;
    }else{
        R r = scale * SQRT((*a / scale) * (*a / scale) + (*b / scale) * (*b / scale));
        if (roe < 0) r = -r;
        *c = *a / r; *s = *b / r;
        R z = 1;
        if (absa > absb) z = *s;
        if (absb >= absa && *c != 0) z = 1 / *c;
        *a = r; *b = z;
    }
#endif
}
/* blas/hala_blas_1.hpp:367-407 rot -> ?rot_ (real) / crot_,zrot_ (complex s): x' = c x + s y ; y' = c y - conj(s) x */
static void NAME(rot1)(T *x, T *y, R c, T s){
    T tx = c * (*x) + s * (*y);
    *y = c * (*y) - CONJ(s) * (*x);
    *x = tx;
}
/* blas/hala_blas_2.hpp:367-374 tpsv('U','N','N') -> ?tpsv_ : packed upper, column by column, back substitution */
static void NAME(tpsv_unn)(int n, const T *ap, T *x){
    size_t kk = (size_t) n * (n + 1) / 2;          /* one past the last packed entry */
    for(int j=n-1; j>=0; j--){
        size_t diag = kk - 1;                       /* AP(j,j) */
        if (x[j] != 0){
            x[j] = x[j] / ap[diag];
            T t = x[j];
            size_t k = diag - 1;
            for(int i=j-1; i>=0; i--, k--) x[i] -= t * ap[k];
        }
        kk -= (size_t) j + 1;
    }
}

/* hex/solvers/hala_solvers_cg.hpp:92-156 solve_cg_core driven as :181-227 wires it, identity preconditioner
 * (z = copy of r). Order of operations, the iteration counter (starts at 1, +1 per operator application)
 * and the stop test  (i == max_iter) || (||r||_2 < tol)  on the recursively updated residual are the reference's. */
static int NAME(cg)(int n, const int *pntr, const int *indx, const T *vals, const T *b, T *x, R tol, int max_iter){
    T *r = (T*) malloc(sizeof(T) * (size_t) n), *p = (T*) malloc(sizeof(T) * (size_t) n);
    T *Ap = (T*) malloc(sizeof(T) * (size_t) n), *z = (T*) malloc(sizeof(T) * (size_t) n);
    for(int i=0; i<n; i++){ p[i] = 0; Ap[i] = 0; }
    NAME(copy)(n, b, 1, r, 1);                              /* r = b */
    NAME(spmv)('N', n, n, 1, pntr, indx, vals, x, 0, p);    /* p = A x */
    NAME(axpy)(n, -1, p, 1, r, 1);                          /* r -= p */
    NAME(copy)(n, r, 1, z, 1);                              /* z = P^-1 r */
    NAME(copy)(n, z, 1, p, 1);                              /* p = z */
    T zr = NAME(dot)(1, n, r, 1, z, 1);
    int iterations = 1, iterate = 1;
    while(iterate){
        NAME(spmv)('N', n, n, 1, pntr, indx, vals, p, 0, Ap);
        iterations++;
        T nzr = NAME(dot)(1, n, p, 1, Ap, 1);
        T a = zr / nzr;
        NAME(axpy)(n,  a, p, 1, x, 1);
        NAME(axpy)(n, -a, Ap, 1, r, 1);
        iterate = !((iterations == max_iter) || (NAME(nrm2)(n, r, 1) < tol));
        if (iterate){
            NAME(copy)(n, r, 1, z, 1);
            nzr = NAME(dot)(1, n, r, 1, z, 1);
            a = nzr / zr;
            NAME(scal)(n, a, p, 1);
            NAME(axpy)(n, 1, z, 1, p, 1);
            zr = nzr;
        }
    }
    free(r); free(p); free(Ap); free(z);
    return iterations;
}

/* hex/solvers/hala_solvers_gmres.hpp:127-230 solve_gmres, identity preconditioner, with krylov_project :47-50
 * (single-pass classical Gram-Schmidt as gemv(op)+gemv('N')) and krylov_combine :60-62.
 * cproj = 0: op = 'T' exactly as the reference; cproj = 1: op = 'C' (the patched oracle for complex data).
 * Reference quirks kept on purpose: Z is rotated only when the inner loop continues (:210-216), the inner
 * residual estimate is |S_j * Z_j| with the un-rotated Z_j (:207), max_iter bounds OUTER iterations (:170). */
static int NAME(gmres)(int n, const int *pntr, const int *indx, const T *vals, const T *b, T *x, R tol,
                       int max_outer, int restart, int cproj){
    T *t = (T*) malloc(sizeof(T) * (size_t) n), *r = (T*) malloc(sizeof(T) * (size_t) n);
    T *W = (T*) malloc(sizeof(T) * (size_t) n * (size_t) restart);
    T *H = (T*) malloc(sizeof(T) * (size_t) restart * (restart + 1));
    T *S = (T*) malloc(sizeof(T) * (size_t) (restart + 1)), *Z = (T*) malloc(sizeof(T) * (size_t) (restart + 1));
    R *C = (R*) malloc(sizeof(R) * (size_t) (restart + 1));
    T *coeffs = (T*) malloc(sizeof(T) * (size_t) (restart + 1));
    R inner_res, outer_res = tol + 1;
    int total = 0, outer = 0;
    while((outer_res > tol) && (outer < max_outer)){
        size_t hsize = 0; int nz = 0, ns = 0;
        NAME(copy)(n, b, 1, t, 1);
        NAME(spmv)('N', n, n, -1, pntr, indx, vals, x, 1, t);       /* t = b - A x */
        NAME(copy)(n, t, 1, r, 1);                                  /* r = P^-1 t */
        total++;
        inner_res = NAME(nrm2)(n, r, 1);
        NAME(scal)(n, (T) ((R) 1 / inner_res), r, 1);
        Z[nz++] = inner_res;
        NAME(copy)(n, r, 1, W, 1);
        int inner = 0;
        while((inner_res > tol) && (inner < restart)){
            for(int i=0; i<n; i++) t[i] = 0;
            NAME(spmv)('N', n, n, 1, pntr, indx, vals, r, 0, t);
            NAME(copy)(n, t, 1, r, 1);
            total++;
            NAME(gemv)(cproj ? 'C' : 'T', n, inner + 1, 1, W, n, r, 0, coeffs);
            NAME(gemv)('N', n, inner + 1, -1, W, n, coeffs, 1, r);
            R nrm = NAME(nrm2)(n, r, 1);
            NAME(scal)(n, (T) ((R) 1 / nrm), r, 1);
            for(int i=0; i<inner; i++) NAME(rot1)(&coeffs[i], &coeffs[i+1], C[i], S[i]);
            T isin, beta = nrm; R icos;
            NAME(rotg)(&coeffs[inner], &beta, &icos, &isin);
            for(int i=0; i<=inner; i++) H[hsize++] = coeffs[i];
            S[ns] = isin; C[ns] = icos; ns++;
            inner_res = ABS(S[ns-1] * Z[nz-1]);
            inner++;
            if ((inner_res > tol) && (inner < restart)){
                NAME(copy)(n, r, 1, W + (size_t) inner * n, 1);
                Z[nz++] = 0;
                NAME(rot1)(&Z[inner-1], &Z[inner], C[ns-1], S[ns-1]);
            }
        }
        if (hsize > 0){
            NAME(tpsv_unn)(nz, H, Z);
            NAME(gemv)('N', n, nz, 1, W, n, Z, 1, x);
        }
        outer++;
        outer_res = inner_res;
    }
    free(t); free(r); free(W); free(H); free(S); free(Z); free(C); free(coeffs);
    return total;
}

/* sparse/hala_sparse_utils.hpp:283-335  sparse_trsv_array<diag,T>: x = alpha * op(T)^-1 b for a CSR that holds exactly
 * one triangle with sorted rows — the diagonal entry is the LAST of its row for uplo 'L' and the FIRST for uplo 'U'
 * (stored, and skipped, even when diag == 'U').  'N' walks the rows (forward for L, backward for U) with a left-to-right
 * sum; 'T'/'C' zero x and scatter column-wise in the opposite row order. */
static void NAME(trsv)(char uplo, char diag, char trans, int n, T alpha, const int *pntr, const int *indx, const T *vals,
                       const T *b, T *x){
    const int unit = (diag == 'U' || diag == 'u'), lower = (uplo == 'L' || uplo == 'l');
    if (n <= 0) return;
    if (trans == 'N' || trans == 'n'){
        if (lower){
            x[0] = unit ? alpha * b[0] : alpha * b[0] / vals[0];
            for(int i=1; i<n; i++){
                T s = 0;
                for(int j=pntr[i]; j<pntr[i+1]-1; j++) s += vals[j] * x[indx[j]];
                x[i] = unit ? (alpha * b[i] - s) : (alpha * b[i] - s) / vals[pntr[i+1]-1];
            }
        }else{
            x[n-1] = unit ? alpha * b[n-1] : alpha * b[n-1] / vals[pntr[n]-1];
            for(int i=n-2; i>=0; i--){
                T s = 0;
                for(int j=pntr[i]+1; j<pntr[i+1]; j++) s += vals[j] * x[indx[j]];
                x[i] = unit ? (alpha * b[i] - s) : (alpha * b[i] - s) / vals[pntr[i]];
            }
        }
    }else{
        const int cj = (trans == 'C' || trans == 'c');
        for(int i=0; i<n; i++) x[i] = 0;
        if (lower){
            for(int i=n-1; i>=0; i--){
                T d = vals[pntr[i+1]-1];
                x[i] = unit ? (alpha * b[i] - x[i]) : (alpha * b[i] - x[i]) / (cj ? CONJ(d) : d);
                for(int j=pntr[i]; j<pntr[i+1]-1; j++) x[indx[j]] += x[i] * (cj ? CONJ(vals[j]) : vals[j]);
            }
        }else{
            for(int i=0; i<n; i++){
                T d = vals[pntr[i]];
                x[i] = unit ? (alpha * b[i] - x[i]) : (alpha * b[i] - x[i]) / (cj ? CONJ(d) : d);
                for(int j=pntr[i]+1; j<pntr[i+1]; j++) x[indx[j]] += x[i] * (cj ? CONJ(vals[j]) : vals[j]);
            }
        }
    }
}

/* The same solve on a GENERAL CSR of which only the `uplo` triangle is used (entries on the other side are skipped, the
 * diagonal is looked up by column index; diag 'U' ignores a stored diagonal) — the cuSPARSE fill-mode semantics that
 * gpu_ilu relies on when it hands the full ILU array to both triangular matrices (gpu/hala_gpu_ilu.hpp:88-89).  For a
 * one-triangle CSR this performs exactly the operations of NAME(trsv) in the same order. */
static void NAME(trsv_general)(char uplo, char diag, char trans, int n, T alpha, const int *pntr, const int *indx, const T *vals,
                               const T *b, T *x){
    const int unit = (diag == 'U' || diag == 'u'), lower = (uplo == 'L' || uplo == 'l');
    const int nt = (trans == 'N' || trans == 'n'), cj = (trans == 'C' || trans == 'c');
    if (!nt) for(int i=0; i<n; i++) x[i] = 0;
    /* 'N' on L and 'T' on U run forward; 'N' on U and 'T' on L run backward */
    const int forward = (nt == lower);
    for(int step=0; step<n; step++){
        const int i = forward ? step : n - 1 - step;
        T d = 1, s = 0;
        for(int j=pntr[i]; j<pntr[i+1]; j++){
            const int c = indx[j];
            if (c == i){ if (!unit) d = vals[j]; }
            else if (nt && (lower ? c < i : c > i)) s += vals[j] * x[c];
        }
        if (nt) x[i] = unit ? (alpha * b[i] - s) : (alpha * b[i] - s) / d;
        else{
            x[i] = unit ? (alpha * b[i] - x[i]) : (alpha * b[i] - x[i]) / (cj ? CONJ(d) : d);
            for(int j=pntr[i]; j<pntr[i+1]; j++){
                const int c = indx[j];
                if (lower ? c < i : c > i) x[c] += x[i] * (cj ? CONJ(vals[j]) : vals[j]);
            }
        }
    }
}

/* sparse/hala_sparse_utils.hpp:228-253 factorize_ilu_array (after get_diagonal_index and a copy of the values): right-looking
 * ILU(0) on sorted rows with the diagonal present.  Returns 0, or 1 + row when a row has no diagonal entry. */
static int NAME(ilu_factor)(int n, const int *pntr, const int *indx, const T *vals, int *diag, T *ilu){
    for(int i=0; i<n; i++){
        int j = pntr[i];
        while (j < pntr[i+1] && indx[j] < i) j++;
        if (j == pntr[i+1] || indx[j] != i) return 1 + i;
        diag[i] = j;
    }
    for(int j=0; j<pntr[n]; j++) ilu[j] = vals[j];
    for(int i=0; i<n; i++){
        T u = ilu[diag[i]];
        for(int j=i+1; j<n; j++){
            int jc = pntr[j];
            while (indx[jc] < i) jc++;          /* the diagonal (column j > i) stops the scan */
            if (indx[jc] == i){
                ilu[jc] /= u;
                T l = ilu[jc];
                int ik = diag[i] + 1, jk = jc + 1;
                while (ik < pntr[i+1] && jk < pntr[j+1]){
                    if (indx[ik] == indx[jk]){ ilu[jk] -= l * ilu[ik]; ik++; jk++; }
                    else if (indx[ik] < indx[jk]) ik++;
                    else jk++;
                }
            }
        }
    }
    return 0;
}
/* sparse/hala_sparse_utils.hpp:261-274 apply_ilu_array: x <- U^-1 L^-1 x (unit lower, stored diagonal upper), in place */
static void NAME(ilu_apply)(int n, const int *pntr, const int *indx, const int *diag, const T *ilu, T *x){
    for(int i=1; i<n; i++)
        for(int j=pntr[i]; j<diag[i]; j++) x[i] -= ilu[j] * x[indx[j]];
    for(int i=n-1; i>=0; i--){
        for(int j=diag[i]+1; j<pntr[i+1]; j++) x[i] -= ilu[j] * x[indx[j]];
        x[i] /= ilu[diag[i]];
    }
}

/* sparse/hala_sparse_utils.hpp:120-160 sparse_gemm_array: C = alpha op(A) op(B) + beta C, one sparse_gemv_array per column of C;
 * op(B) = B^T / B^H first copies the row of B into a contiguous (conjugated) vector. C is M x N (ldc), op(A) is M x K. */
static void NAME(spmm)(char transa, char transb, int M, int N, int K, T alpha, const int *pntr, const int *indx, const T *vals,
                       const T *B, int ldb, T beta, T *C, int ldc){
    const int an = (transa == 'N' || transa == 'n'), bn = (transb == 'N' || transb == 'n'), bc = (transb == 'C' || transb == 'c');
    T *x = (T*) malloc(sizeof(T) * (size_t) (K > 0 ? K : 1));
    for(int i=0; i<N; i++){
        const T *col = B + (size_t) i * ldb;
        if (!bn){
            for(int k=0; k<K; k++){ T v = B[i + (size_t) k * ldb]; x[k] = bc ? CONJ(v) : v; }
            col = x;
        }
        if (an) NAME(spmv)('N', M, K, alpha, pntr, indx, vals, col, beta, C + (size_t) i * ldc);
        else    NAME(spmv)(transa, K, M, alpha, pntr, indx, vals, col, beta, C + (size_t) i * ldc);
    }
    free(x);
}
