/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float32 arithmetic) of the FCMA correlation hot path of
 * brainiak/brainiak @ 123f6e1.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; brainiak_b200 never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the reference's own known-answer tests (tests/fcma/test_voxel_selection.py:55-67 golden
 *       normaliser block, test_util.py corrcoef check), and
 *   (2) outputs of the unmodified reference (oracle/_ref, built by oracle/build_ref.sh from
 *       /root/reference) committed as fixtures under tests/golden/ by tests/golden/make_golden.py.
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* a4: correlation block.
 * Reference: voxelselector.py:307-323 -> cython_blas.pyx:115-116
 *   sgemm('N','T', M=V2, N=nb, K=T, A=raw2[e] (ld V2), B=&raw[e][0,start] (ld V), C=&corr[0,e,0], ldc=V2*E)
 * i.e. corr[i, e, j] = sum_t raw[e][t, start+i] * raw2[e][t, j], float32 in, float32 accumulate.
 * The reference's sgemm (SciPy's OpenBLAS, FMA micro-kernels) accumulates every output element as a
 * sequential fp32 FMA chain over t; tests/test_oracle.py checks that this restatement reproduces the
 * reference's outputs (tests/golden) BIT FOR BIT.
 * layout 0: out[nb][E][V2] (VoxelSelector); layout 1: out[E][nb][V2] (Classifier,
 * classifier.py:166-178 -> cython_blas.pyx:477-478 with ldc=V2).
 * raw / raw2: E pointers to C-contiguous [T_e][V] / [T_e][V2]; T[e] may differ per epoch
 * (voxelselector.py:317 uses mat.shape[0]).
 */
void oracle_corr_block(const float *const *raw, const float *const *raw2, const int *T, int E,
                       long V, long V2, long start, long nb, int layout, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (long i = 0; i < nb; i++) {
        for (int e = 0; e < E; e++) {
            const float *a = raw[e];
            const float *b = raw2[e];
            float *dst = layout == 0 ? out + ((size_t)i * E + e) * V2
                                     : out + ((size_t)e * nb + i) * V2;
            for (long j = 0; j < V2; j++) dst[j] = 0.0f;
            for (int t = 0; t < T[e]; t++) {
                float av = a[(size_t)t * V + start + i];
                const float *brow = b + (size_t)t * V2;
                for (long j = 0; j < V2; j++) dst[j] = __builtin_fmaf(av, brow[j], dst[j]);
            }
        }
    }
}

/* Same contraction with float64 accumulation: used for error attribution only. */
void oracle_corr_block_f64(const float *const *raw, const float *const *raw2, const int *T, int E,
                           long V, long V2, long start, long nb, int layout, double *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (long i = 0; i < nb; i++) {
        for (int e = 0; e < E; e++) {
            const float *a = raw[e];
            const float *b = raw2[e];
            double *dst = layout == 0 ? out + ((size_t)i * E + e) * V2
                                      : out + ((size_t)e * nb + i) * V2;
            for (long j = 0; j < V2; j++) dst[j] = 0.0;
            for (int t = 0; t < T[e]; t++) {
                double av = a[(size_t)t * V + start + i];
                const float *brow = b + (size_t)t * V2;
                for (long j = 0; j < V2; j++) dst[j] += av * (double)brow[j];
            }
        }
    }
}

/* a6: Fisher-z + within-subject z-score, in place on data[n0][E][n2].
 * Reference: fcma_extension.cc:52-84 (within_subject_norm_native).
 *   nSubjs = E / eps (integer division; trailing epochs untouched, :52)
 *   num = 1+r, den = 1-r, each clamped to 1e-4 if <= 0 (:68-72; the literal 1e-4 is a double
 *   that is converted to float on assignment)
 *   z = 0.5f*logf(num/den); mean += z; std_dev += z*z   (sequential float32, :73-75)
 *   mean /= eps; var = std_dev/eps - mean*mean; inv = var<=0 ? 0 : 1/sqrt(var)  (:76-78;
 *   `sqrt` on a float argument in C++ <cmath> resolves to the float overload)
 *   z = (z - mean) * inv                                 (:79-82)
 */
void oracle_within_subject_norm(float *data, long n0, int E, long n2, int eps)
{
    if (eps <= 0) return;
    long nSubjs = E / eps;
#pragma omp parallel for schedule(static)
    for (long v = 0; v < n0 * nSubjs; v++) {
        long s = v % nSubjs;
        long i = v / nSubjs;
        float *mat = data + (size_t)i * E * n2;
        for (long j = 0; j < n2; j++) {
            float mean = 0.0f, std_dev = 0.0f;
            for (long b = s * eps; b < (s + 1) * eps; b++) {
                float r = mat[(size_t)b * n2 + j];
                float num = 1.0f + r;
                float den = 1.0f - r;
                num = (num <= 0.0f) ? (float)1e-4 : num;
                den = (den <= 0.0f) ? (float)1e-4 : den;
                float z = 0.5f * logf(num / den);
                mat[(size_t)b * n2 + j] = z;
                mean += z;
                std_dev += z * z;
            }
            mean = mean / (float)eps;
            std_dev = std_dev / (float)eps - mean * mean;
            float inv = (std_dev <= 0.0f) ? 0.0f : 1.0f / sqrtf(std_dev);
            for (long b = s * eps; b < (s + 1) * eps; b++)
                mat[(size_t)b * n2 + j] = (mat[(size_t)b * n2 + j] - mean) * inv;
        }
    }
}

/* a7 / a11: linear kernel K = beta*K + Z Z^T for one [E][K] slab (full square written).
 * Reference: cython_blas.pyx:197-207 ssyrk('L','T',N=E,K=k,alpha=1,A,lda=k,beta,C,ldc=E) + mirror.
 * float32 accumulate, k ascending.
 */
void oracle_kernel_matrix(const float *z, int E, long k, float beta, float *K)
{
    for (int a = 0; a < E; a++) {
        for (int b = 0; b <= a; b++) {
            const float *za = z + (size_t)a * k;
            const float *zb = z + (size_t)b * k;
            float acc = 0.0f;
            for (long j = 0; j < k; j++) acc += za[j] * zb[j];
            float v = (beta == 0.0f ? 0.0f : beta * K[a * E + b]) + acc;
            K[a * E + b] = v;
            K[b * E + a] = v;
        }
    }
}

/* batched: K[i] = Z_i Z_i^T for i < nb (voxelselector.py:400-408 loop), Z = [nb][E][V2]. */
void oracle_kernel_matrices(const float *z, long nb, int E, long V2, float *K)
{
#pragma omp parallel for schedule(dynamic)
    for (long i = 0; i < nb; i++)
        oracle_kernel_matrix(z + (size_t)i * E * V2, E, V2, 0.0f, K + (size_t)i * E * E);
}

/* float64-accumulating variant for error attribution. */
void oracle_kernel_matrices_f64(const float *z, long nb, int E, long V2, double *K)
{
#pragma omp parallel for schedule(dynamic)
    for (long i = 0; i < nb; i++) {
        const float *zi = z + (size_t)i * E * V2;
        double *Ki = K + (size_t)i * E * E;
        for (int a = 0; a < E; a++)
            for (int b = 0; b <= a; b++) {
                double acc = 0.0;
                for (long j = 0; j < V2; j++)
                    acc += (double)zi[(size_t)a * V2 + j] * (double)zi[(size_t)b * V2 + j];
                Ki[a * E + b] = acc;
                Ki[b * E + a] = acc;
            }
    }
}

/* a14: per-epoch normalisation of one [T][V] block, in place.
 * Reference: preprocessing.py:80-84: zscore(axis=0, ddof=0) -> nan_to_num -> / sqrt(T).
 * scipy's zscore on a float32 array computes mean/std with numpy float32 reductions (pairwise
 * summation), so bits can differ in the last ulp; the restatement accumulates in float64 and
 * rounds once, which is within 1 ulp of any float32 summation order.
 */
void oracle_epoch_normalize(float *mat, int T, long V)
{
    double rs = sqrt((double)T);
#pragma omp parallel for schedule(static)
    for (long v = 0; v < V; v++) {
        double m = 0.0;
        for (int t = 0; t < T; t++) m += mat[(size_t)t * V + v];
        m /= T;
        double q = 0.0;
        for (int t = 0; t < T; t++) {
            double d = mat[(size_t)t * V + v] - m;
            q += d * d;
        }
        double sd = sqrt(q / T);
        for (int t = 0; t < T; t++) {
            double z = (mat[(size_t)t * V + v] - m) / sd; /* 0/0 -> nan -> 0 (nan_to_num) */
            if (!(z == z)) z = 0.0;
            else if (isinf(z)) z = z > 0 ? 3.4028234663852886e38 : -3.4028234663852886e38;
            mat[(size_t)t * V + v] = (float)(z / rs);
        }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
