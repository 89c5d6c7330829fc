/*
 * oracle/qc_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C restatement of the reference's support-radius neighbour search
 *   torch_quadconv/quadconv.py:165-199  (QuadConv._compute_eval_indices)
 *   torch_quadconv/quadconv.py:153-154  (QuadConv._bump_arg = linalg.vector_norm over the last dim)
 * and of the reference's pooling reduction
 *   torch_quadconv/mesh_pool.py:88-93   (scatter_reduce_ amax, include_self=False / gather).
 *
 * The reference evaluates, for every (out, in) point pair, in fp32:
 *     z   = out - in                         (quadconv.py:178, component-wise fp32 subtract)
 *     n   = ||z||_2                          (quadconv.py:180)
 *     hit = n <= (float) decay_param         (quadconv.py:182, python scalar demoted to fp32)
 * and lists the hits row-major (out index, then in index) with torch.nonzero (quadconv.py:183).
 * SURVEY.md hard-part 1 records that ATen's CPU vector_norm is bitwise
 *     sqrtf(fmaf(z2,z2, fmaf(z1,z1, z0*z0)))   (d=3),  sqrtf(fmaf(z1,z1, z0*z0)) (d=2),  fabsf(z0) (d=1)
 * tests/test_oracle.py pins this restatement against tests/golden/*.npz, which were produced by the
 * live reference (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/build.py). -ffp-contract=off so
 * the only fused operations are the explicit fmaf() calls.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline float qc_norm(const float *o, const float *p, int d)
{
    if (d == 1) return fabsf(o[0] - p[0]);
    float z0 = o[0] - p[0];
    float acc = z0 * z0;
    for (int k = 1; k < d; ++k) {
        float z = o[k] - p[k];
        acc = fmaf(z, z, acc);
    }
    return sqrtf(acc);
}

/* Pass 1: per-out-point hit counts. quadconv.py:178-183 */
void qc_oracle_count(const float *out_pts, const float *in_pts, int64_t No, int64_t Ni, int d,
                     float r32, int64_t *row_counts)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t o = 0; o < No; ++o) {
        int64_t c = 0;
        const float *po = out_pts + o * d;
        for (int64_t i = 0; i < Ni; ++i)
            c += (qc_norm(po, in_pts + i * d, d) <= r32);
        row_counts[o] = c;
    }
}

/* Pass 2: fill [P,2] int64 (col0 = out idx, col1 = in idx), row-major order as torch.nonzero. */
void qc_oracle_fill(const float *out_pts, const float *in_pts, int64_t No, int64_t Ni, int d,
                    float r32, const int64_t *row_ptr, int64_t *pairs)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t o = 0; o < No; ++o) {
        int64_t w = row_ptr[o];
        const float *po = out_pts + o * d;
        for (int64_t i = 0; i < Ni; ++i) {
            if (qc_norm(po, in_pts + i * d, d) <= r32) {
                pairs[2 * w] = o;
                pairs[2 * w + 1] = i;
                ++w;
            }
        }
    }
}

/* mesh_pool.py:88-89: out = zeros(B,C,Nout); out.scatter_reduce_(2, index, x, 'amax', include_self=False)
 * with index[b,c,k] = group[k] for all b,c.  Groups never empty in the reference (each keeper is in
 * its own group), but an empty group keeps the 0 the reference initialises with. */
void qc_oracle_segmax(const float *x, const int64_t *group, int64_t BC, int64_t Nin, int64_t Nout,
                      float *out)
{
    for (int64_t r = 0; r < BC; ++r) {
        float *y = out + r * Nout;
        const float *xr = x + r * Nin;
        for (int64_t g = 0; g < Nout; ++g) y[g] = 0.0f;
        /* include_self=False: first contribution overwrites, later ones take max */
        for (int64_t pass = 0; pass < 2; ++pass) {
            for (int64_t k = 0; k < Nin; ++k) {
                int64_t g = group[k];
                if (pass == 0) y[g] = -INFINITY;
                else if (xr[k] > y[g]) y[g] = xr[k];
            }
        }
    }
}

/* mesh_pool.py:93: out = gather(x, 2, index) */
void qc_oracle_gather(const float *x, const int64_t *index, int64_t BC, int64_t Nin, int64_t Nout,
                      float *out)
{
    for (int64_t r = 0; r < BC; ++r)
        for (int64_t k = 0; k < Nout; ++k)
            out[r * Nout + k] = x[r * Nin + index[k]];
}
