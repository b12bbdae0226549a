/*
 * kelvin_b200.h -- C ABI of the B200-native FT-CCSD hot path.
 *
 * The reference (awhite862/kelvin) has no FFI of its own: its hot path is
 * NumPy einsum -> BLAS, reached through the Python functions listed below.
 * Each entry point here names the reference interface whose arithmetic it
 * replaces (paths relative to the reference repo root).  All pointers are
 * DEVICE pointers owned by the caller (PyTorch's allocator) unless marked
 * "host"; `stream` is a cudaStream_t passed as void*.  Every function returns
 * 0 on success or a negative code (-1 bad argument, -2 CUDA error; text via
 * kb200_last_error()).  No hidden allocation, no global state, FP64 throughout.
 */
#ifndef KELVIN_B200_H
#define KELVIN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int kb200_version(void);
const char* kb200_last_error(void);
/* Number of kernels this library has launched since load / since last reset
 * (bench.py's "gpu_launches"). */
int64_t kb200_launch_count(void);
void kb200_launch_count_reset(void);
/* A captured plan (CUDA graph) replays its kernels without passing through this library: the
 * caller adds the number of kernels the graph holds (counted while it was captured) per replay. */
void kb200_launch_count_add(int64_t n);

/* ------------------------------------------------------------------------
 * Contraction plans.
 *
 * Replaces: the bodies of cqcpy.cc_equations._Stanton / _u_Stanton /
 * _Lambda_opt / _uccsd_Lambda_opt / _LS_TS / ccsd_{1,2}rdm_*_opt, called from
 * kelvin/ft_cc_equations.py:106,153-155,399,442-445,407,456,718-720,744-751,
 * i.e. every pyscf.lib.einsum -> dgemm of the per-tau-point residual.
 *
 * One op is a batched tensor contraction with composite ("gathered") indices
 *   C[b][cm[m]+cn[n]] = beta*C + alpha * sum_k A[b][am[m]+ak[k]] * B[b][bk[k]+bn[n]]
 * where am/ak/bk/bn/cm/cn are uint32 element-offset tables (one entry per
 * composite index value) living in one device buffer `tables`; the index
 * permutations of the reference's einsum strings are folded into these tables,
 * never materialised.  b runs over imaginary-time grid points (stride 0 for
 * the tau-independent integrals).  kind==1 is the index-permuted axpby
 *   C[b][cm[m]+cn[n]] = beta*C + alpha * A[b][am[m]+ak[n]].
 * ------------------------------------------------------------------------ */
typedef struct kb200_op {
    int32_t kind;          /* 0 contraction (DMMA GEMM), 1 permuted axpby,
                              2 contraction with K,N <= 64 (streaming rank-K update),
                              3 one term of a fused elementwise sum
                                C = beta*C + sum_t alpha_t X_t[tAm[m]+tAk[n]] * Y_t[tBk[m]+tBn[n]]
                                (a = X, b = Y or -1; `group` consecutive entries, same C/M/N;
                                 beta of the first entry applies, a_mode = read X m-fast) */
    int32_t a, b, c;       /* tensor slot numbers (b unused for kind 1)        */
    int64_t a_off, b_off, c_off;   /* constant element offsets inside the slot */
    int32_t M, N, K, batch;
    int64_t bsA, bsB, bsC; /* batch strides in elements                        */
    int64_t tAm, tAk, tBk, tBn, tCm, tCn;  /* table starts (uint32 units)      */
    double alpha, beta;
    int32_t a_mode;        /* 0: A gathers contiguously along k, 1: along m    */
    int32_t b_mode;        /* 0: B gathers contiguously along k, 1: along n    */
    int32_t tile;          /* CTA tile id: 0 128x128, 1 128x32, 5 64x64 (see kb200.cu) */
    int32_t splitk;        /* >=1; >1 uses the workspace + deterministic reduce */
    int32_t group;         /* kind 0: this op and the next group-1 ops (same tile/a_mode/b_mode,
                              splitk 1, mutually independent) share one launch; 0/1 = alone */
    int32_t reserved;      /* tile 7: bit 0 = rows (not columns) of C are the contiguous direction */
    int64_t ldA, ldB;      /* > 0: the operand is a plain matrix -- its row and k tables are affine
                              and the contiguous index (k for mode 0, row/column for mode 1) has
                              stride 1; ld = pitch of the other index in elements.  With an even
                              pitch and a 16-byte aligned base the kernel feeds the operand by TMA
                              (cp.async.bulk.tensor through a tensor map built at launch) instead
                              of the offset tables.  0: gathered through the tables. */
} kb200_op;

/* Bytes of workspace kb200_plan_run needs for these ops (split-K partials). */
int64_t kb200_plan_workspace_bytes(const kb200_op* ops /*host*/, int nops);

/* Run ops[0..nops) in order on `stream`.  slots[i] is the device base pointer
 * of tensor slot i (host array of device pointers). */
int kb200_plan_run(const kb200_op* ops /*host*/, int nops,
                   const uint32_t* tables, double* const* slots /*host*/, int nslots,
                   double* workspace, int64_t workspace_bytes, void* stream);

/* Streams a plan run uses: 1 = everything in plan order on the caller's stream (KB200_STREAMS=1
 * in the environment selects it); n >= 3 = two streams for the wide (m^6) launches plus n - 2
 * (at most 6) for the small ones, default 2 + 6.  The side streams fork from and join the
 * caller's stream, with every read/write order on every slot (and on each split-K workspace
 * region) kept by events: same results, bit for bit.  Returns the old value. */
int kb200_set_plan_streams(int n);

/* Same as kb200_plan_run, but brackets every op with CUDA events on `stream`,
 * synchronises, and returns the device time of each op in op_ms[nops] (host).
 * Used by bench.py for the live roofline measurement; not on the product path. */
int kb200_plan_run_timed(const kb200_op* ops /*host*/, int nops,
                         const uint32_t* tables, double* const* slots /*host*/, int nslots,
                         double* workspace, int64_t workspace_bytes, void* stream,
                         float* op_ms /*host*/);

/* ------------------------------------------------------------------------
 * Imaginary-time integration.
 * Replaces kelvin/quadrature.py:292-317 (int_tbar1, int_tbar2):
 *   out[y,p] = sum_x G[y,x] * w(y,x,p) * tbar[x,p],
 *   w = exp(D[p]*(ti[x]-ti[y])) for x<y, 1 otherwise   (quirk Q4).
 * n = elements per grid point.  ti[ng], G[ng*ng] row-major, D[n].
 * mode bit 0 = 0: one exp per (y,x) pair, exactly the reference's formula;
 *            = 1: ng-1 exps per element, weights built as running products.
 * mode bit 1 (+2): caller guarantees G[y,x] == 0 for x > y (true for every rule of
 *            kelvin/quadrature.py); the kernel then skips the upper triangle.
 * ------------------------------------------------------------------------ */
int kb200_int_tbar(int ng, int64_t n, const double* tbar, const double* D,
                   const double* ti, const double* G, double* out, int mode, void* stream);

/* Rows y0 <= y < y1 only (out has y1-y0 rows): the tau-sharded form -- each GPU
 * integrates its own grid points from the all-gathered tbar. */
int kb200_int_tbar_rows(int ng, int64_t n, const double* tbar, const double* D,
                        const double* ti, const double* G, double* out, int y0, int y1,
                        int mode, void* stream);

/* As kb200_int_tbar_rows with explicit row strides (elements): tbar row x starts at
 * tbar + x*tstride, out row y - y0 at out + (y - y0)*ostride -- the blocks of several tensors may
 * share one (ng, Ntot) exchange buffer (tau-sharded runs all-gather ONE buffer). */
int kb200_int_tbar_strided(int ng, int64_t n, const double* tbar, int64_t tstride,
                           const double* D, const double* ti, const double* G, double* out,
                           int64_t ostride, int y0, int y1, int mode, void* stream);

/* The fused amplitude update of one block.  Replaces, in ONE pass over tbar and the amplitudes,
 * kelvin/quadrature.py:292-317 (integration), kelvin/cc_utils.py:278-295 (residual norms,
 * damping) and the block's term of kelvin/ft_cc_energy.py:35-72:
 *   new[y]  = sum_x G[y,x] w(y,x) tbar[x]                      (never stored)
 *   out4[0] = ||new - amp||^2, out4[1] = ||amp||^2 (before), amp <- alpha*amp + (1-alpha)*new,
 *   out4[2] = ||amp||^2 (after),
 *   out4[3] = sum_y g[y] sum_p (c2*amp[y,p] + c11*T1x[y,a,i]*T1y[y,b,j]) * W[p]   (W == NULL: 0)
 * over rows y0 <= y < y1; p = (a,b,i,j) over (n/(nvb*noa*nob), nvb, noa, nob) when T1x/T1y are
 * given (row strides t1xs/t1ys; they must already hold the UPDATED singles).  `scratch` needs
 * kb200_reduce_scratch_doubles() doubles. */
int kb200_int_tbar_update(int ng, int64_t n, const double* tbar, int64_t tstride,
                          const double* D, const double* ti, const double* G, double* amp,
                          int64_t astride, int y0, int y1, double alpha, const double* W,
                          const double* T1x, const double* T1y, int64_t t1xs, int64_t t1ys,
                          int nvb, int noa, int nob, const double* g, double c2, double c11,
                          double* out4, double* scratch, int mode, void* stream);

/* The *_h forms additionally take HOST copies of ti[ng], g[ng] and G[ng*ng] (NULL = not given):
 * for the grid sizes of the reference's benchmarks (ng = 10, 16), all rows and a lower-triangular
 * G they run a fully unrolled kernel that holds the quadrature in its parameter space -- same
 * arithmetic, same summation order. */
int kb200_int_tbar_strided_h(int ng, int64_t n, const double* tbar, int64_t tstride,
                             const double* D, const double* ti, const double* G, double* out,
                             int64_t ostride, int y0, int y1, int mode, const double* ti_h /*host*/,
                             const double* G_h /*host*/, void* stream);
int kb200_int_tbar_update_h(int ng, int64_t n, const double* tbar, int64_t tstride,
                            const double* D, const double* ti, const double* G, double* amp,
                            int64_t astride, int y0, int y1, double alpha, const double* W,
                            const double* T1x, const double* T1y, int64_t t1xs, int64_t t1ys,
                            int nvb, int noa, int nob, const double* g, double c2, double c11,
                            double* out4, double* scratch, int mode, const double* ti_h /*host*/,
                            const double* g_h /*host*/, const double* G_h /*host*/, void* stream);

/* Replaces kelvin/quadrature.py:320-345 (int_L1, int_L2):
 *   out[s,q] = (1/g[s]) sum_y g[y]*G[y,s]*w(s,y,q)*L[y,q],
 *   w = exp(D[perm(q)]*(ti[s]-ti[y])) for y>=s, 1 otherwise.
 * L and out are (ng, d0,d1,d2,d3) C-contiguous; D is addressed as
 * D[i0*ds0+i1*ds1+i2*ds2+i3*ds3] so that the reference's 'yabij,yijab->yijab'
 * index swap needs no transposed copy. */
int kb200_int_L(int ng, const int32_t dims[4] /*host*/, const int64_t dstride[4] /*host*/,
                const double* L, const double* D, const double* ti, const double* g,
                const double* G, double* out, int mode, void* stream);

int kb200_int_L_rows(int ng, const int32_t dims[4] /*host*/, const int64_t dstride[4] /*host*/,
                     const double* L, const double* D, const double* ti, const double* g,
                     const double* G, double* out, int s0, int s1, int mode, void* stream);

int kb200_int_L_strided(int ng, const int32_t dims[4] /*host*/, const int64_t dstride[4] /*host*/,
                        const double* L, int64_t lstride, const double* D, const double* ti,
                        const double* g, const double* G, double* out, int64_t ostride, int s0,
                        int s1, int mode, void* stream);

int kb200_int_L_strided_h(int ng, const int32_t dims[4] /*host*/, const int64_t dstride[4] /*host*/,
                          const double* L, int64_t lstride, const double* D, const double* ti,
                          const double* g, const double* G, double* out, int64_t ostride, int s0,
                          int s1, int mode, const double* ti_h /*host*/, const double* g_h /*host*/,
                          const double* G_h /*host*/, void* stream);

/* ------------------------------------------------------------------------
 * Energy functional pieces.  Replaces kelvin/ft_cc_energy.py:7-32,35-72.
 *   out[0] = sum_y g[y] sum_{p} (c2*T2[y,p] + c11*T1x[y,a,i]*T1y[y,b,j]) * Iabij[p]
 * with p=(a,b,i,j) over dims (nva,nvb,noa,nob); Iabij is the oovv block
 * pre-permuted to abij order (kind-1 plan op).  T1x/T1y may be NULL (c11=0).
 *   kb200_dot_g: out[0] = sum_y g[y] sum_p X[y,p]*F[p]   (the T1.f_ov term)
 * `scratch` needs kb200_reduce_scratch_doubles() doubles. */
int64_t kb200_reduce_scratch_doubles(void);
int kb200_energy_pair(int ng, int nva, int nvb, int noa, int nob,
                      const double* T2, const double* T1x, const double* T1y,
                      const double* Iabij, const double* g, double c2, double c11,
                      double* out, double* scratch, void* stream);
int kb200_dot_g(int ng, int64_t n, const double* X, const double* F, const double* g,
                double* out, double* scratch, void* stream);

/* ------------------------------------------------------------------------
 * Damping + residual norms.  Replaces kelvin/cc_utils.py:145-151,278-295,
 * 458-464,533-547:  out[0]=||new-old||^2, out[1]=||old||^2 (before update),
 * then old <- alpha*old+(1-alpha)*new, out[2]=||old||^2 (after update). */
int kb200_damp_norms(int64_t n, double* old, const double* neu, double alpha,
                     double* out3, double* scratch, void* stream);

/* The same over ng grid points of n elements with row strides (elements): the blocks of one
 * quantity may be column ranges of a wider (ng, Ntot) exchange buffer. */
int kb200_damp_norms_rows(int ng, int64_t n, double* old, int64_t ostride, const double* neu,
                          int64_t nstride, double alpha, double* out3, double* scratch,
                          void* stream);

/* ------------------------------------------------------------------------
 * Integral dressing.  Replaces kelvin/cc_utils.py:584-601,713-775:
 *   out[p,q,r,s] = eri[p,q,r,s]*s0[p]*s1[q]*s2[r]*s3[s]   (dims d[4]);
 *   kb200_dress2: out[p,q] = (f[p,q] - (p==q)*e[p]) * s0[p]*s1[q]. */
int kb200_dress4(const int32_t d[4] /*host*/, const double* eri, const double* s0,
                 const double* s1, const double* s2, const double* s3, double* out,
                 void* stream);
int kb200_dress2(int n0, int n1, const double* f, const double* e, const double* s0,
                 const double* s1, double* out, void* stream);

/* ------------------------------------------------------------------------
 * Active-space (athresh) index subsets.  Replaces the numpy.ix_ gathers of
 * kelvin/cc_utils.py:825-938 (ft_active_integrals / uft_active_integrals) and the
 * numpy.ix_ scatters of :1469-1503,1568-1625 (g/u_n2rdm_full_active):
 *   gather4:      out[i0,i1,i2,i3]  = src[sum_k idx_k[i_k]*src_stride[k]] * prod_k scale_k[i_k]
 *   scatter4_add: dst[sum_k idx_k[i_k]*dst_stride[k]] += alpha * src[i0,i1,i2,i3] * prod_k scale_k[i_k]
 * d[], strides and the two pointer arrays are host arrays; idx_k (int32) / scale_k are device
 * vectors of length d[k], nullptr meaning identity / 1.  Index lists must not repeat entries. */
int kb200_gather4(const int32_t d[4], const int64_t src_stride[4], const double* src,
                  const int32_t* const idx[4], const double* const scale[4], double* out,
                  void* stream);
int kb200_scatter4_add(const int32_t d[4], const int64_t dst_stride[4], const double* src,
                       const int32_t* const idx[4], const double* const scale[4], double alpha,
                       double* dst, void* stream);

/* out[y,p] = sum over nothing: weighted grid sums used by the RDM drivers
 * (kelvin/ft_cc_equations.py:713,737): out[p] = sum_y g[y]*X[y,p]. */
int kb200_gsum(int ng, int64_t n, const double* X, const double* g, double* out, void* stream);

/* "Keep one index" partial trace, replaces the einsum('cdab,abcd->c', P, I)-type
 * contractions of kelvin/cc_utils.py:1648-1685,1746-1895 and the per-grid-point
 * pairings einsum('vijab,vabij->v', L2, T2) of kelvin/ccsd.py:1121-1146,1214-1258:
 *   out[k] = beta*out[k] + alpha * sum_{i1..i4} A[k*sA[0] + sum_d i_d*sA[d]] * B[k*sB[0] + ...]
 * over i_d < dims[d-1]; strides in elements (any index permutation).  nkeep <= 65536. */
int kb200_dot_keep(int nkeep, const int32_t dims[4] /*host*/, const int64_t sA[5] /*host*/,
                   const int64_t sB[5] /*host*/, const double* A, const double* B,
                   double alpha, double beta, double* out, double* scratch, void* stream);

/* Elementwise X[y,p] *= D[p]  (kelvin/ccsd.py:1138-1139,1238-1242). */
int kb200_scale_by(int ng, int64_t n, double* X, const double* D, void* stream);

/* Symmetry defect of two strided 5-index views (element strides; unused leading dims = 1):
 *   out[0] = max |X - Y|, out[1] = max |X|   (device doubles, overwritten).
 * Used for the closed-shell test of unrestricted inputs: Fa == Fb, Ia == Ib,
 * Iabab.wxyz[p,q,r,s] == Iabab.xwzy[q,p,s,r], T1a == T1b, T2aa == T2bb,
 * T2ab[y,a,B,i,J] == T2ab[y,B,a,J,i] -- the condition under which kelvin's unrestricted
 * loops (kelvin/cc_utils.py:245-317, 483-566) compute every beta block twice. */
int kb200_max_absdiff(const int32_t dims[5] /*host*/, const int64_t sx[5] /*host*/,
                      const int64_t sy[5] /*host*/, const double* X, const double* Y,
                      double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
