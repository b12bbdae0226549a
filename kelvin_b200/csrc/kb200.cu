// kb200.cu -- C ABI + streaming kernels of the B200-native FT-CCSD hot path.
// See include/kelvin_b200.h for the contract and the reference call sites.
#include "../../include/kelvin_b200.h"
#ifndef KB200_LK_NW
#define KB200_LK_NW 4
#endif
#ifndef KB200_LK_MINB
#define KB200_LK_MINB 2
#endif
#include "kb200_gemm.cuh"

#include <atomic>
#include <cstdlib>
#include <vector>
#include <cstdio>
#include <cstring>

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return -2;
}
#define KB_CHECK_LAUNCH(where)                               \
    do {                                                     \
        g_launches.fetch_add(1, std::memory_order_relaxed);  \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return cuda_fail(e__, where); \
    } while (0)

constexpr int RED_BLOCKS = 1184;   // 8 CTAs per SM x 148 SMs
constexpr int RED_THREADS = 256;

// ---------------------------------------------------------------------------
// block reduction helpers (warp shuffles; fixed order => deterministic)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double* out /*[NV][gridDim.x]*/) {
    __shared__ double sh[NV][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) sh[i][w] = s;
    }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = (lane < nw) ? sh[i][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) out[(size_t)i * gridDim.x + blockIdx.x] = s;
        }
    }
}

// final pass: out[i] = sum_j part[i][j], one warp per value, fixed order
__global__ void final_sum_kernel(const double* part, int nblocks, int nv, double* out) {
    int i = blockIdx.x;
    if (i >= nv) return;
    double s = 0.0;
    for (int j = threadIdx.x; j < nblocks; j += 32) s += part[(size_t)i * nblocks + j];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[i] = s;
}

// ---------------------------------------------------------------------------
// imaginary-time integration (quadrature.py:292-345)
//
// One thread per amplitude element.  The grid rows of the output are handled YC at a time with
// their accumulators and running weights in REGISTERS; the input rows stream past once per
// chunk straight from global memory (coalesced, independent of the arithmetic, so the loads
// run ahead of the exp/FMA chain).  Nothing scales with ng except the loop lengths: any grid
// size runs (the reference's tests go to ngrid = 2400).  ti and G sit in shared memory when
// they fit (ng <= 72), else they are read through the read-only cache (warp-uniform addresses).
// MODE 1: E_k = exp(D (tau_{k-1} - tau_k)) once per input row and chunk, weights as running
// products (relative deviation <= ng eps); MODE 0: one exp per (x, y) pair, the reference's
// literal formula.  UPD: the fused amplitude update -- the integrated row never goes to memory:
// residual norm, damping, new norm and the energy contraction are done on the spot
// (cc_utils.py:278-299 in one pass over T-bar and T).
// ---------------------------------------------------------------------------
#ifndef KB200_INT_THREADS
#define KB200_INT_THREADS 128
#endif
#ifndef KB200_INT_MINB
#define KB200_INT_MINB 4
#endif
#ifndef KB200_INT_XB
#define KB200_INT_XB 4
#endif
constexpr int INT_THREADS = KB200_INT_THREADS;
constexpr int INT_XB = KB200_INT_XB;  // input rows whose loads are in flight together
constexpr int INT_SMEM_NG = 72;      // ng up to which ti and G are staged in shared memory

struct IntArgs {
    int ng;
    long long n;
    const double* src;      // tbar / L: ng rows
    long long sstride;
    const double* D;
    const double* ti;
    const double* g;        // int_L and the energy term
    const double* G;
    double* out;            // plain: rows r0..r1 of the result;  UPD: the amplitudes (all ng rows)
    long long ostride;
    int r0, r1;
    int lower;
    // D addressed through strides (int_L: D is stored (v.., o..), L (o.., v..))
    int dd[4];
    long long ds[4];
    int dstrided;
    // fused update
    double alpha;
    const double* W;        // energy weights per element (f_ai or <ij||ab> in abij order) or null
    const double* T1x;      // c11 term: T1x[y, a, i] * T1y[y, b, j]; null = none
    const double* T1y;
    long long t1xs, t1ys;
    int nvb, noa, nob;
    double c2, c11;
    double* part;           // [4][gridDim.x]
};

__device__ __forceinline__ double int_D(const IntArgs& a, long long p) {
    if (!a.dstrided) return a.D[p];
    unsigned r = (unsigned)p;                   // n < 2^31 (checked on the host)
    unsigned i3 = r % (unsigned)a.dd[3]; r /= (unsigned)a.dd[3];
    unsigned i2 = r % (unsigned)a.dd[2]; r /= (unsigned)a.dd[2];
    unsigned i1 = r % (unsigned)a.dd[1]; r /= (unsigned)a.dd[1];
    unsigned i0 = r;
    return a.D[i0 * a.ds[0] + i1 * a.ds[1] + i2 * a.ds[2] + i3 * a.ds[3]];
}

template <int YC, int MODE, bool UPD>
__global__ void __launch_bounds__(INT_THREADS, KB200_INT_MINB)
    int_tbar_kernel(const __grid_constant__ IntArgs a) {
    extern __shared__ double sm[];
    const int ng = a.ng;
    const bool staged = ng <= INT_SMEM_NG;
    const double* tis = a.ti;
    const double* Gs = a.G;
    if (staged) {
        double* t_ = sm;
        double* G_ = sm + ng;
        for (int i = threadIdx.x; i < ng; i += INT_THREADS) t_[i] = a.ti[i];
        for (int i = threadIdx.x; i < ng * ng; i += INT_THREADS) G_[i] = a.G[i];
        __syncthreads();
        tis = t_;
        Gs = G_;
    }
    double nrm[4] = {0.0, 0.0, 0.0, 0.0};
    const double oma = 1.0 - a.alpha;
    for (long long p = (long long)blockIdx.x * INT_THREADS + threadIdx.x; p < a.n;
         p += (long long)gridDim.x * INT_THREADS) {
        const double d = int_D(a, p);
        const double* src = a.src + p;
        double esum = 0.0;
        unsigned ia = 0, ib = 0, ii = 0, ij = 0;
        if (UPD && a.T1x != nullptr) {
            unsigned r = (unsigned)p;
            ij = r % (unsigned)a.nob; r /= (unsigned)a.nob;
            ii = r % (unsigned)a.noa; r /= (unsigned)a.noa;
            ib = r % (unsigned)a.nvb; r /= (unsigned)a.nvb;
            ia = r;
        }
        for (int c0 = a.r0; c0 < a.r1; c0 += YC) {
            const int ny = min(YC, a.r1 - c0);
            double acc[YC], w[YC], dg[YC], ov[YC];
            // every load of the chunk that does not depend on arithmetic is issued up front
            // (diagonal inputs, old amplitudes), the lower-triangle inputs in blocks of XB rows:
            // dozens of independent 8-byte loads per thread are in flight instead of one
#pragma unroll
            for (int j = 0; j < YC; ++j) {
                acc[j] = 0.0;
                w[j] = 1.0;
                dg[j] = (j < ny) ? src[(size_t)(c0 + j) * a.sstride] : 0.0;
                ov[j] = (UPD && j < ny) ? a.out[(size_t)(c0 + j) * a.ostride + p] : 0.0;
            }
            // x < y: descending x, the weight of row y picks up one factor per step
            for (int xb = c0 + ny - 2; xb >= 0; xb -= INT_XB) {
                double tbv[INT_XB];
#pragma unroll
                for (int u = 0; u < INT_XB; ++u)
                    tbv[u] = (xb - u >= 0) ? src[(size_t)(xb - u) * a.sstride] : 0.0;
#pragma unroll
                for (int u = 0; u < INT_XB; ++u) {
                    const int x = xb - u;
                    if (x < 0) break;
                    const double tb = tbv[u];
                    double e = 1.0;
                    if (MODE == 1) e = exp(d * (tis[x] - tis[x + 1]));
#pragma unroll
                    for (int j = 0; j < YC; ++j) {
                        const int y = c0 + j;
                        if (j < ny && y > x) {
                            if (MODE == 1) w[j] *= e;
                            const double gw = Gs[y * ng + x];
                            if (gw != 0.0) {
                                const double wt = (MODE == 1) ? w[j] : exp(d * (tis[x] - tis[y]));
                                acc[j] += gw * wt * tb;
                            }
                        }
                    }
                }
            }
            // x >= y: weight 1 (quirk Q4); only x == y survives for a lower-triangular G
#pragma unroll
            for (int j = 0; j < YC; ++j) {
                const int y = c0 + j;
                if (j < ny) {
                    const double gd = Gs[y * ng + y];
                    if (gd != 0.0) acc[j] += gd * dg[j];
                    if (!a.lower)
                        for (int x = y + 1; x < ng; ++x) {
                            const double gw = Gs[y * ng + x];
                            if (gw != 0.0) acc[j] += gw * src[(size_t)x * a.sstride];
                        }
                }
            }
#pragma unroll
            for (int j = 0; j < YC; ++j) {
                const int y = c0 + j;
                if (j < ny) {
                    if (!UPD) {
                        a.out[(size_t)(y - a.r0) * a.ostride + p] = acc[j];
                    } else {
                        const double o = ov[j];
                        const double df = acc[j] - o;
                        nrm[0] += df * df;
                        nrm[1] += o * o;
                        const double u = a.alpha * o + oma * acc[j];
                        a.out[(size_t)y * a.ostride + p] = u;
                        nrm[2] += u * u;
                        if (a.W != nullptr) {
                            double v = a.c2 * u;
                            if (a.T1x != nullptr)
                                v += a.c11 * a.T1x[(size_t)y * a.t1xs + (size_t)ia * a.noa + ii] *
                                     a.T1y[(size_t)y * a.t1ys + (size_t)ib * a.nob + ij];
                            esum += a.g[y] * v;
                        }
                    }
                }
            }
        }
        if (UPD && a.W != nullptr) nrm[3] += esum * a.W[p];
    }
    if (UPD) block_sum_store<4>(nrm, a.part);
}

// Lambda-bar[s] = (1/g_s) sum_y g_y G[y,s] w(s,y) L[y],  w = exp(D (tau_s - tau_y)) for y >= s
// (running product over ascending y), 1 for y < s (quadrature.py:320-345).
template <int YC, int MODE>
__global__ void __launch_bounds__(INT_THREADS, KB200_INT_MINB)
    int_L_kernel(const __grid_constant__ IntArgs a) {
    extern __shared__ double sm[];
    const int ng = a.ng;
    const bool staged = ng <= INT_SMEM_NG;
    const double* tis = a.ti;
    const double* Gs = a.G;
    const double* gs = a.g;
    if (staged) {
        double* t_ = sm;
        double* g_ = sm + ng;
        double* G_ = sm + 2 * ng;
        for (int i = threadIdx.x; i < ng; i += INT_THREADS) {
            t_[i] = a.ti[i];
            g_[i] = a.g[i];
        }
        for (int i = threadIdx.x; i < ng * ng; i += INT_THREADS) G_[i] = a.G[i];
        __syncthreads();
        tis = t_;
        gs = g_;
        Gs = G_;
    }
    for (long long p = (long long)blockIdx.x * INT_THREADS + threadIdx.x; p < a.n;
         p += (long long)gridDim.x * INT_THREADS) {
        const double d = int_D(a, p);
        const double* src = a.src + p;
        for (int c0 = a.r0; c0 < a.r1; c0 += YC) {
            const int ns = min(YC, a.r1 - c0);
            double acc[YC], w[YC];
#pragma unroll
            for (int j = 0; j < YC; ++j) {
                acc[j] = 0.0;
                w[j] = 1.0;
            }
            if (!a.lower) {
                // G[y,s] with y < s is the upper triangle: weight 1
#pragma unroll
                for (int j = 0; j < YC; ++j) {
                    const int s = c0 + j;
                    if (j < ns)
                        for (int y = 0; y < s; ++y) {
                            const double gw = gs[y] * Gs[y * ng + s];
                            if (gw != 0.0) acc[j] += gw * src[(size_t)y * a.sstride];
                        }
                }
            }
            for (int yb = c0; yb < ng; yb += INT_XB) {
                double lvv[INT_XB];
#pragma unroll
                for (int u = 0; u < INT_XB; ++u)
                    lvv[u] = (yb + u < ng) ? src[(size_t)(yb + u) * a.sstride] : 0.0;
#pragma unroll
                for (int u = 0; u < INT_XB; ++u) {
                    const int y = yb + u;
                    if (y >= ng) break;
                    const double lv = lvv[u];
                    double e = 1.0;
                    if (MODE == 1 && y > 0) e = exp(d * (tis[y - 1] - tis[y]));
#pragma unroll
                    for (int j = 0; j < YC; ++j) {
                        const int s = c0 + j;
                        if (j < ns && y >= s) {
                            if (MODE == 1 && y > s) w[j] *= e;
                            const double gw = gs[y] * Gs[y * ng + s];
                            if (gw != 0.0) {
                                const double wt = (MODE == 1) ? w[j] : exp(d * (tis[s] - tis[y]));
                                acc[j] += gw * wt * lv;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < YC; ++j) {
                const int s = c0 + j;
                if (j < ns) a.out[(size_t)(s - a.r0) * a.ostride + p] = acc[j] / gs[s];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// The same integrations with the grid size as a template parameter (NG = 10: the reference's
// benchmarks; 16: UEG-57), all rows, lower-triangular G.  Everything unrolls: the quadrature
// matrix, the grid and the weights sit in the kernel's parameter space and enter the FP64
// instructions as constant-bank operands, the NG inputs (and old amplitudes) of an element are
// NG independent loads issued up front, and what remains per element is NG-1 exps and
// NG(NG-1)/2 multiply + multiply-add pairs -- no index arithmetic, no branches, no shared memory.
// Same arithmetic and summation order as the generic kernels.
// ---------------------------------------------------------------------------
template <int NG>
struct IntFixed {
    double G[NG * NG];      // int_tbar: G[y][x];  int_L: g[y] * G[y][s] at [y][s]
    double ti[NG];
    double g[NG];
};

template <int NG, int MODE, bool UPD>
__global__ void __launch_bounds__(INT_THREADS)
    int_tbar_fixed_kernel(const __grid_constant__ IntArgs a, const __grid_constant__ IntFixed<NG> c) {
    double nrm[4] = {0.0, 0.0, 0.0, 0.0};
    const double oma = 1.0 - a.alpha;
    for (long long p = (long long)blockIdx.x * INT_THREADS + threadIdx.x; p < a.n;
         p += (long long)gridDim.x * INT_THREADS) {
        const double d = a.D[p];
        const double* src = a.src + p;
        double tb[NG], ov[NG], e[NG];
#pragma unroll
        for (int x = 0; x < NG; ++x) tb[x] = src[(size_t)x * a.sstride];
        if (UPD) {
#pragma unroll
            for (int y = 0; y < NG; ++y) ov[y] = a.out[(size_t)y * a.ostride + p];
        }
        unsigned ia = 0, ib = 0, ii = 0, ij = 0;
        if (UPD && a.T1x != nullptr) {
            unsigned r = (unsigned)p;
            ij = r % (unsigned)a.nob; r /= (unsigned)a.nob;
            ii = r % (unsigned)a.noa; r /= (unsigned)a.noa;
            ib = r % (unsigned)a.nvb; r /= (unsigned)a.nvb;
            ia = r;
        }
        e[0] = 1.0;
        if (MODE == 1) {
#pragma unroll
            for (int k = 1; k < NG; ++k) e[k] = exp(d * (c.ti[k - 1] - c.ti[k]));
        }
        double esum = 0.0;
#pragma unroll
        for (int y = 0; y < NG; ++y) {
            double acc = 0.0, w = 1.0;
#pragma unroll
            for (int x = y - 1; x >= 0; --x) {
                double wt;
                if (MODE == 1) {
                    w *= e[x + 1];
                    wt = w;
                } else {
                    wt = exp(d * (c.ti[x] - c.ti[y]));
                }
                acc += c.G[y * NG + x] * wt * tb[x];
            }
            acc += c.G[y * NG + y] * tb[y];
            if (!UPD) {
                a.out[(size_t)y * a.ostride + p] = acc;
            } else {
                const double o = ov[y];
                const double df = acc - o;
                nrm[0] += df * df;
                nrm[1] += o * o;
                const double u = a.alpha * o + oma * acc;
                a.out[(size_t)y * a.ostride + p] = u;
                nrm[2] += u * u;
                if (a.W != nullptr) {
                    double v = a.c2 * u;
                    if (a.T1x != nullptr)
                        v += a.c11 * a.T1x[(size_t)y * a.t1xs + (size_t)ia * a.noa + ii] *
                             a.T1y[(size_t)y * a.t1ys + (size_t)ib * a.nob + ij];
                    esum += c.g[y] * v;
                }
            }
        }
        if (UPD && a.W != nullptr) nrm[3] += esum * a.W[p];
    }
    if (UPD) block_sum_store<4>(nrm, a.part);
}

template <int NG, int MODE>
__global__ void __launch_bounds__(INT_THREADS)
    int_L_fixed_kernel(const __grid_constant__ IntArgs a, const __grid_constant__ IntFixed<NG> c) {
    for (long long p = (long long)blockIdx.x * INT_THREADS + threadIdx.x; p < a.n;
         p += (long long)gridDim.x * INT_THREADS) {
        const double d = int_D(a, p);
        const double* src = a.src + p;
        double lv[NG], e[NG];
#pragma unroll
        for (int y = 0; y < NG; ++y) lv[y] = src[(size_t)y * a.sstride];
        e[0] = 1.0;
        if (MODE == 1) {
#pragma unroll
            for (int k = 1; k < NG; ++k) e[k] = exp(d * (c.ti[k - 1] - c.ti[k]));
        }
#pragma unroll
        for (int s = 0; s < NG; ++s) {
            double acc = 0.0, w = 1.0;
#pragma unroll
            for (int y = s; y < NG; ++y) {
                double wt;
                if (MODE == 1) {
                    if (y > s) w *= e[y];
                    wt = w;
                } else {
                    wt = exp(d * (c.ti[s] - c.ti[y]));
                }
                acc += c.G[y * NG + s] * wt * lv[y];
            }
            a.out[(size_t)s * a.ostride + p] = acc / c.g[s];
        }
    }
}

// ---------------------------------------------------------------------------
// energy functional (ft_cc_energy.py:7-72)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
    energy_pair_kernel(int ng, int nva, int nvb, int noa, int nob,
                       const double* __restrict__ T2, const double* __restrict__ T1x,
                       const double* __restrict__ T1y, const double* __restrict__ I,
                       const double* __restrict__ g, double c2, double c11, double* part) {
    const long long n = (long long)nva * nvb * noa * nob;
    const long long n1x = (long long)nva * noa, n1y = (long long)nvb * nob;
    double acc[1] = {0.0};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned r = (unsigned)p;                  // n < 2^31 (checked on the host)
        unsigned j = r % (unsigned)nob; r /= (unsigned)nob;
        unsigned i = r % (unsigned)noa; r /= (unsigned)noa;
        unsigned bb = r % (unsigned)nvb; r /= (unsigned)nvb;
        unsigned a = r;
        double s = 0.0;
        for (int y = 0; y < ng; ++y) {
            double v = c2 * T2[(size_t)y * n + p];
            if (T1x != nullptr)
                v += c11 * T1x[(size_t)y * n1x + (size_t)a * noa + i] * T1y[(size_t)y * n1y + (size_t)bb * nob + j];
            s += __ldg(g + y) * v;
        }
        acc[0] += s * I[p];
    }
    block_sum_store<1>(acc, part);
}

__global__ void __launch_bounds__(RED_THREADS)
    dot_g_kernel(int ng, long long n, const double* __restrict__ X, const double* __restrict__ F,
                 const double* __restrict__ g, double* part) {
    double acc[1] = {0.0};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int y = 0; y < ng; ++y) s += __ldg(g + y) * X[(size_t)y * n + p];
        acc[0] += s * F[p];
    }
    block_sum_store<1>(acc, part);
}

// ---------------------------------------------------------------------------
// damping + norms (cc_utils.py:145-151,278-295)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
    damp_norms_kernel(long long n, double* old, const double* neu,
                      double alpha, double* part) {
    double acc[3] = {0.0, 0.0, 0.0};
    const double oma = 1.0 - alpha;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        double o = old[p], w = neu[p];
        double d = w - o;
        acc[0] += d * d;
        acc[1] += o * o;
        double u = alpha * o + oma * w;
        old[p] = u;
        acc[2] += u * u;
    }
    block_sum_store<3>(acc, part);
}

// the same over ng rows with row strides (blocks that are column ranges of a wider buffer)
__global__ void __launch_bounds__(RED_THREADS)
    damp_norms_rows_kernel(int ng, long long n, double* old, long long os, const double* neu,
                           long long ns, double alpha, double* part) {
    double acc[3] = {0.0, 0.0, 0.0};
    const double oma = 1.0 - alpha;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        for (int y = 0; y < ng; ++y) {
            double* op = old + (size_t)y * os + p;
            const double o = *op, w = neu[(size_t)y * ns + p];
            const double d = w - o;
            acc[0] += d * d;
            acc[1] += o * o;
            const double u = alpha * o + oma * w;
            *op = u;
            acc[2] += u * u;
        }
    }
    block_sum_store<3>(acc, part);
}

// ---------------------------------------------------------------------------
// dressing (cc_utils.py:584-601,713-775)
// ---------------------------------------------------------------------------
__global__ void dress4_kernel(int d0, int d1, int d2, int d3, const double* __restrict__ eri,
                              const double* __restrict__ s0, const double* __restrict__ s1,
                              const double* __restrict__ s2, const double* __restrict__ s3,
                              double* __restrict__ out) {
    const long long n = (long long)d0 * d1 * d2 * d3;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned r = (unsigned)p;                  // n < 2^31 (checked on the host)
        unsigned l = r % (unsigned)d3; r /= (unsigned)d3;
        unsigned k = r % (unsigned)d2; r /= (unsigned)d2;
        unsigned j = r % (unsigned)d1; r /= (unsigned)d1;
        unsigned i = r;
        out[p] = eri[p] * s0[i] * s1[j] * s2[k] * s3[l];
    }
}

// Index-subset variants for the active-space (athresh) path: every axis may carry an index
// list (device int32, nullptr = identity) into a larger strided array.
struct Sub4 {
    int d[4];
    long long st[4];
    const int* idx[4];
    const double* sc[4];
};

__global__ void gather4_kernel(Sub4 q, const double* __restrict__ src, double* __restrict__ out) {
    const long long n = (long long)q.d[0] * q.d[1] * q.d[2] * q.d[3];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned r = (unsigned)p;
        unsigned i[4];
        i[3] = r % (unsigned)q.d[3]; r /= (unsigned)q.d[3];
        i[2] = r % (unsigned)q.d[2]; r /= (unsigned)q.d[2];
        i[1] = r % (unsigned)q.d[1]; r /= (unsigned)q.d[1];
        i[0] = r;
        long long off = 0;
        double f = 1.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            off += (long long)(q.idx[k] ? q.idx[k][i[k]] : (int)i[k]) * q.st[k];
            if (q.sc[k]) f *= q.sc[k][i[k]];
        }
        out[p] = src[off] * f;
    }
}

__global__ void scatter4_add_kernel(Sub4 q, double alpha, const double* __restrict__ src,
                                    double* __restrict__ dst) {
    const long long n = (long long)q.d[0] * q.d[1] * q.d[2] * q.d[3];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned r = (unsigned)p;
        unsigned i[4];
        i[3] = r % (unsigned)q.d[3]; r /= (unsigned)q.d[3];
        i[2] = r % (unsigned)q.d[2]; r /= (unsigned)q.d[2];
        i[1] = r % (unsigned)q.d[1]; r /= (unsigned)q.d[1];
        i[0] = r;
        long long off = 0;
        double f = alpha;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            off += (long long)(q.idx[k] ? q.idx[k][i[k]] : (int)i[k]) * q.st[k];
            if (q.sc[k]) f *= q.sc[k][i[k]];
        }
        dst[off] += src[p] * f;      // index lists hold distinct entries: no two threads collide
    }
}

__global__ void dress2_kernel(int n0, int n1, const double* __restrict__ f,
                              const double* __restrict__ e, const double* __restrict__ s0,
                              const double* __restrict__ s1, double* __restrict__ out) {
    const long long n = (long long)n0 * n1;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        int j = (int)(p % n1);
        int i = (int)(p / n1);
        double v = f[p];
        if (i == j) v -= e[i];
        out[p] = v * s0[i] * s1[j];
    }
}

__global__ void gsum_kernel(int ng, long long n, const double* __restrict__ X,
                            const double* __restrict__ g, double* __restrict__ out) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int y = 0; y < ng; ++y) s += __ldg(g + y) * X[(size_t)y * n + p];
        out[p] = s;
    }
}

__global__ void scale_by_kernel(int ng, long long n, double* __restrict__ X,
                                const double* __restrict__ D) {
    const long long tot = n * ng;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot;
         q += (long long)gridDim.x * blockDim.x)
        X[q] *= D[q % n];
}

// ---------------------------------------------------------------------------
// Symmetry defect of two strided views (closed-shell check of the unrestricted inputs):
//   out[0] = max |X[i] - Y[i]|, out[1] = max |X[i]| over a 5-index box; out is zeroed by the
//   caller.  Non-negative doubles order like their bit patterns, so the maxima are atomicMax
//   on the unsigned image.
// ---------------------------------------------------------------------------
struct Box5 {
    int d[5];
    long long sx[5];
    long long sy[5];
};

__global__ void max_absdiff_kernel(Box5 q, const double* __restrict__ X,
                                   const double* __restrict__ Y, unsigned long long* out) {
    const long long tot = (long long)q.d[0] * q.d[1] * q.d[2] * q.d[3] * q.d[4];
    double md = 0.0, mx = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < tot;
         p += (long long)gridDim.x * blockDim.x) {
        long long r = p, ox = 0, oy = 0;
#pragma unroll
        for (int k = 4; k >= 0; --k) {
            const long long i = r % q.d[k];
            r /= q.d[k];
            ox += i * q.sx[k];
            oy += i * q.sy[k];
        }
        const double x = X[ox];
        md = fmax(md, fabs(x - Y[oy]));
        mx = fmax(mx, fabs(x));
    }
    for (int o = 16; o > 0; o >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, (unsigned long long)__double_as_longlong(md));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(mx));
    }
}

// ---------------------------------------------------------------------------
// "keep one index" partial traces (cc_utils.py:1648-1685,1746-1895) and the
// per-grid-point <L, T> pairings of ccsd.py:1121-1146,1214-1258:
//   out[k] = beta*out[k] + alpha * sum_{i1..i4} A[k*sa0 + sum i_d*sa_d] * B[k*sb0 + sum i_d*sb_d]
// ---------------------------------------------------------------------------
struct Keep5 {
    int d[4];
    long long sa[5];
    long long sb[5];
};

__global__ void __launch_bounds__(RED_THREADS)
    dot_keep_kernel(Keep5 q, const double* __restrict__ A, const double* __restrict__ B,
                    double* part /*[nkeep][gridDim.x]*/) {
    const int k = blockIdx.y;
    const long long inner = (long long)q.d[0] * q.d[1] * q.d[2] * q.d[3];
    const double* Ak = A + k * q.sa[0];
    const double* Bk = B + k * q.sb[0];
    double acc = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < inner;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned r = (unsigned)p;                  // inner < 2^31 (checked on the host)
        unsigned i4 = r % (unsigned)q.d[3]; r /= (unsigned)q.d[3];
        unsigned i3 = r % (unsigned)q.d[2]; r /= (unsigned)q.d[2];
        unsigned i2 = r % (unsigned)q.d[1]; r /= (unsigned)q.d[1];
        unsigned i1 = r;
        acc += Ak[i1 * q.sa[1] + i2 * q.sa[2] + i3 * q.sa[3] + i4 * q.sa[4]] *
               Bk[i1 * q.sb[1] + i2 * q.sb[2] + i3 * q.sb[3] + i4 * q.sb[4]];
    }
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    acc = warp_sum(acc);
    if (lane == 0) sh[w] = acc;
    __syncthreads();
    if (w == 0) {
        double s2 = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        s2 = warp_sum(s2);
        if (lane == 0) part[(size_t)k * gridDim.x + blockIdx.x] = s2;
    }
}

__global__ void final_axpby_kernel(const double* part, int nblocks, int nv, double alpha,
                                   double beta, double* out) {
    int i = blockIdx.x;
    if (i >= nv) return;
    double s = 0.0;
    for (int j = threadIdx.x; j < nblocks; j += 32) s += part[(size_t)i * nblocks + j];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[i] = (beta != 0.0 ? beta * out[i] : 0.0) + alpha * s;
}

int grid_for(long long n, int threads, int cap = 148 * 16) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

// ---------------------------------------------------------------------------
// GEMM dispatch
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// TMA tensor maps for plain operands
// ---------------------------------------------------------------------------
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

TmaEncodeFn tma_encoder() {
    static TmaEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* off = getenv("KB200_TMA");
        if (off && atoi(off) == 0) return nullptr;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TmaEncodeFn)p;
    }
    return fn;
}

// The operand as a 3-d tensor (contiguous index: nc elements; strided index: ns rows of pitch
// ld; grid point: nb of stride bs) with a box of boxc x boxs x 1.  False when the layout does
// not meet TMA's alignment rules (16-byte base and strides).
bool make_tmap(CUtensorMap* tm, const double* base, long long nc, long long ns, long long ld, int nb,
               long long bs, int boxc, int boxs) {
    TmaEncodeFn enc = tma_encoder();
    if (!enc || ld <= 0 || (ld & 1) || ((uintptr_t)base & 15) || nc <= 0 || ns <= 0) return false;
    if (nb > 1 && (bs <= 0 || (bs & 1))) return false;
    if (nc >= (1LL << 32) || ns >= (1LL << 32) || ld * 8 >= (1LL << 40) || bs * 8 >= (1LL << 40))
        return false;
    if (boxc > 256 || boxs > 256) return false;
    cuuint64_t dims[3] = {(cuuint64_t)nc, (cuuint64_t)ns, (cuuint64_t)(nb > 1 ? nb : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(nb > 1 ? bs : ns * ld) * 8};
    cuuint32_t box[3] = {(cuuint32_t)boxc, (cuuint32_t)boxs, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int WMs, int WNs, int WM, int WN, int ST, bool ILV, int MINB = 1>
int launch_gemm_modes(kb200::GemmGroup& grp, const long long* ldA, const long long* ldB, int splitk,
                      cudaStream_t st) {
    using namespace kb200;
    constexpr int BMt = WMs * WM, BNt = WNs * WN, NT = WMs * WNs * 32;
    constexpr int smem = ST * (StageMax<BMt, NT>::value + StageMax<BNt, NT>::value) * 8 + 2 * ST * 8;
    for (int m = 0; m < grp.n; ++m) {
        GemmParams& p = grp.p[m];
        p.tmaA = p.tmaB = 0;
        if (ldA[m] > 0) {
            // mode 0: [rows][k contiguous], box (BK + 4) x BM; mode 1: [k][rows contiguous],
            // box (BM + 4) x BK -- the padded pitches of TileLoader
            p.tmaA = p.amode == 0
                         ? make_tmap(&grp.tm[m][0], p.A, p.K, p.M, ldA[m], p.bsA ? p.batch : 1, p.bsA,
                                     BK + 4, BMt)
                         : make_tmap(&grp.tm[m][0], p.A, p.M, p.K, ldA[m], p.bsA ? p.batch : 1, p.bsA,
                                     BMt + 4, BK);
        }
        if (ldB[m] > 0) {
            p.tmaB = p.bmode == 0
                         ? make_tmap(&grp.tm[m][1], p.B, p.K, p.N, ldB[m], p.bsB ? p.batch : 1, p.bsB,
                                     BK + 4, BNt)
                         : make_tmap(&grp.tm[m][1], p.B, p.N, p.K, ldB[m], p.bsB ? p.batch : 1, p.bsB,
                                     BNt + 4, BK);
        }
        // the kernel feeds a member either entirely by TMA or entirely by gathers
        if (!(p.tmaA && p.tmaB)) p.tmaA = p.tmaB = 0;
    }
    auto kern = gemm_tab_kernel<WMs, WNs, WM, WN, ST, ILV, MINB>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm)");
        configured = true;
    }
    dim3 grid(grp.fend[grp.n - 1] + grp.rend[grp.n - 1], splitk, 1);
    kern<<<grid, NT, smem, st>>>(grp);
    KB_CHECK_LAUNCH("gemm_tab_kernel");
    return 0;
}

constexpr int LK_NW = KB200_LK_NW;       // warps per CTA of the long-K kernel
constexpr int LK_MINB = KB200_LK_MINB;   // CTAs per SM it is compiled for

template <int T, int AM, int BM_>
int launch_longk_inst(const kb200::GemmParams& p, cudaStream_t st) {
    using namespace kb200;
    constexpr int NT = LK_NW * 32;
    constexpr int smem = LK_STAGES * (LongKLoader<8 * T, AM, NT>::STAGE + LongKLoader<8 * T, BM_, NT>::STAGE) * 8;
    auto kern = longk_kernel<T, AM, BM_, LK_NW, LK_MINB>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(longk)");
        configured = true;
    }
    dim3 grid(p.splitk, p.batch, 1);
    kern<<<grid, NT, smem, st>>>(p);
    KB_CHECK_LAUNCH("longk_kernel");
    return 0;
}

template <int T>
int launch_longk_modes(const kb200::GemmParams& p, int am, int bm, cudaStream_t st) {
    if (am == 0 && bm == 0) return launch_longk_inst<T, 0, 0>(p, st);
    if (am == 0 && bm == 1) return launch_longk_inst<T, 0, 1>(p, st);
    if (am == 1 && bm == 0) return launch_longk_inst<T, 1, 0>(p, st);
    return launch_longk_inst<T, 1, 1>(p, st);
}

template <int KT, int NI>
int launch_skinny(const kb200::GemmParams& p, int a_mode, int c_mode, cudaStream_t st) {
    using namespace kb200;
    constexpr int smem = SkinnyCfg<KT, NI>::SMEM;
    auto kern = skinny_kernel<KT, NI>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(skinny)");
        configured = true;
    }
    // one CTA per SM over all batches; every CTA stages B once and its warps walk 8-row tiles
    int ntiles = (p.M + 7) / 8;
    int gx = 148 / p.batch;            // never more CTAs than SMs: a second wave would double the time
    int maxgx = (ntiles + SK_WARPS - 1) / SK_WARPS;
    if (gx > maxgx) gx = maxgx;
    if (gx < 1) gx = 1;
    dim3 grid(gx, p.batch, 1);
    kern<<<grid, SK_WARPS * 32, smem, st>>>(p, a_mode, c_mode);
    KB_CHECK_LAUNCH("skinny_kernel");
    return 0;
}

// tile ids: 0 = 128x128, 8 warps (64x32 warp tiles), interleaved loads
//           1 = 128x32,  8 warps (16x32)
//           2 = 128x128, 16 warps (32x32 warp tiles), interleaved loads
//           3 = 128x128, 8 warps, loads issued up front (experiment)
//           4 = 128x64,  8 warps (32x32 warp tiles), 2 CTAs/SM
//           5 = 64x64,   4 warps (32x32 warp tiles), 3 stages, several CTAs/SM: bandwidth-bound
//               shapes (skinny N, or tiny outputs with long split K)
//           6 = whole output (<= 40x40) per CTA, split over K only (longk_kernel)
//           7 = skinny streaming update, K, N <= 40, 8 rows per warp (skinny_kernel)
int tile_bm(int tile) { return tile == 5 ? 64 : 128; }
int tile_bn(int tile) { return tile == 1 ? 32 : ((tile == 4 || tile == 5) ? 64 : 128); }

int64_t op_workspace(const kb200_op& o) {
    if (o.kind != 0 || o.splitk <= 1) return 0;   // kinds 1, 2 need no workspace
    return (int64_t)o.batch * o.splitk * (int64_t)o.M * o.N * 8;
}

}  // namespace

namespace {

int int_grid(long long n) { return grid_for(n, INT_THREADS, 148 * 16); }

size_t int_smem(int ng, int nvec) {
    return ng <= INT_SMEM_NG ? ((size_t)ng * ng + (size_t)nvec * ng) * 8 : 0;
}

template <int YC, bool UPD>
void launch_int_tbar(const IntArgs& a, int mode, cudaStream_t st) {
    const int grid = int_grid(a.n);
    const size_t smem = int_smem(a.ng, 1);
    if (mode == 1)
        int_tbar_kernel<YC, 1, UPD><<<grid, INT_THREADS, smem, st>>>(a);
    else
        int_tbar_kernel<YC, 0, UPD><<<grid, INT_THREADS, smem, st>>>(a);
}

template <int YC>
void launch_int_L(const IntArgs& a, int mode, cudaStream_t st) {
    const int grid = int_grid(a.n);
    const size_t smem = int_smem(a.ng, 2);
    if (mode == 1)
        int_L_kernel<YC, 1><<<grid, INT_THREADS, smem, st>>>(a);
    else
        int_L_kernel<YC, 0><<<grid, INT_THREADS, smem, st>>>(a);
}

// Fixed-grid fast path: host copies of ti / g / G given, all rows, G lower triangular.
template <int NG>
bool fill_fixed(IntFixed<NG>& c, const double* ti_h, const double* g_h, const double* G_h, bool forL) {
    for (int y = 0; y < NG; ++y) {
        c.ti[y] = ti_h[y];
        c.g[y] = g_h ? g_h[y] : 0.0;
        for (int x = 0; x < NG; ++x) {
            if (x > y && G_h[y * NG + x] != 0.0) return false;      // not lower triangular
            c.G[y * NG + x] = forL ? g_h[y] * G_h[y * NG + x] : G_h[y * NG + x];
        }
    }
    return true;
}

template <int NG, bool UPD>
bool launch_int_tbar_fixed(const IntArgs& a, int mode, const double* ti_h, const double* g_h,
                           const double* G_h, cudaStream_t st) {
    IntFixed<NG> c;
    if (!fill_fixed<NG>(c, ti_h, g_h, G_h, false)) return false;
    const int grid = int_grid(a.n);
    if (mode == 1)
        int_tbar_fixed_kernel<NG, 1, UPD><<<grid, INT_THREADS, 0, st>>>(a, c);
    else
        int_tbar_fixed_kernel<NG, 0, UPD><<<grid, INT_THREADS, 0, st>>>(a, c);
    return true;
}

template <int NG>
bool launch_int_L_fixed(const IntArgs& a, int mode, const double* ti_h, const double* g_h,
                        const double* G_h, cudaStream_t st) {
    IntFixed<NG> c;
    if (!fill_fixed<NG>(c, ti_h, g_h, G_h, true)) return false;
    const int grid = int_grid(a.n);
    if (mode == 1)
        int_L_fixed_kernel<NG, 1><<<grid, INT_THREADS, 0, st>>>(a, c);
    else
        int_L_fixed_kernel<NG, 0><<<grid, INT_THREADS, 0, st>>>(a, c);
    return true;
}

IntArgs int_args(int ng, int64_t n, const double* src, int64_t sstride, const double* D,
                 const double* ti, const double* g, const double* G, double* out,
                 int64_t ostride, int r0, int r1, int lower) {
    IntArgs a;
    memset(&a, 0, sizeof(a));
    a.ng = ng; a.n = n; a.src = src; a.sstride = sstride; a.D = D; a.ti = ti; a.g = g; a.G = G;
    a.out = out; a.ostride = ostride; a.r0 = r0; a.r1 = r1; a.lower = lower;
    return a;
}

}  // namespace

extern "C" {

int kb200_version(void) { return 100; }
const char* kb200_last_error(void) { return g_err; }
int64_t kb200_launch_count(void) { return g_launches.load(); }
void kb200_launch_count_reset(void) { g_launches.store(0); }
void kb200_launch_count_add(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t kb200_reduce_scratch_doubles(void) { return 65536; }

int kb200_dot_keep(int nkeep, const int32_t dims[4], const int64_t sA[5], const int64_t sB[5],
                   const double* A, const double* B, double alpha, double beta, double* out,
                   double* scratch, void* stream) {
    if (nkeep <= 0) return fail(-1, "dot_keep: bad nkeep");
    Keep5 q;
    long long inner = 1;
    for (int i = 0; i < 4; ++i) {
        if (dims[i] <= 0) return fail(-1, "dot_keep: bad dims");
        q.d[i] = dims[i];
        inner *= dims[i];
    }
    if (inner >= (1LL << 31)) return fail(-1, "dot_keep: inner extent too large");
    for (int i = 0; i < 5; ++i) {
        q.sa[i] = sA[i];
        q.sb[i] = sB[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long want = (inner + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
    long long cap = 65536 / nkeep;
    if (cap < 1) return fail(-1, "dot_keep: nkeep too large");
    int nsplit = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    if (nsplit > 1024) nsplit = 1024;
    dim3 grid(nsplit, nkeep);
    dot_keep_kernel<<<grid, RED_THREADS, 0, st>>>(q, A, B, scratch);
    KB_CHECK_LAUNCH("dot_keep_kernel");
    final_axpby_kernel<<<nkeep, 32, 0, st>>>(scratch, nsplit, nkeep, alpha, beta, out);
    KB_CHECK_LAUNCH("final_axpby_kernel");
    return 0;
}

// Split-K partials: the wide launches share one region (they are ordered against each other
// through it), the small launches have one region per small-launch stream so that independent
// long-K reductions do not queue behind one another.
constexpr int MS_NBIG = 2, MS_NSMALL_MAX = 6;

static bool op_is_wide(const kb200_op& o) {
    return o.kind == 0 && (o.tile == 0 || o.tile == 2 || o.tile == 3);
}

static void workspace_layout(const kb200_op* ops, int nops, int64_t* wide, int64_t* small) {
    int64_t w = 0, v = 0;
    for (int i = 0; i < nops; ++i) {
        int64_t x = op_workspace(ops[i]);
        if (op_is_wide(ops[i])) { if (x > w) w = x; }
        else if (x > v) v = x;
    }
    *wide = (w + 255) & ~(int64_t)255;
    *small = (v + 255) & ~(int64_t)255;
}

int64_t kb200_plan_workspace_bytes(const kb200_op* ops, int nops) {
    int64_t w, v;
    workspace_layout(ops, nops, &w, &v);
    return w + MS_NSMALL_MAX * v;
}

// ---------------------------------------------------------------------------
// Concurrent execution of a plan's independent launches.
//
// The plan arrives in a dependency-respecting order (plan.py: Lowered._schedule).  With one
// stream every launch waits for the tail of the one before it: at small tau batches (tau-sharded
// runs: 1-3 grid points per GPU) an m^6 launch is only 1-3 waves of CTAs and the ~100 small
// kernels of a residual each fill a fraction of the machine.  Here the big contraction launches
// alternate between two streams and everything else goes to a third; read-after-write,
// write-after-write and write-after-read orders on every slot (and on the split-K workspace) are
// enforced with events, so the arithmetic and its order per output are unchanged
// (bit-identical results).  The caller's stream is stream 0; the side streams fork from it at
// the start of the plan and join it at the end, so the call keeps stream semantics.
// ---------------------------------------------------------------------------
struct MultiStream {
    static constexpr int NS = MS_NBIG + MS_NSMALL_MAX;
    int nsmall = MS_NSMALL_MAX;                     // small-launch streams in use (1 .. MS_NSMALL_MAX)
    cudaStream_t s[NS] = {nullptr};
    bool made = false;
    std::vector<cudaEvent_t> pool;      // one event per launch of the current plan
    int used = 0;
    // per launch
    std::vector<int> l_stream, l_seq;
    // per slot (then one pseudo-slot per workspace region: wide, small stream 0, 1, ...)
    std::vector<int> last_writer;               // launch id or -1
    std::vector<int> last_reader;               // [slot*NS + stream] launch id or -1
    int seq[NS];                                // launches issued per stream
    int synced[NS][NS];                         // synced[S][T]: S has waited for T's launch seq <= this
    int last_on[NS];                            // last launch id per stream
    int nbig = 0;
    cudaEvent_t fork_ev = nullptr;
    bool forked[NS];

    int init(cudaStream_t caller, int nslots, int nstreams) {
        if (!made) {
            // the side streams inherit the priority of the caller's stream (a plan on a
            // high-priority stream must not queue its launches behind another plan's)
            int prio = 0;
            if (cudaStreamGetPriority(caller, &prio) != cudaSuccess) prio = 0;
            for (int k = 1; k < NS; ++k)
                if (cudaStreamCreateWithPriority(&s[k], cudaStreamNonBlocking, prio) != cudaSuccess)
                    return fail(-2, "plan: cannot create side stream");
            if (cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming) != cudaSuccess)
                return fail(-2, "plan: cannot create event");
            made = true;
        }
        nsmall = nstreams - MS_NBIG;
        if (nsmall < 1) nsmall = 1;
        if (nsmall > MS_NSMALL_MAX) nsmall = MS_NSMALL_MAX;
        s[0] = caller;
        used = 0;
        nbig = 0;
        l_stream.clear();
        l_seq.clear();
        last_writer.assign(nslots + 1 + MS_NSMALL_MAX, -1);
        last_reader.assign((size_t)(nslots + 1 + MS_NSMALL_MAX) * NS, -1);
        for (int a = 0; a < NS; ++a) {
            seq[a] = 0;
            last_on[a] = -1;
            forked[a] = (a == 0);
            for (int b = 0; b < NS; ++b) synced[a][b] = 0;
        }
        if (cudaEventRecord(fork_ev, caller) != cudaSuccess) return fail(-2, "plan: fork record");
        return 0;
    }
    int wait_for(int S, int launch) {
        if (launch < 0) return 0;
        const int T = l_stream[launch];
        if (T == S || l_seq[launch] <= synced[S][T]) return 0;
        if (cudaStreamWaitEvent(s[S], pool[launch], 0) != cudaSuccess)
            return fail(-2, "plan: stream wait");
        synced[S][T] = l_seq[launch];
        return 0;
    }
    // Small launches: the stream whose tail is one of this launch's own dependencies (queueing
    // behind it costs nothing), else the least recently used small stream.  Inside a captured
    // graph the streams only define edges: this keeps the edges close to the true dependencies.
    int pick_small(const kb200_op* ops, int i, int n) const {
        int best = -1, best_id = -1;
        for (int m = 0; m < n; ++m) {
            const kb200_op& q = ops[i + m];
            int dep[3] = {last_writer[q.c], last_writer[q.a],
                          (q.kind != 1 && q.b >= 0) ? last_writer[q.b] : -1};
            for (int k = 0; k < 3; ++k) {
                const int id = dep[k];
                if (id < 0) continue;
                const int T = l_stream[id];
                if (T >= MS_NBIG && T < MS_NBIG + nsmall && last_on[T] == id && id > best_id) {
                    best = T;
                    best_id = id;
                }
            }
        }
        if (best >= 0) return best;
        best = MS_NBIG;
        for (int T = MS_NBIG; T < MS_NBIG + nsmall; ++T)
            if (last_on[T] < last_on[best]) best = T;
        return best;
    }
    static int workspace_slot(int nslots, int S) { return S < MS_NBIG ? nslots : nslots + 1 + (S - MS_NBIG); }
    // stream for the launch made of ops[i .. i+n): inserts the waits it needs
    int begin(const kb200_op* ops, int i, int n, int nslots, cudaStream_t* out) {
        const kb200_op& o = ops[i];
        for (int m = 0; m < n; ++m) {
            const kb200_op& q = ops[i + m];
            if (q.a < 0 || q.a >= nslots || q.c < 0 || q.c >= nslots ||
                (q.kind != 1 && (q.b >= nslots || (q.kind != 3 && q.b < 0))))
                return fail(-1, "plan: bad slot");
        }
        const int S = op_is_wide(o) ? (nbig++ & 1) : pick_small(ops, i, n);
        if (!forked[S]) {
            if (cudaStreamWaitEvent(s[S], fork_ev, 0) != cudaSuccess)
                return fail(-2, "plan: fork wait");
            forked[S] = true;
        }
        for (int m = 0; m < n; ++m) {
            const kb200_op& q = ops[i + m];
            int rd[3], nr = 0;
            rd[nr++] = q.a;
            if (q.kind != 1 && q.b >= 0) rd[nr++] = q.b;
            if (q.beta != 0.0) rd[nr++] = q.c;
            for (int k = 0; k < nr; ++k) {
                if (rd[k] < 0 || rd[k] >= nslots) return fail(-1, "plan: bad slot");
                if (wait_for(S, last_writer[rd[k]])) return -2;
            }
            int wr[2], nw = 0;
            if (q.c < 0 || q.c >= nslots) return fail(-1, "plan: bad slot");
            wr[nw++] = q.c;
            if (op_workspace(q) > 0) wr[nw++] = workspace_slot(nslots, S);   // split-K partials
            for (int k = 0; k < nw; ++k) {
                if (wait_for(S, last_writer[wr[k]])) return -2;
                for (int T = 0; T < NS; ++T)
                    if (wait_for(S, last_reader[(size_t)wr[k] * NS + T])) return -2;
            }
        }
        *out = s[S];
        return S;
    }
    int end(const kb200_op* ops, int i, int n, int nslots, int S) {
        const int id = (int)l_stream.size();
        if ((int)pool.size() <= id) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
                return fail(-2, "plan: cannot create event");
            pool.push_back(e);
        }
        if (cudaEventRecord(pool[id], s[S]) != cudaSuccess) return fail(-2, "plan: event record");
        l_stream.push_back(S);
        l_seq.push_back(++seq[S]);
        last_on[S] = id;
        for (int m = 0; m < n; ++m) {
            const kb200_op& q = ops[i + m];
            last_reader[(size_t)q.a * NS + S] = id;
            if (q.kind != 1 && q.b >= 0) last_reader[(size_t)q.b * NS + S] = id;
        }
        for (int m = 0; m < n; ++m) {
            const kb200_op& q = ops[i + m];
            int wr[2], nw = 0;
            wr[nw++] = q.c;
            if (op_workspace(q) > 0) wr[nw++] = workspace_slot(nslots, S);
            for (int k = 0; k < nw; ++k) {
                last_writer[wr[k]] = id;
                for (int T = 0; T < NS; ++T) last_reader[(size_t)wr[k] * NS + T] = -1;
            }
        }
        return 0;
    }
    // the caller's stream waits for everything issued on the side streams
    int join() {
        int rc = 0;
        for (int T = 1; T < NS; ++T)
            if (last_on[T] >= 0 &&
                cudaStreamWaitEvent(s[0], pool[last_on[T]], 0) != cudaSuccess)
                rc = fail(-2, "plan: join wait");
        return rc;
    }
};

// one set of side streams per (device, caller stream): two plans that run concurrently on
// different caller streams (a rank's own grid points and the grid point it shares with the other
// ranks, parallel.py) must not queue behind each other on shared side streams
constexpr int MS_PER_DEV = 4;
static MultiStream g_ms[16][MS_PER_DEV];
static cudaStream_t g_ms_owner[16][MS_PER_DEV];
static int g_ms_used[16] = {0};

static MultiStream* ms_for(int dev, cudaStream_t caller) {
    for (int k = 0; k < g_ms_used[dev]; ++k)
        if (g_ms_owner[dev][k] == caller) return &g_ms[dev][k];
    if (g_ms_used[dev] < MS_PER_DEV) {
        const int k = g_ms_used[dev]++;
        g_ms_owner[dev][k] = caller;
        return &g_ms[dev][k];
    }
    return nullptr;          // more caller streams than sets: that plan runs on one stream
}

static int g_plan_streams = -1;

// 1 = everything on the caller's stream; n >= 3 = two streams for the wide launches plus n - 2
// for the small ones; anything else = the default (2 + 6)
static int clamp_streams(int n) {
    if (n == 1) return 1;
    if (n < 3) return MS_NBIG + MS_NSMALL_MAX;
    return n > MS_NBIG + MS_NSMALL_MAX ? MS_NBIG + MS_NSMALL_MAX : n;
}

static int plan_streams() {
    if (g_plan_streams < 0) {
        const char* e = getenv("KB200_STREAMS");
        g_plan_streams = clamp_streams(e ? atoi(e) : 0);
    }
    return g_plan_streams;
}

int kb200_set_plan_streams(int n) {
    const int old = plan_streams();
    g_plan_streams = clamp_streams(n);
    return old;
}

static int run_plan_body(const kb200_op* ops, int nops, const uint32_t* tables,
                         double* const* slots, int nslots, double* workspace,
                         int64_t workspace_bytes, cudaStream_t st0, cudaEvent_t* ev,
                         MultiStream* ms) {
    int prev_i = -1, prev_n = 0, prev_S = 0;
    int64_t ws_wide = 0, ws_small = 0;
    workspace_layout(ops, nops, &ws_wide, &ws_small);
    if (ws_wide + MS_NSMALL_MAX * ws_small > 0 &&
        (workspace == nullptr || ws_wide + MS_NSMALL_MAX * ws_small > workspace_bytes))
        return fail(-1, "plan: workspace too small");
    for (int i = 0; i < nops; ++i) {
        const kb200_op& o = ops[i];
        cudaStream_t st = st0;
        int S_cur = 0;
        if (ms) {
            if (prev_i >= 0 && ms->end(ops, prev_i, prev_n, nslots, prev_S)) return -2;
            int n = ((o.kind == 0 || o.kind == 3) && o.group > 1) ? o.group : 1;
            if (i + n > nops) return fail(-1, "plan: bad group");
            int S = ms->begin(ops, i, n, nslots, &st);
            if (S < 0) return S;
            prev_i = i; prev_n = n; prev_S = S;
            S_cur = S;
        }
        // split-K partials of this launch: the wide region, or the region of its small stream
        double* ws_op = workspace;
        if (workspace != nullptr && !op_is_wide(o))
            ws_op = workspace + (ws_wide + (int64_t)(S_cur >= MS_NBIG ? S_cur - MS_NBIG : 0) * ws_small) / 8;
        if (ev) cudaEventRecord(ev[2 * i], st);
        if (o.a < 0 || o.a >= nslots || o.c < 0 || o.c >= nslots) return fail(-1, "plan: bad slot");
        if (o.M <= 0 || o.N <= 0 || o.batch <= 0) return fail(-1, "plan: empty op");
        if (o.kind == 0) {
            // a group leader carries the number of consecutive, mutually independent ops of the
            // same kernel configuration that share its launch
            int ng = o.group > 1 ? o.group : 1;
            if (ng > kb200::MAX_GROUP || i + ng > nops) return fail(-1, "plan: bad group");
            kb200::GemmGroup grp;
            long long ldA[kb200::MAX_GROUP], ldB[kb200::MAX_GROUP];
            grp.n = ng;
            int BMt = tile_bm(o.tile), BNt = tile_bn(o.tile);
            int fsum = 0, rsum = 0;
            for (int m = 0; m < ng; ++m) {
                const kb200_op& q = ops[i + m];
                if (q.kind != 0 || q.tile != o.tile || (ng > 1 && q.splitk > 1))
                    return fail(-1, "plan: inconsistent group");
                if (q.a < 0 || q.a >= nslots || q.c < 0 || q.c >= nslots || q.b < 0 || q.b >= nslots ||
                    q.K <= 0 || q.M <= 0 || q.N <= 0 || q.batch <= 0)
                    return fail(-1, "plan: bad contraction");
                kb200::GemmParams& p = grp.p[m];
                p.A = slots[q.a] + q.a_off;
                p.B = slots[q.b] + q.b_off;
                p.C = slots[q.c] + q.c_off;
                p.am = tables + q.tAm; p.ak = tables + q.tAk;
                p.bk = tables + q.tBk; p.bn = tables + q.tBn;
                p.cm = tables + q.tCm; p.cn = tables + q.tCn;
                p.M = q.M; p.N = q.N; p.K = q.K;
                p.splitk = q.splitk < 1 ? 1 : q.splitk;
                int nkt = (q.K + kb200::BK - 1) / kb200::BK;
                if (p.splitk > nkt) p.splitk = nkt;
                p.kchunk = ((nkt + p.splitk - 1) / p.splitk) * kb200::BK;
                p.splitk = (q.K + p.kchunk - 1) / p.kchunk;
                p.bsA = q.bsA; p.bsB = q.bsB; p.bsC = q.bsC;
                p.alpha = q.alpha; p.beta = q.beta;
                p.partial = ws_op;
                p.tilesM = (q.M + BMt - 1) / BMt;
                p.tilesN = (q.N + BNt - 1) / BNt;
                p.batch = q.batch;
                p.amode = q.a_mode;
                p.bmode = q.b_mode;
                ldA[m] = q.ldA;
                ldB[m] = q.ldB;
                int fm = p.tilesM - ((q.M % BMt) ? 1 : 0), fn = p.tilesN - ((q.N % BNt) ? 1 : 0);
                fsum += fm * fn * q.batch;
                rsum += (p.tilesM * p.tilesN - fm * fn) * q.batch;
                grp.fend[m] = fsum;
                grp.rend[m] = rsum;
            }
            for (int m = ng; m < kb200::MAX_GROUP; ++m) {
                grp.p[m] = grp.p[0];
                grp.fend[m] = fsum;
                grp.rend[m] = rsum;
            }
            const kb200::GemmParams& p = grp.p[0];
            if (p.splitk > 1) {
                int64_t need = (int64_t)o.batch * p.splitk * (int64_t)o.M * o.N * 8;
                if (ws_op == nullptr || need > (op_is_wide(o) ? ws_wide : ws_small))
                    return fail(-1, "plan: workspace too small");
            }
            int rc;
            if (o.tile == 0)
                rc = launch_gemm_modes<2, 4, 64, 32, 4, true>(grp, ldA, ldB, p.splitk, st);
            else if (o.tile == 1)
                rc = launch_gemm_modes<8, 1, 16, 32, 4, true>(grp, ldA, ldB, p.splitk, st);
            else if (o.tile == 2)
                rc = launch_gemm_modes<4, 4, 32, 32, 4, true>(grp, ldA, ldB, p.splitk, st);
            else if (o.tile == 3)
                rc = launch_gemm_modes<2, 4, 64, 32, 4, false>(grp, ldA, ldB, p.splitk, st);
            else if (o.tile == 4)
                rc = launch_gemm_modes<4, 2, 32, 32, 3, true, 2>(grp, ldA, ldB, p.splitk, st);
            else if (o.tile == 5)
                rc = launch_gemm_modes<2, 2, 32, 32, 3, true, 3>(grp, ldA, ldB, p.splitk, st);
            else if (o.tile == 6) {
                if (ng != 1 || o.M > 40 || o.N > 40) return fail(-1, "plan: bad long-K op");
                if (o.M <= 24 && o.N <= 24)
                    rc = launch_longk_modes<3>(p, o.a_mode, o.b_mode, st);
                else
                    rc = launch_longk_modes<5>(p, o.a_mode, o.b_mode, st);
            } else if (o.tile == 7) {
                if (ng != 1 || o.K > 40 || o.N > 40 || p.splitk != 1) return fail(-1, "plan: bad skinny op");
                // reserved bit 0: consecutive rows of C are adjacent in memory (else its columns are)
                const int cmode = o.reserved & 1;
                if (o.K <= 20 && o.N <= 24)
                    rc = launch_skinny<5, 3>(p, o.a_mode, cmode, st);
                else if (o.K <= 36)
                    rc = launch_skinny<9, 5>(p, o.a_mode, cmode, st);
                else
                    rc = launch_skinny<10, 5>(p, o.a_mode, cmode, st);
            } else
                return fail(-1, "plan: unknown tile id");
            if (rc) return rc;
            if (p.splitk > 1) {
                long long total = (long long)o.M * o.N * o.batch;
                if (p.splitk >= 32 && total <= 32LL * 148 * 8)
                    kb200::splitk_reduce_wide_kernel<<<(unsigned)((total + 31) / 32), dim3(32, 16), 0, st>>>(p, o.batch);
                else
                    kb200::splitk_reduce_kernel<<<grid_for(total, 256), 256, 0, st>>>(p, o.batch);
                KB_CHECK_LAUNCH("splitk_reduce_kernel");
            }
            if (ev) {
                // the group's time is recorded on its leader; the other members read 0
                cudaEventRecord(ev[2 * i + 1], st);
                for (int m = 1; m < ng; ++m) {
                    cudaEventRecord(ev[2 * (i + m)], st);
                    cudaEventRecord(ev[2 * (i + m) + 1], st);
                }
            }
            i += ng - 1;
            continue;
        } else if (o.kind == 2) {
            if (o.b < 0 || o.b >= nslots || o.K <= 0 || o.K > 64 || o.N > 64)
                return fail(-1, "plan: bad rank-k op");
            kb200::GemmParams p;
            p.A = slots[o.a] + o.a_off;
            p.B = slots[o.b] + o.b_off;
            p.C = slots[o.c] + o.c_off;
            p.am = tables + o.tAm; p.ak = tables + o.tAk;
            p.bk = tables + o.tBk; p.bn = tables + o.tBn;
            p.cm = tables + o.tCm; p.cn = tables + o.tCn;
            p.M = o.M; p.N = o.N; p.K = o.K;
            p.splitk = 1; p.kchunk = 0;
            p.bsA = o.bsA; p.bsB = o.bsB; p.bsC = o.bsC;
            p.alpha = o.alpha; p.beta = o.beta;
            p.partial = nullptr; p.tilesM = p.tilesN = 0; p.batch = o.batch;
            // each CTA stages B once and streams several 256-row blocks: aim at ~4 CTAs per SM
            int nblk = (o.M + 255) / 256;
            int per = (int)(((long long)nblk * o.batch + 148 * 4 - 1) / (148 * 4));
            if (per < 1) per = 1;
            if (per > 16) per = 16;
            dim3 grid((nblk + per - 1) / per, 1, o.batch);
            if (o.K <= 40)
                kb200::rankk_kernel<40><<<grid, 256, 0, st>>>(p, per);
            else
                kb200::rankk_kernel<64><<<grid, 256, 0, st>>>(p, per);
            KB_CHECK_LAUNCH("rankk_kernel");
        } else if (o.kind == 3) {
            int ng = o.group > 1 ? o.group : 1;
            if (ng > kb200::EW_MAX_TERMS || i + ng > nops) return fail(-1, "plan: bad fused group");
            kb200::EwParams q;
            q.nterms = ng;
            q.C = slots[o.c] + o.c_off;
            q.cm = tables + o.tCm; q.cn = tables + o.tCn;
            q.bsC = o.bsC; q.beta = o.beta; q.M = o.M; q.N = o.N;
            for (int m = 0; m < ng; ++m) {
                const kb200_op& r = ops[i + m];
                if (r.kind != 3 || r.c != o.c || r.M != o.M || r.N != o.N || r.tCm != o.tCm ||
                    r.tCn != o.tCn || r.a < 0 || r.a >= nslots || r.b >= nslots)
                    return fail(-1, "plan: inconsistent fused group");
                kb200::EwTerm& t = q.t[m];
                t.X = slots[r.a] + r.a_off;
                t.xm = tables + r.tAm; t.xn = tables + r.tAk;
                t.bsX = r.bsA;
                if (r.b >= 0) {
                    t.Y = slots[r.b] + r.b_off;
                    t.ym = tables + r.tBk; t.yn = tables + r.tBn;
                    t.bsY = r.bsB;
                } else {
                    t.Y = nullptr; t.ym = t.yn = nullptr; t.bsY = 0;
                }
                t.alpha = r.alpha;
                t.transposed = r.a_mode;
            }
            for (int m = ng; m < kb200::EW_MAX_TERMS; ++m) q.t[m] = q.t[0];
            dim3 grid((o.M + 31) / 32, (o.N + 31) / 32, o.batch);
            if (grid.y > 65535) return fail(-1, "plan: fused N too large");
            kb200::fused_ew_kernel<<<grid, 256, 0, st>>>(q);
            KB_CHECK_LAUNCH("fused_ew_kernel");
            if (ev) {
                cudaEventRecord(ev[2 * i + 1], st);
                for (int m = 1; m < ng; ++m) {
                    cudaEventRecord(ev[2 * (i + m)], st);
                    cudaEventRecord(ev[2 * (i + m) + 1], st);
                }
            }
            i += ng - 1;
            continue;
        } else if (o.kind == 1) {
            kb200::PermParams p;
            p.A = slots[o.a] + o.a_off;
            p.C = slots[o.c] + o.c_off;
            p.am = tables + o.tAm; p.an = tables + o.tAk;
            p.cm = tables + o.tCm; p.cn = tables + o.tCn;
            p.M = o.M; p.N = o.N;
            p.bsA = o.bsA; p.bsC = o.bsC;
            p.alpha = o.alpha; p.beta = o.beta;
            p.a_mode = o.a_mode; p.c_mode = o.b_mode;
            dim3 grid((o.M + 31) / 32, (o.N + 31) / 32, o.batch);
            if (grid.y > 65535) return fail(-1, "plan: permute N too large");
            kb200::permute_axpby_kernel<<<grid, 256, 0, st>>>(p);
            KB_CHECK_LAUNCH("permute_axpby_kernel");
        } else {
            return fail(-1, "plan: unknown op kind");
        }
        if (ev) cudaEventRecord(ev[2 * i + 1], st);
    }
    if (ms && prev_i >= 0 && ms->end(ops, prev_i, prev_n, nslots, prev_S)) return -2;
    return 0;
}

static int run_plan_impl(const kb200_op* ops, int nops, const uint32_t* tables,
                         double* const* slots, int nslots, double* workspace,
                         int64_t workspace_bytes, cudaStream_t st, cudaEvent_t* ev) {
    MultiStream* ms = nullptr;
    int dev = 0;
    if (!ev && nops > 1 && plan_streams() > 1 && cudaGetDevice(&dev) == cudaSuccess && dev < 16) {
        ms = ms_for(dev, st);
        if (ms && ms->init(st, nslots, plan_streams())) return -2;
    }
    int rc = run_plan_body(ops, nops, tables, slots, nslots, workspace, workspace_bytes, st, ev, ms);
    // join even after an error: nothing may stay in flight on the side streams unordered
    if (ms && ms->join() && rc == 0) rc = -2;
    if (ms && rc != 0) cudaDeviceSynchronize();
    return rc;
}

int kb200_plan_run(const kb200_op* ops, int nops, const uint32_t* tables, double* const* slots,
                   int nslots, double* workspace, int64_t workspace_bytes, void* stream) {
    return run_plan_impl(ops, nops, tables, slots, nslots, workspace, workspace_bytes,
                         (cudaStream_t)stream, nullptr);
}

int kb200_plan_run_timed(const kb200_op* ops, int nops, const uint32_t* tables,
                         double* const* slots, int nslots, double* workspace,
                         int64_t workspace_bytes, void* stream, float* op_ms) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t* ev = new cudaEvent_t[2 * (size_t)nops];
    for (int i = 0; i < 2 * nops; ++i) cudaEventCreate(&ev[i]);
    int rc = run_plan_impl(ops, nops, tables, slots, nslots, workspace, workspace_bytes, st, ev);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == 0 && e != cudaSuccess) rc = cuda_fail(e, "plan_run_timed sync");
    for (int i = 0; i < nops; ++i) {
        float ms = 0.f;
        if (rc == 0) cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
        op_ms[i] = ms;
    }
    for (int i = 0; i < 2 * nops; ++i) cudaEventDestroy(ev[i]);
    delete[] ev;
    return rc;
}

int kb200_int_tbar(int ng, int64_t n, const double* tbar, const double* D, const double* ti,
                   const double* G, double* out, int mode, void* stream) {
    return kb200_int_tbar_strided(ng, n, tbar, n, D, ti, G, out, n, 0, ng, mode, stream);
}

int kb200_int_tbar_rows(int ng, int64_t n, const double* tbar, const double* D, const double* ti,
                        const double* G, double* out, int y0, int y1, int mode, void* stream) {
    return kb200_int_tbar_strided(ng, n, tbar, n, D, ti, G, out, n, y0, y1, mode, stream);
}

int kb200_int_tbar_strided(int ng, int64_t n, const double* tbar, int64_t tstride,
                           const double* D, const double* ti, const double* G, double* out,
                           int64_t ostride, int y0, int y1, int mode, void* stream) {
    return kb200_int_tbar_strided_h(ng, n, tbar, tstride, D, ti, G, out, ostride, y0, y1, mode,
                                    nullptr, nullptr, stream);
}

int kb200_int_tbar_strided_h(int ng, int64_t n, const double* tbar, int64_t tstride,
                             const double* D, const double* ti, const double* G, double* out,
                             int64_t ostride, int y0, int y1, int mode, const double* ti_h,
                             const double* G_h, void* stream) {
    // mode bit 1 (value 2): the caller guarantees G[y,x] == 0 for x > y (every quadrature of
    // kelvin/quadrature.py), which lets the kernel skip scanning the upper triangle
    const int lower = (mode & 2) ? 1 : 0;
    mode &= 1;
    if (ng <= 0 || n < 0 || y0 < 0 || y1 > ng || y0 > y1 || tstride < n || ostride < n)
        return fail(-1, "int_tbar: bad size");
    if (y0 == y1 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    IntArgs a = int_args(ng, n, tbar, tstride, D, ti, nullptr, G, out, ostride, y0, y1, lower);
    if (ti_h && G_h && y0 == 0 && y1 == ng) {
        bool done = false;
        if (ng == 10) done = launch_int_tbar_fixed<10, false>(a, mode, ti_h, nullptr, G_h, st);
        else if (ng == 16) done = launch_int_tbar_fixed<16, false>(a, mode, ti_h, nullptr, G_h, st);
        if (done) {
            KB_CHECK_LAUNCH("int_tbar_fixed_kernel");
            return 0;
        }
    }
    if (y1 - y0 <= 4)
        launch_int_tbar<4, false>(a, mode, st);
    else if (y1 - y0 <= 8)
        launch_int_tbar<8, false>(a, mode, st);
    else if (y1 - y0 <= 12)
        launch_int_tbar<12, false>(a, mode, st);
    else
        launch_int_tbar<16, false>(a, mode, st);
    KB_CHECK_LAUNCH("int_tbar_kernel");
    return 0;
}

int kb200_int_tbar_update(int ng, int64_t n, const double* tbar, int64_t tstride,
                          const double* D, const double* ti, const double* G, double* amp,
                          int64_t astride, int y0, int y1, double alpha, const double* W,
                          const double* T1x, const double* T1y, int64_t t1xs, int64_t t1ys,
                          int nvb, int noa, int nob, const double* g, double c2, double c11,
                          double* out4, double* scratch, int mode, void* stream) {
    return kb200_int_tbar_update_h(ng, n, tbar, tstride, D, ti, G, amp, astride, y0, y1, alpha, W, T1x,
                                   T1y, t1xs, t1ys, nvb, noa, nob, g, c2, c11, out4, scratch, mode,
                                   nullptr, nullptr, nullptr, stream);
}

int kb200_int_tbar_update_h(int ng, int64_t n, const double* tbar, int64_t tstride,
                            const double* D, const double* ti, const double* G, double* amp,
                            int64_t astride, int y0, int y1, double alpha, const double* W,
                            const double* T1x, const double* T1y, int64_t t1xs, int64_t t1ys,
                            int nvb, int noa, int nob, const double* g, double c2, double c11,
                            double* out4, double* scratch, int mode, const double* ti_h,
                            const double* g_h, const double* G_h, void* stream) {
    const int lower = (mode & 2) ? 1 : 0;
    mode &= 1;
    if (ng <= 0 || n <= 0 || y0 < 0 || y1 > ng || y0 >= y1 || tstride < n || astride < n)
        return fail(-1, "int_tbar_update: bad size");
    if ((T1x == nullptr) != (T1y == nullptr)) return fail(-1, "int_tbar_update: T1x/T1y");
    if (W != nullptr && g == nullptr) return fail(-1, "int_tbar_update: energy term needs g");
    if (T1x != nullptr && (n >= (1LL << 31) || nvb <= 0 || noa <= 0 || nob <= 0 ||
                           n % ((long long)nvb * noa * nob) != 0))
        return fail(-1, "int_tbar_update: bad block dims");
    cudaStream_t st = (cudaStream_t)stream;
    IntArgs a = int_args(ng, n, tbar, tstride, D, ti, g, G, amp, astride, y0, y1, lower);
    a.alpha = alpha; a.W = W; a.T1x = T1x; a.T1y = T1y; a.t1xs = t1xs; a.t1ys = t1ys;
    a.nvb = nvb; a.noa = noa; a.nob = nob; a.c2 = c2; a.c11 = c11; a.part = scratch;
    const int grid = int_grid(n);
    if (4LL * grid > 65536) return fail(-1, "int_tbar_update: scratch too small");
    bool done = false;
    if (ti_h && G_h && (g_h || W == nullptr) && y0 == 0 && y1 == ng) {
        if (ng == 10) done = launch_int_tbar_fixed<10, true>(a, mode, ti_h, g_h, G_h, st);
        else if (ng == 16) done = launch_int_tbar_fixed<16, true>(a, mode, ti_h, g_h, G_h, st);
    }
    if (done) {
    } else if (y1 - y0 <= 4)
        launch_int_tbar<4, true>(a, mode, st);
    else if (y1 - y0 <= 8)
        launch_int_tbar<8, true>(a, mode, st);
    else if (y1 - y0 <= 12)
        launch_int_tbar<12, true>(a, mode, st);
    else
        launch_int_tbar<16, true>(a, mode, st);
    KB_CHECK_LAUNCH("int_tbar_kernel(update)");
    final_sum_kernel<<<4, 32, 0, st>>>(scratch, grid, 4, out4);
    KB_CHECK_LAUNCH("final_sum_kernel");
    return 0;
}

int kb200_int_L(int ng, const int32_t dims[4], const int64_t dstride[4], const double* L,
                const double* D, const double* ti, const double* g, const double* G, double* out,
                int mode, void* stream) {
    return kb200_int_L_rows(ng, dims, dstride, L, D, ti, g, G, out, 0, ng, mode, stream);
}

int kb200_int_L_rows(int ng, const int32_t dims[4], const int64_t dstride[4], const double* L,
                     const double* D, const double* ti, const double* g, const double* G,
                     double* out, int s0, int s1, int mode, void* stream) {
    long long n = 1;
    for (int i = 0; i < 4; ++i) n *= dims[i] > 0 ? dims[i] : 0;
    return kb200_int_L_strided(ng, dims, dstride, L, n, D, ti, g, G, out, n, s0, s1, mode, stream);
}

int kb200_int_L_strided(int ng, const int32_t dims[4], const int64_t dstride[4], const double* L,
                        int64_t lstride, const double* D, const double* ti, const double* g,
                        const double* G, double* out, int64_t ostride, int s0, int s1, int mode,
                        void* stream) {
    return kb200_int_L_strided_h(ng, dims, dstride, L, lstride, D, ti, g, G, out, ostride, s0, s1, mode,
                                 nullptr, nullptr, nullptr, stream);
}

int kb200_int_L_strided_h(int ng, const int32_t dims[4], const int64_t dstride[4], const double* L,
                          int64_t lstride, const double* D, const double* ti, const double* g,
                          const double* G, double* out, int64_t ostride, int s0, int s1, int mode,
                          const double* ti_h, const double* g_h, const double* G_h, void* stream) {
    const int lower = (mode & 2) ? 1 : 0;
    mode &= 1;
    if (ng <= 0 || s0 < 0 || s1 > ng || s0 > s1) return fail(-1, "int_L: bad size");
    if (s0 == s1) return 0;
    long long n = 1;
    for (int i = 0; i < 4; ++i) {
        if (dims[i] <= 0) return fail(-1, "int_L: bad dims");
        n *= dims[i];
    }
    if (n >= (1LL << 31)) return fail(-1, "int_L: block too large");
    if (lstride < n || ostride < n) return fail(-1, "int_L: bad stride");
    cudaStream_t st = (cudaStream_t)stream;
    IntArgs a = int_args(ng, n, L, lstride, D, ti, g, G, out, ostride, s0, s1, lower);
    a.dstrided = 1;
    for (int i = 0; i < 4; ++i) {
        a.dd[i] = dims[i];
        a.ds[i] = dstride[i];
    }
    if (ti_h && g_h && G_h && s0 == 0 && s1 == ng) {
        bool done = false;
        if (ng == 10) done = launch_int_L_fixed<10>(a, mode, ti_h, g_h, G_h, st);
        else if (ng == 16) done = launch_int_L_fixed<16>(a, mode, ti_h, g_h, G_h, st);
        if (done) {
            KB_CHECK_LAUNCH("int_L_fixed_kernel");
            return 0;
        }
    }
    if (s1 - s0 <= 4)
        launch_int_L<4>(a, mode, st);
    else if (s1 - s0 <= 8)
        launch_int_L<8>(a, mode, st);
    else if (s1 - s0 <= 12)
        launch_int_L<12>(a, mode, st);
    else
        launch_int_L<16>(a, mode, st);
    KB_CHECK_LAUNCH("int_L_kernel");
    return 0;
}

int kb200_energy_pair(int ng, int nva, int nvb, int noa, int nob, const double* T2,
                      const double* T1x, const double* T1y, const double* Iabij, const double* g,
                      double c2, double c11, double* out, double* scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long n = (long long)nva * nvb * noa * nob;
    if (n <= 0 || n >= (1LL << 31)) return fail(-1, "energy_pair: bad dims");
    int grid = grid_for(n, RED_THREADS, RED_BLOCKS);
    if ((T1x == nullptr) != (T1y == nullptr)) return fail(-1, "energy_pair: T1x/T1y");
    energy_pair_kernel<<<grid, RED_THREADS, 0, st>>>(ng, nva, nvb, noa, nob, T2, T1x, T1y, Iabij, g,
                                                    c2, c11, scratch);
    KB_CHECK_LAUNCH("energy_pair_kernel");
    final_sum_kernel<<<1, 32, 0, st>>>(scratch, grid, 1, out);
    KB_CHECK_LAUNCH("final_sum_kernel");
    return 0;
}

int kb200_dot_g(int ng, int64_t n, const double* X, const double* F, const double* g, double* out,
                double* scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(n, RED_THREADS, RED_BLOCKS);
    dot_g_kernel<<<grid, RED_THREADS, 0, st>>>(ng, n, X, F, g, scratch);
    KB_CHECK_LAUNCH("dot_g_kernel");
    final_sum_kernel<<<1, 32, 0, st>>>(scratch, grid, 1, out);
    KB_CHECK_LAUNCH("final_sum_kernel");
    return 0;
}

int kb200_damp_norms(int64_t n, double* old, const double* neu, double alpha, double* out3,
                     double* scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(n, RED_THREADS, RED_BLOCKS);
    damp_norms_kernel<<<grid, RED_THREADS, 0, st>>>(n, old, neu, alpha, scratch);
    KB_CHECK_LAUNCH("damp_norms_kernel");
    final_sum_kernel<<<3, 32, 0, st>>>(scratch, grid, 3, out3);
    KB_CHECK_LAUNCH("final_sum_kernel");
    return 0;
}

int kb200_damp_norms_rows(int ng, int64_t n, double* old, int64_t ostride, const double* neu,
                          int64_t nstride, double alpha, double* out3, double* scratch,
                          void* stream) {
    if (ng <= 0 || n <= 0 || ostride < n || nstride < n) return fail(-1, "damp_norms_rows: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(n, RED_THREADS, RED_BLOCKS);
    damp_norms_rows_kernel<<<grid, RED_THREADS, 0, st>>>(ng, n, old, ostride, neu, nstride, alpha,
                                                         scratch);
    KB_CHECK_LAUNCH("damp_norms_rows_kernel");
    final_sum_kernel<<<3, 32, 0, st>>>(scratch, grid, 3, out3);
    KB_CHECK_LAUNCH("final_sum_kernel");
    return 0;
}

int kb200_dress4(const int32_t d[4], const double* eri, const double* s0, const double* s1,
                 const double* s2, const double* s3, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long n = (long long)d[0] * d[1] * d[2] * d[3];
    if (n <= 0 || n >= (1LL << 31)) return fail(-1, "dress4: bad dims");
    dress4_kernel<<<grid_for(n, 256), 256, 0, st>>>(d[0], d[1], d[2], d[3], eri, s0, s1, s2, s3, out);
    KB_CHECK_LAUNCH("dress4_kernel");
    return 0;
}

static int fill_sub4(Sub4& q, const int32_t d[4], const int64_t st[4], const int32_t* const idx[4],
                     const double* const sc[4]) {
    long long n = 1;
    for (int k = 0; k < 4; ++k) {
        if (d[k] <= 0) return -1;
        q.d[k] = d[k];
        q.st[k] = st[k];
        q.idx[k] = idx ? idx[k] : nullptr;
        q.sc[k] = sc ? sc[k] : nullptr;
        n *= d[k];
    }
    return n < (1LL << 31) ? 0 : -1;
}

int kb200_gather4(const int32_t d[4], const int64_t src_stride[4], const double* src,
                  const int32_t* const idx[4], const double* const scale[4], double* out,
                  void* stream) {
    Sub4 q;
    if (fill_sub4(q, d, src_stride, idx, scale)) return fail(-1, "gather4: bad dims");
    long long n = (long long)d[0] * d[1] * d[2] * d[3];
    gather4_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(q, src, out);
    KB_CHECK_LAUNCH("gather4_kernel");
    return 0;
}

int kb200_scatter4_add(const int32_t d[4], const int64_t dst_stride[4], const double* src,
                       const int32_t* const idx[4], const double* const scale[4], double alpha,
                       double* dst, void* stream) {
    Sub4 q;
    if (fill_sub4(q, d, dst_stride, idx, scale)) return fail(-1, "scatter4_add: bad dims");
    long long n = (long long)d[0] * d[1] * d[2] * d[3];
    scatter4_add_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(q, alpha, src, dst);
    KB_CHECK_LAUNCH("scatter4_add_kernel");
    return 0;
}

int kb200_dress2(int n0, int n1, const double* f, const double* e, const double* s0,
                 const double* s1, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long n = (long long)n0 * n1;
    if (n <= 0) return fail(-1, "dress2: bad dims");
    dress2_kernel<<<grid_for(n, 256), 256, 0, st>>>(n0, n1, f, e, s0, s1, out);
    KB_CHECK_LAUNCH("dress2_kernel");
    return 0;
}

int kb200_gsum(int ng, int64_t n, const double* X, const double* g, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return fail(-1, "gsum: bad size");
    gsum_kernel<<<grid_for(n, 256), 256, 0, st>>>(ng, n, X, g, out);
    KB_CHECK_LAUNCH("gsum_kernel");
    return 0;
}

int kb200_scale_by(int ng, int64_t n, double* X, const double* D, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return fail(-1, "scale_by: bad size");
    scale_by_kernel<<<grid_for(n * ng, 256), 256, 0, st>>>(ng, n, X, D);
    KB_CHECK_LAUNCH("scale_by_kernel");
    return 0;
}

int kb200_max_absdiff(const int32_t dims[5], const int64_t sx[5], const int64_t sy[5],
                      const double* X, const double* Y, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Box5 q;
    long long tot = 1;
    for (int k = 0; k < 5; ++k) {
        if (dims[k] <= 0) return fail(-1, "max_absdiff: bad size");
        q.d[k] = dims[k];
        q.sx[k] = sx[k];
        q.sy[k] = sy[k];
        tot *= dims[k];
    }
    if (cudaMemsetAsync(out, 0, 2 * sizeof(double), st) != cudaSuccess)
        return fail(-2, "max_absdiff: memset");
    max_absdiff_kernel<<<grid_for(tot, 256), 256, 0, st>>>(q, X, Y, (unsigned long long*)out);
    KB_CHECK_LAUNCH("max_absdiff_kernel");
    return 0;
}

}  // extern "C"
