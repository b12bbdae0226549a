// kb200_gemm.cuh -- FP64 tensor-core (DMMA) contraction kernel with gathered
// ("composite index") operands for sm_100a.
//
// C[b][cm[m]+cn[n]] = beta*C + alpha * sum_k A[b][am[m]+ak[k]] * B[b][bk[k]+bn[n]]
//
// * FP64 on sm_100a has no tcgen05/TMEM path (tcgen05.mma has no .f64 kind);
//   every mma.sync f64 shape lowers to DMMA.8x8x4 (checked with cuobjdump), so
//   the kernel issues m8n8k4 directly.
// * Operands are gathered straight from the natural abij/ijab/... storage of
//   the amplitudes and integrals through uint32 offset tables, so the index
//   transposes of the reference's einsum strings are folded into the loads.
//   (TMA cannot describe these tensors: with an odd orbital count every
//   stride is 8 mod 16 bytes, and cp.async.bulk needs 16-byte alignment --
//   so the feed is 8-byte cp.async into a 4-stage shared-memory ring.)
// * Shared-memory tiles are stored k-contiguous ([row][BK+4]) or
//   row-contiguous ([k][R+4]) -- whichever matches the operand's contiguous
//   direction in HBM, so global reads stay coalesced; both paddings make the
//   DMMA fragment loads bank-conflict free (stride == 4 mod 16 doubles).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kb200 {

constexpr int BK = 16;

struct GemmParams {
    const double* A;
    const double* B;
    double* C;
    const uint32_t* am;
    const uint32_t* ak;
    const uint32_t* bk;
    const uint32_t* bn;
    const uint32_t* cm;
    const uint32_t* cn;
    int M, N, K;
    int splitk, kchunk;          // kchunk is a multiple of BK
    long long bsA, bsB, bsC;
    double alpha, beta;
    double* partial;             // [batch][splitk][M][N] when splitk > 1
    int tilesM, tilesN;
    int batch;
    int amode, bmode;            // operand contiguity (see TileLoader); per member of a launch
    int tmaA, tmaB;              // 1: the operand is a plain matrix with 16-byte aligned rows and
                                 // arrives through its TMA tensor map (GemmGroup::tm), else gathered
};

// Up to KB200_MAX_GROUP independent contractions of one kernel configuration share a launch:
// at m = 33 a single block GEMM is 5.5 waves of CTAs, a group of four 21.9, so the idle tail of
// the last wave is paid once per group instead of once per contraction.
constexpr int MAX_GROUP = 8;
struct GemmGroup {
    // TMA tensor maps of the plain operands, [member][A, B]: 3-d tensors (contiguous index, strided
    // index, grid point).  The boxes over-fetch 4 elements of the contiguous index, which
    // reproduces the padded shared-memory pitches of TileLoader (conflict-free fragment loads)
    // with one cp.async.bulk.tensor per operand tile; out-of-range rows / k are zero-filled.
    alignas(64) CUtensorMap tm[MAX_GROUP][2];
    GemmParams p[MAX_GROUP];
    int fend[MAX_GROUP];         // running end of the full-tile CTA ranges of the members
    int rend[MAX_GROUP];         // same for the ragged-tile CTAs (these come last in the grid)
    int n;
};

// ---- mbarrier / TMA primitives (sm_90+: cp.async.bulk.tensor -> SASS UTMALDG) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// One operand tile (R rows x BK k-values) loader.  MODE 0: k-contiguous in
// HBM -> smem [R][BK+4]; MODE 1: row-contiguous -> smem [BK][R+4].
// * k-offsets of the tile after the one being loaded are fetched one k-tile
//   ahead (prefetch_k / advance_k), so table lookups never sit in front of a
//   cp.async;
// * a tile's cp.asyncs are issued in NPART slices between the DMMA groups of the
//   tile being consumed (load_part);
// * interior tiles take a FAST path (no zero-fill predicate, row pointers kept
//   in registers: one IMAD.WIDE + one LDGSTS per element) -- the DMMA warps are
//   in-order, so every non-DMMA instruction they execute is tensor-pipe idle
//   time unless the other warp of the sub-partition covers it.
__device__ __forceinline__ void cp_async8_u32(uint32_t sdst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8_u32_z(uint32_t sdst, const double* gsrc, bool valid) {
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sdst), "l"(gsrc), "r"(sz));
}

template <int R, int NT, int MODE>
struct TileLoader {
    static constexpr int LDK = BK + 4;
    static constexpr int LDR = R + 4;
    static constexpr int STAGE = (MODE == 0) ? R * LDK : BK * LDR;
    static constexpr int NPART = BK / 4;
    // MODE 0 mapping
    static constexpr int RSTEP = NT / BK;            // rows covered per pass
    static constexpr int RPT = (R + RSTEP - 1) / RSTEP;
    // MODE 1 mapping
    static constexpr int KSTEP = (NT / R) > 0 ? (NT / R) : 1;
    static constexpr int KPT = BK / KSTEP;
    static_assert(MODE == 0 || (NT % R == 0 && BK % KSTEP == 0), "bad tile/threads");
    static constexpr int CNT = (MODE == 0) ? RPT : KPT;       // cp.asyncs per thread per tile
    static constexpr int PER = (CNT + NPART - 1) / NPART;     // per slice
    static constexpr int NKO = (MODE == 0) ? 1 : KPT;
    static constexpr int NRP = (MODE == 0) ? RPT : 1;

    const uint32_t* ktab;
    const double* rptr[NRP];  // row base pointers (invalid rows point at the operand base)
    uint32_t koff[NKO];       // k-offsets of the tile being loaded
    uint32_t koff2[NKO];      // k-offsets of the tile after it (in flight)
    uint32_t soff;            // this thread's byte offset inside a stage
    unsigned rvalid;
    unsigned kvalid, kvalid2;
    bool rows_full;           // every row of this tile is inside the operand (CTA-uniform)

    __device__ __forceinline__ void init(const double* base, const uint32_t* rtab,
                                         const uint32_t* ktab_, int row0, int nrows, int tid) {
        ktab = ktab_;
        rvalid = 0;
        kvalid = kvalid2 = 0;
        rows_full = (row0 + R <= nrows);
        if (MODE == 0) {
            int r0 = tid / BK;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                int r = r0 + i * RSTEP;
                int row = row0 + r;
                bool v = (r < R) && (row < nrows);
                rptr[i] = base + (v ? rtab[row] : 0u);
                rvalid |= (v ? 1u : 0u) << i;
            }
            soff = (uint32_t)((r0 * LDK + tid % BK) * 8);
        } else {
            int r = tid % R;
            int row = row0 + r;
            bool v = row < nrows;
            rptr[0] = base + (v ? rtab[row] : 0u);
            rvalid = v ? 1u : 0u;
            soff = (uint32_t)(((tid / R) * LDR + r) * 8);
        }
    }

    // start fetching the k-offsets of the tile starting at k0 (into koff2)
    __device__ __forceinline__ void prefetch_k(int k0, int kend, int tid) {
        kvalid2 = 0;
        if (MODE == 0) {
            int k = k0 + tid % BK;
            bool kv = k < kend;
            koff2[0] = kv ? __ldg(ktab + k) : 0u;
            kvalid2 = kv ? 1u : 0u;
        } else {
            int kk0 = tid / R;
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                int k = k0 + kk0 + i * KSTEP;
                bool kv = k < kend;
                koff2[i] = kv ? __ldg(ktab + k) : 0u;
                kvalid2 |= (kv ? 1u : 0u) << i;
            }
        }
    }
    // make the prefetched offsets current
    __device__ __forceinline__ void advance_k() {
#pragma unroll
        for (int i = 0; i < NKO; ++i) koff[i] = koff2[i];
        kvalid = kvalid2;
    }

    // issue slice `part` (0..NPART-1) of the tile whose k-offsets are current.
    // sstage: shared-space byte address of the destination stage.
    template <bool FAST>
    __device__ __forceinline__ void load_part(uint32_t sstage, int part) const {
        const uint32_t sb = sstage + soff;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = part * PER + j;
                if (i < RPT && (i * RSTEP < R)) {
                    const uint32_t sd = sb + (uint32_t)(i * RSTEP * LDK * 8);
                    if (FAST) {
                        cp_async8_u32(sd, rptr[i] + koff[0]);
                    } else {
                        bool v = (kvalid & 1u) && ((rvalid >> i) & 1u);
                        cp_async8_u32_z(sd, rptr[i] + koff[0], v);
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = part * PER + j;
                if (i < KPT) {
                    const uint32_t sd = sb + (uint32_t)(i * KSTEP * LDR * 8);
                    if (FAST) {
                        cp_async8_u32(sd, rptr[0] + koff[i]);
                    } else {
                        bool v = ((kvalid >> i) & 1u) && (rvalid & 1u);
                        cp_async8_u32_z(sd, rptr[0] + koff[i], v);
                    }
                }
            }
        }
    }

    __device__ __forceinline__ void load_all(uint32_t sstage) const {
#pragma unroll
        for (int part = 0; part < NPART; ++part) load_part<false>(sstage, part);
    }

    // fragment element (row r, k index kk) of a stage
    __device__ __forceinline__ static double frag(const double* stage, int r, int kk) {
        return (MODE == 0) ? stage[r * LDK + kk] : stage[kk * LDR + r];
    }
};

// shared-memory doubles one stage of an R-row operand tile needs in either contiguity mode
template <int R, int NT>
struct StageMax {
    static constexpr int value = TileLoader<R, NT, 0>::STAGE > TileLoader<R, NT, 1>::STAGE
                                     ? TileLoader<R, NT, 0>::STAGE
                                     : TileLoader<R, NT, 1>::STAGE;
};

// The tile loop for one member and one (AMODE, BMODE) pair; the kernel below picks the pair of
// the member its CTA works on at run time (CTA-uniform), so contractions with different operand
// contiguities share a launch: at small tau batches (tau-sharded runs) a launch of one or two
// block GEMMs is a fraction of a wave of CTAs.
template <int WARPS_M, int WARPS_N, int WM, int WN, int AMODE, int BMODE, int STAGES, bool ILV,
          bool TMA>
__device__ __forceinline__ void gemm_tab_body(const GemmParams& p, const CUtensorMap* tmA,
                                              const CUtensorMap* tmB, int f_, bool in_full,
                                              double* smem) {
    constexpr int BM = WARPS_M * WM;
    constexpr int BN = WARPS_N * WN;
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int MI = WM / 8;
    constexpr int NI = WN / 8;
    using LA = TileLoader<BM, NT, AMODE>;
    using LB = TileLoader<BN, NT, BMODE>;

    double* As = smem;
    double* Bs = smem + STAGES * StageMax<BM, NT>::value;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wmi = warp % WARPS_M;
    const int wni = warp / WARPS_M;
    const int g = lane >> 2;
    const int t = lane & 3;
    int b, tm, tn;
    {
        const int rm = (p.M % BM) ? 1 : 0, rn = (p.N % BN) ? 1 : 0;
        const int fm = p.tilesM - rm, fn = p.tilesN - rn;
        const int nf = fm * fn, nr = p.tilesM * p.tilesN - nf;
        if (in_full) {
            b = f_ / nf;
            const int t_ = f_ - b * nf;
            tm = t_ % fm;
            tn = t_ / fm;
        } else {
            const int g_ = f_;
            b = g_ / nr;
            const int r_ = g_ - b * nr;
            if (rm && r_ < p.tilesN) {
                tm = p.tilesM - 1;
                tn = r_;
            } else {
                tn = p.tilesN - 1;
                tm = r_ - (rm ? p.tilesN : 0);
            }
        }
    }
    const int ks = blockIdx.y;
    const int m0 = tm * BM;
    const int n0 = tn * BN;
    const int kbeg = ks * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    const int nk = (kend - kbeg + BK - 1) / BK;

    // validity of this warp's row / column groups (warp-uniform)
    unsigned mval = 0, nval = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
        if (m0 + (i * WARPS_M + wmi) * 8 < p.M) mval |= 1u << i;
#pragma unroll
    for (int j = 0; j < NI; ++j)
        if (n0 + (j * WARPS_N + wni) * 8 < p.N) nval |= 1u << j;

    const int mcnt = __popc(mval), ncnt = __popc(nval);   // valid groups form a prefix

    // operand feed: TMA (one elected thread issues cp.async.bulk.tensor per operand tile, the
    // stage's mbarrier counts the bytes) for plain operands, gathered 8-byte cp.async otherwise
    // (compile-time per body: both operands by TMA, or both gathered -- run-time tests inside
    // the tile loop keep the compiler from interleaving the copies with the DMMA groups)
    constexpr bool ta = TMA, tb = TMA;
    constexpr bool tany = TMA;
    LA la;
    LB lb;
    la.rows_full = true;
    lb.rows_full = true;
    if (!ta) la.init(p.A + (long long)b * p.bsA, p.am, p.ak, m0, p.M, tid);
    if (!tb) lb.init(p.B + (long long)b * p.bsB, p.bn, p.bk, n0, p.N, tid);
    const uint32_t As_u = (uint32_t)__cvta_generic_to_shared(As);
    const uint32_t Bs_u = (uint32_t)__cvta_generic_to_shared(Bs);
    const bool rows_full = la.rows_full && lb.rows_full;
    const uint32_t bar_u = (uint32_t)__cvta_generic_to_shared(
        smem + STAGES * (StageMax<BM, NT>::value + StageMax<BN, NT>::value));
    // full[s]: the stage's bytes have landed (1 arrival + transaction count);
    // empty[s]: every warp is done reading the stage (NT / 32 arrivals)
    const uint32_t empty_u = bar_u + 8 * STAGES;
    if (tany) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(bar_u + 8 * s, 1);
                mbar_init(empty_u + 8 * s, NT / 32);
            }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
    }
    const uint32_t tx_bytes = (ta ? LA::STAGE * 8 : 0) + (tb ? LB::STAGE * 8 : 0);
    const int bA = p.bsA ? b : 0, bB = p.bsB ? b : 0;       // grid-point coordinate (0: shared)
    auto tma_issue = [&](int s, int k0) {
        const uint32_t bar = bar_u + 8 * s;
        mbar_expect_tx(bar, tx_bytes);
        if (ta) {
            if (AMODE == 0)
                tma_load_3d(As_u + s * LA::STAGE * 8, tmA, bar, k0, m0, bA);
            else
                tma_load_3d(As_u + s * LA::STAGE * 8, tmA, bar, m0, k0, bA);
        }
        if (tb) {
            if (BMODE == 0)
                tma_load_3d(Bs_u + s * LB::STAGE * 8, tmB, bar, k0, n0, bB);
            else
                tma_load_3d(Bs_u + s * LB::STAGE * 8, tmB, bar, n0, k0, bB);
        }
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (!ta) la.prefetch_k(kbeg, kend, tid);
    if (!tb) lb.prefetch_k(kbeg, kend, tid);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (!ta) {
            la.advance_k();
            la.prefetch_k(kbeg + (s + 1) * BK, kend, tid);
        }
        if (!tb) {
            lb.advance_k();
            lb.prefetch_k(kbeg + (s + 1) * BK, kend, tid);
        }
        if (s < nk) {
            if (!ta) la.load_all(As_u + s * LA::STAGE * 8);
            if (!tb) lb.load_all(Bs_u + s * LB::STAGE * 8);
            if (tany && tid == 0) tma_issue(s, kbeg + s * BK);
        }
        cp_async_commit();
    }

    const bool full = (mval == (1u << MI) - 1u) && (nval == (1u << NI) - 1u);
    if (TMA) {
        // TMA tile loop: NO block barrier.  Every warp waits for the stage's bytes (full
        // barrier), multiplies, and releases the stage (empty barrier); one thread refills the
        // stage released in the previous iteration once all warps have left it.  The warps may
        // drift apart by up to STAGES - 1 k-tiles.
        for (int kt = 0; kt < nk; ++kt) {
            const int st_ = kt % STAGES;
            mbar_wait(bar_u + 8 * st_, (kt / STAGES) & 1);
            const int nxt = kt + STAGES - 1;
            if (tid == 0 && nxt < nk) {
                if (kt >= 1) mbar_wait(empty_u + 8 * ((kt - 1) % STAGES), ((kt - 1) / STAGES) & 1);
                tma_issue(nxt % STAGES, kbeg + nxt * BK);
            }
            const double* as = As + st_ * LA::STAGE;
            const double* bs = Bs + st_ * LB::STAGE;
            if (full) {
#pragma unroll
                for (int k4 = 0; k4 < BK / 4; ++k4) {
                    double af[MI], bf[NI];
#pragma unroll
                    for (int i = 0; i < MI; ++i)
                        af[i] = LA::frag(as, (i * WARPS_M + wmi) * 8 + g, k4 * 4 + t);
#pragma unroll
                    for (int j = 0; j < NI; ++j)
                        bf[j] = LB::frag(bs, (j * WARPS_N + wni) * 8 + g, k4 * 4 + t);
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NI; ++j)
                            dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            } else {
#pragma unroll
                for (int k4 = 0; k4 < BK / 4; ++k4) {
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        if (i < mcnt) {
                            double a = LA::frag(as, (i * WARPS_M + wmi) * 8 + g, k4 * 4 + t);
#pragma unroll
                            for (int j = 0; j < NI; ++j) {
                                if (j < ncnt) {
                                    double bb = LB::frag(bs, (j * WARPS_N + wni) * 8 + g, k4 * 4 + t);
                                    dmma884(acc[i][j][0], acc[i][j][1], a, bb);
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_u + 8 * st_);
        }
    } else
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
#ifndef KB200_EXP_NOSYNC
        __syncthreads();
#endif
        const int nxt = kt + STAGES - 1;
#ifdef KB200_EXP_NOLOAD
        const bool do_load = false;
#else
        const bool do_load = nxt < nk;
#endif
        if (!ta) {
            la.advance_k();
            if (nxt + 1 < nk) la.prefetch_k(kbeg + (nxt + 1) * BK, kend, tid);
        }
        if (!tb) {
            lb.advance_k();
            if (nxt + 1 < nk) lb.prefetch_k(kbeg + (nxt + 1) * BK, kend, tid);
        }
        if (tany && do_load && tid == 0) tma_issue(nxt % STAGES, kbeg + nxt * BK);
        const uint32_t an = As_u + (nxt % STAGES) * LA::STAGE * 8;
        const uint32_t bn = Bs_u + (nxt % STAGES) * LB::STAGE * 8;
        const double* as = As + (kt % STAGES) * LA::STAGE;
        const double* bs = Bs + (kt % STAGES) * LB::STAGE;
        // interior tile and interior k-tile: no predicates anywhere in the loop body
        const bool fast = rows_full && (kbeg + (nxt + 1) * BK <= kend);
        if (full && fast && do_load) {
#pragma unroll
            for (int k4 = 0; k4 < BK / 4; ++k4) {
                double af[MI], bf[NI];
#pragma unroll
                for (int i = 0; i < MI; ++i)
                    af[i] = LA::frag(as, (i * WARPS_M + wmi) * 8 + g, k4 * 4 + t);
#pragma unroll
                for (int j = 0; j < NI; ++j)
                    bf[j] = LB::frag(bs, (j * WARPS_N + wni) * 8 + g, k4 * 4 + t);
                if (ILV) {
                    if (!ta) la.template load_part<true>(an, k4);
                    if (!tb) lb.template load_part<true>(bn, k4);
                } else if (k4 == 0) {
#pragma unroll
                    for (int q = 0; q < BK / 4; ++q) {
                        if (!ta) la.template load_part<true>(an, q);
                        if (!tb) lb.template load_part<true>(bn, q);
                    }
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        } else {
            if (do_load) {
                if (!ta) la.load_all(an);
                if (!tb) lb.load_all(bn);
            }
            if (full) {
#pragma unroll
                for (int k4 = 0; k4 < BK / 4; ++k4) {
                    double af[MI], bf[NI];
#pragma unroll
                    for (int i = 0; i < MI; ++i)
                        af[i] = LA::frag(as, (i * WARPS_M + wmi) * 8 + g, k4 * 4 + t);
#pragma unroll
                    for (int j = 0; j < NI; ++j)
                        bf[j] = LB::frag(bs, (j * WARPS_N + wni) * 8 + g, k4 * 4 + t);
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NI; ++j)
                            dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            } else {
#pragma unroll
                for (int k4 = 0; k4 < BK / 4; ++k4) {
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        if (i < mcnt) {
                            double a = LA::frag(as, (i * WARPS_M + wmi) * 8 + g, k4 * 4 + t);
#pragma unroll
                            for (int j = 0; j < NI; ++j) {
                                if (j < ncnt) {
                                    double bb = LB::frag(bs, (j * WARPS_N + wni) * 8 + g, k4 * 4 + t);
                                    dmma884(acc[i][j][0], acc[i][j][1], a, bb);
                                }
                            }
                        }
                    }
                }
            }
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row = g][col = 2t, 2t+1] of each 8x8 tile
    if (p.splitk == 1) {
        double* C = p.C + (long long)b * p.bsC;
        uint32_t co[NI][2];
        unsigned cval = 0;
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                int col = n0 + (j * WARPS_N + wni) * 8 + 2 * t + c;
                bool v = col < p.N;
                co[j][c] = v ? __ldg(p.cn + col) : 0u;
                cval |= (v ? 1u : 0u) << (2 * j + c);
            }
        uint32_t ro[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int row = m0 + (i * WARPS_M + wmi) * 8 + g;
            ro[i] = (row < p.M) ? __ldg(p.cm + row) : 0xffffffffu;
        }
        const bool rmw = p.beta != 0.0;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int row = m0 + (i * WARPS_M + wmi) * 8 + g;
            if (row >= p.M) continue;
            double* Cr = C + ro[i];
            double old[NI][2];
            if (rmw) {
#pragma unroll
                for (int j = 0; j < NI; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        old[j][c] = ((cval >> (2 * j + c)) & 1u) ? Cr[co[j][c]] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if ((cval >> (2 * j + c)) & 1u) {
                        double v = p.alpha * acc[i][j][c];
                        if (rmw) v += p.beta * old[j][c];
                        Cr[co[j][c]] = v;
                    }
        }
    } else {
        double* P = p.partial + ((size_t)b * p.splitk + ks) * (size_t)p.M * p.N;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int row = m0 + (i * WARPS_M + wmi) * 8 + g;
            if (row >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NI; ++j) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    int col = n0 + (j * WARPS_N + wni) * 8 + 2 * t + c;
                    if (col < p.N) P[(size_t)row * p.N + col] = acc[i][j][c];
                }
            }
        }
    }
}

// Row groups (8 rows) are dealt to the warps round-robin, so that in a ragged
// edge tile every warp loses the same share of work and whole invalid groups
// are skipped: edge tiles cost in proportion to their valid area.
template <int WARPS_M, int WARPS_N, int WM, int WN, int STAGES, bool ILV, int MINB>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB)
    gemm_tab_kernel(const __grid_constant__ GemmGroup grp) {
    extern __shared__ __align__(128) double smem[];
    // CTA order: member by member, batch-major over the full tiles (CTAs that run together
    // share one batch's operands in L2), then the ragged -- cheaper -- edge tiles of all
    // members and batches, so the final partial wave is filled with the short CTAs.
    int mi = 0;
    int f_ = blockIdx.x;
    bool in_full;
    {
        const int total_full = grp.fend[grp.n - 1];
        in_full = f_ < total_full;
        if (in_full) {
            while (f_ >= grp.fend[mi]) ++mi;
            if (mi) f_ -= grp.fend[mi - 1];
        } else {
            f_ -= total_full;
            while (f_ >= grp.rend[mi]) ++mi;
            if (mi) f_ -= grp.rend[mi - 1];
        }
    }
    const GemmParams& p = grp.p[mi];
    const bool tma = p.tmaA && p.tmaB;
#define GEMM_TAB_BODY(AM, BM_)                                                                      \
    do {                                                                                            \
        if (tma)                                                                                    \
            gemm_tab_body<WARPS_M, WARPS_N, WM, WN, AM, BM_, STAGES, ILV, true>(                    \
                p, &grp.tm[mi][0], &grp.tm[mi][1], f_, in_full, smem);                              \
        else                                                                                        \
            gemm_tab_body<WARPS_M, WARPS_N, WM, WN, AM, BM_, STAGES, ILV, false>(                   \
                p, &grp.tm[mi][0], &grp.tm[mi][1], f_, in_full, smem);                              \
    } while (0)
    if (p.amode == 0) {
        if (p.bmode == 0)
            GEMM_TAB_BODY(0, 0);
        else
            GEMM_TAB_BODY(0, 1);
    } else {
        if (p.bmode == 0)
            GEMM_TAB_BODY(1, 0);
        else
            GEMM_TAB_BODY(1, 1);
    }
#undef GEMM_TAB_BODY
}

// ---------------------------------------------------------------------------
// Long-K contraction into a tiny output (tile id 6):
//   C[m, n] (+)= alpha * sum_k A[m, k] * B[k, n],   M, N <= 8*T (T = 3 or 5), K ~ n^3.
// These are the F_vv / F_oo builds, the singles residual and their adjoints: one operand is an
// n^4 tensor streamed once from HBM, the output is n^2.  The 128x128 / 64x64 tiles waste
// 3/4 of their DMMAs here; this kernel holds the whole (padded) output in the accumulators
// of every warp and splits the K range instead: grid = (splitk, batch), each CTA walks its
// K chunk in stages of 32, warp w owns the k4-step w of every stage (25 DMMAs for 10
// fragment loads), the eight partial outputs are tree-reduced through shared memory and
// the CTA result goes to the split-K workspace (deterministic reduction as for tile 0..5).
// ---------------------------------------------------------------------------
constexpr int LK_BK = 32;
constexpr int LK_STAGES = 4;

template <int R, int MODE, int NT>
struct LongKLoader {
    static constexpr int LDK = LK_BK + 4;
    static constexpr int LDR = R + 4;
    static constexpr int STAGE = (MODE == 0) ? R * LDK : LK_BK * LDR;
    static constexpr int ITERS = (R * LK_BK) / NT;
    static constexpr int NKV = (MODE == 0) ? 1 : ITERS;   // distinct k offsets a thread needs per stage
    static_assert((R * LK_BK) % NT == 0 && NT % LK_BK == 0, "tile must divide the CTA");
    // Element (row, kk) of a stage <- base[rowtab[row] + ktab[k0 + kk]].  A thread's (row, kk)
    // assignments are the same in every stage; its k offsets for the next stage are fetched one
    // stage ahead (prefetch/advance) so that no cp.async waits on a table load.
    const double* base;
    const uint32_t* ktab;
    uint32_t roff[ITERS];
    int kk[NKV];
    int dst[ITERS];
    unsigned rmask;
    uint32_t kcur[NKV], knxt[NKV];

    __device__ __forceinline__ void init(const double* base_, const uint32_t* __restrict__ rowtab,
                                         int nrows, const uint32_t* ktab_, int tid) {
        base = base_;
        ktab = ktab_;
        rmask = 0;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int e = it * NT + tid;
            const int row = (MODE == 0) ? e / LK_BK : e % R;
            const int k_ = (MODE == 0) ? e % LK_BK : e / R;
            if (MODE == 0) {
                if (it == 0) kk[0] = k_;
            } else {
                kk[it] = k_;
            }
            dst[it] = (MODE == 0) ? row * LDK + k_ : k_ * LDR + row;
            const bool v = row < nrows;
            roff[it] = v ? __ldg(rowtab + row) : 0u;
            rmask |= (v ? 1u : 0u) << it;
        }
#pragma unroll
        for (int v = 0; v < NKV; ++v) kcur[v] = knxt[v] = 0xffffffffu;
    }
    __device__ __forceinline__ void prefetch(int k0, int kend) {
#pragma unroll
        for (int v = 0; v < NKV; ++v) {
            const int k = k0 + kk[v];
            knxt[v] = (k < kend) ? __ldg(ktab + k) : 0xffffffffu;
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int v = 0; v < NKV; ++v) kcur[v] = knxt[v];
    }
    __device__ __forceinline__ void issue(uint32_t stage_u) const {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t ko = kcur[(MODE == 0) ? 0 : it];
            const bool valid = ((rmask >> it) & 1u) && (ko != 0xffffffffu);
            const double* src = valid ? base + (size_t)roff[it] + ko : base;
            unsigned sa = stage_u + dst[it] * 8;
            int sz = valid ? 8 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(src), "r"(sz));
        }
    }
    __device__ __forceinline__ static double frag(const double* stage, int r, int k_) {
        return (MODE == 0) ? stage[r * LDK + k_] : stage[k_ * LDR + r];
    }
};

template <int T, int AMODE, int BMODE, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) longk_kernel(const GemmParams p) {
    constexpr int R = 8 * T;
    constexpr int NT = NW * 32;
    constexpr int KSTEPS = (LK_BK / 4) / NW;      // k4-steps of a stage owned by one warp
    static_assert((LK_BK / 4) % NW == 0, "warps must divide the k4-steps of a stage");
    using LA = LongKLoader<R, AMODE, NT>;
    using LB = LongKLoader<R, BMODE, NT>;
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + LK_STAGES * LA::STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int ks = blockIdx.x;
    const int b = blockIdx.y;
    const int kbeg = ks * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    const int nk = (kend - kbeg + LK_BK - 1) / LK_BK;

    LA la;
    LB lb;
    la.init(p.A + (long long)b * p.bsA, p.am, p.M, p.ak, tid);
    lb.init(p.B + (long long)b * p.bsB, p.bn, p.N, p.bk, tid);
    const uint32_t As_u = (uint32_t)__cvta_generic_to_shared(As);
    const uint32_t Bs_u = (uint32_t)__cvta_generic_to_shared(Bs);

    double acc[T][T][2];
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    la.prefetch(kbeg, kend);
    lb.prefetch(kbeg, kend);
#pragma unroll
    for (int s = 0; s < LK_STAGES - 1; ++s) {
        la.advance();
        lb.advance();
        la.prefetch(kbeg + (s + 1) * LK_BK, kend);
        lb.prefetch(kbeg + (s + 1) * LK_BK, kend);
        if (s < nk) {
            la.issue(As_u + s * LA::STAGE * 8);
            lb.issue(Bs_u + s * LB::STAGE * 8);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<LK_STAGES - 2>();
        __syncthreads();
        const int nxt = kt + LK_STAGES - 1;
        la.advance();
        lb.advance();
        if (nxt + 1 < nk) {
            la.prefetch(kbeg + (nxt + 1) * LK_BK, kend);
            lb.prefetch(kbeg + (nxt + 1) * LK_BK, kend);
        }
        if (nxt < nk) {
            la.issue(As_u + (nxt % LK_STAGES) * LA::STAGE * 8);
            lb.issue(Bs_u + (nxt % LK_STAGES) * LB::STAGE * 8);
        }
        cp_async_commit();
        const double* as = As + (kt % LK_STAGES) * LA::STAGE;
        const double* bs = Bs + (kt % LK_STAGES) * LB::STAGE;
#pragma unroll
        for (int q = 0; q < KSTEPS; ++q) {
            const int k4 = q * NW + warp;
            double af[T], bf[T];
#pragma unroll
            for (int i = 0; i < T; ++i) af[i] = LA::frag(as, i * 8 + g, k4 * 4 + t);
#pragma unroll
            for (int j = 0; j < T; ++j) bf[j] = LB::frag(bs, j * 8 + g, k4 * 4 + t);
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // tree reduction of the per-warp partial outputs (deterministic order)
    double2* red = reinterpret_cast<double2*>(smem);
#pragma unroll
    for (int half = NW / 2; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j)
                    red[((warp - half) * T * T + i * T + j) * 32 + lane] =
                        make_double2(acc[i][j][0], acc[i][j][1]);
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) {
                    const double2 v = red[(warp * T * T + i * T + j) * 32 + lane];
                    acc[i][j][0] += v.x;
                    acc[i][j][1] += v.y;
                }
        }
        __syncthreads();
    }
    if (warp != 0) return;
    if (p.splitk == 1) {
        double* C = p.C + (long long)b * p.bsC;
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const int row = i * 8 + g;
            if (row >= p.M) continue;
            const uint32_t ro = p.cm[row];
#pragma unroll
            for (int j = 0; j < T; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = j * 8 + 2 * t + c;
                    if (col < p.N) {
                        double* dst = C + ro + p.cn[col];
                        double v = p.alpha * acc[i][j][c];
                        if (p.beta != 0.0) v += p.beta * (*dst);
                        *dst = v;
                    }
                }
        }
    } else {
        double* P = p.partial + ((size_t)b * p.splitk + ks) * (size_t)p.M * p.N;
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const int row = i * 8 + g;
            if (row >= p.M) continue;
#pragma unroll
            for (int j = 0; j < T; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = j * 8 + 2 * t + c;
                    if (col < p.N) P[(size_t)row * p.N + col] = acc[i][j][c];
                }
        }
    }
}

// ---------------------------------------------------------------------------
// Skinny streaming contraction (tile id 7): the n^5 "dressing" terms
//   C[m, n] (+)= alpha * sum_k A[m, k] * B[k, n],   M ~ n^3 rows,  K <= 4*KT,  N <= 8*NI (<= 40).
// 24 bytes of HBM traffic per C element for ~2K flops: bandwidth-bound, every A and C element is
// touched exactly once, so nothing is staged except B (K x N, a few KB, once per CTA).
// One warp owns 8 rows at a time: its A fragments come straight from global memory into the
// DMMA operand registers (4 consecutive k per row = one 32-byte sector), B fragments from shared
// memory, C is read-modify-written in the accumulator layout.  No barriers in the row loop;
// latency is hidden by 16 resident warps per SM.
// ---------------------------------------------------------------------------
constexpr int SK_STAGES = 4;
constexpr int SK_WARPS = 8;

template <int KT, int NI>
struct SkinnyCfg {
    static constexpr int KPAD = KT * 4;
    // pitches with conflict-free fragment reads (see LongKLoader / TileLoader)
    static constexpr int AP = (KPAD % 16 == 4 || KPAD % 16 == 12) ? KPAD : KPAD + 4;
    static constexpr int CP = NI * 8;
    static constexpr int STAGE = 8 * AP + 8 * CP + 8;          // doubles: A tile, C tile, 8 row offsets (+pad)
    static constexpr int BP = AP;
    static constexpr int SMEM = (NI * 8 * BP + SK_WARPS * SK_STAGES * STAGE) * 8;
};

template <int KT, int NI>
__global__ void __launch_bounds__(SK_WARPS * 32, 1) skinny_kernel(const GemmParams p, int a_mode, int c_mode) {
    using Cfg = SkinnyCfg<KT, NI>;
    constexpr int KPAD = Cfg::KPAD, AP = Cfg::AP, CP = Cfg::CP, BP = Cfg::BP;
    constexpr int AIT = (8 * KPAD) / 32;       // cp.async per lane for the A tile (= KT)
    constexpr int CIT = (8 * CP) / 32;         // ... for the C tile (= 2 NI)
    extern __shared__ double smem[];
    double* Bs = smem;                                         // [NI*8][BP], k contiguous
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int b = blockIdx.y;
    const double* __restrict__ A = p.A + (long long)b * p.bsA;
    const double* __restrict__ B = p.B + (long long)b * p.bsB;
    double* __restrict__ C = p.C + (long long)b * p.bsC;
    double* wst = smem + NI * 8 * BP + warp * SK_STAGES * Cfg::STAGE;   // this warp's stages

    for (int idx = tid; idx < NI * 8 * KPAD; idx += SK_WARPS * 32) {
        const int n = idx / KPAD, k = idx - n * KPAD;
        Bs[n * BP + k] = (n < p.N && k < p.K) ? B[(size_t)p.bk[k] + p.bn[n]] : 0.0;
    }
    // per-lane constants of the copy pattern (the same in every tile): source offset along k / n,
    // destination slot, source lane of the row offset
    uint32_t a_ko[AIT];
    int a_dst[AIT];
    int a_row[AIT];
#pragma unroll
    for (int i = 0; i < AIT; ++i) {
        const int e = i * 32 + lane;
        const int row = a_mode ? (e & 7) : e / KPAD;
        const int k = a_mode ? (e >> 3) : e - row * KPAD;
        a_ko[i] = (k < p.K) ? __ldg(p.ak + k) : 0xffffffffu;
        a_dst[i] = (row * AP + k) * 8;
        a_row[i] = row;
    }
    // accumulator-layout column offsets (epilogue) and C-tile copy pattern
    uint32_t cno[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = j * 8 + 2 * t + c;
            cno[j][c] = (col < p.N) ? __ldg(p.cn + col) : 0xffffffffu;
        }
    __syncthreads();
    // B fragments stay in registers for the whole kernel
    double bf[KT][NI];
#pragma unroll
    for (int k4 = 0; k4 < KT; ++k4)
#pragma unroll
        for (int j = 0; j < NI; ++j) bf[k4][j] = Bs[(j * 8 + g) * BP + k4 * 4 + t];

    const bool rmw = p.beta != 0.0;
    const int ntiles = (p.M + 7) >> 3;
    const int stride = gridDim.x * SK_WARPS;
    const int first = blockIdx.x * SK_WARPS + warp;

    // row offsets of a tile: lanes 0-7 hold am[row], lanes 8-15 cm[row] (fetched one tile ahead)
    auto fetch_rows = [&](int mt) -> uint32_t {
        const int row = mt * 8 + (lane & 7);
        if (mt >= ntiles || row >= p.M || lane >= 16) return 0xffffffffu;
        return __ldg((lane < 8 ? p.am : p.cm) + row);
    };
    auto issue = [&](int mt, int s, uint32_t rows) {
        if (mt >= ntiles) return;
        double* st = wst + s * Cfg::STAGE;
        const uint32_t st_u = (uint32_t)__cvta_generic_to_shared(st);
        uint32_t* roffs = reinterpret_cast<uint32_t*>(st + 8 * AP + 8 * CP);
        if (lane >= 8 && lane < 16) roffs[lane - 8] = rows;           // C row offsets for the epilogue
#pragma unroll
        for (int i = 0; i < AIT; ++i) {
            const uint32_t ro = __shfl_sync(0xffffffffu, rows, a_row[i]);
            const bool v = (ro != 0xffffffffu) && (a_ko[i] != 0xffffffffu);
            const double* src = v ? A + (size_t)ro + a_ko[i] : A;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(st_u + a_dst[i]), "l"(src),
                         "r"(v ? 8 : 0));
        }
        if (rmw) {
            // old C values in the accumulator layout: lane (g, t) fetches its own 2 NI elements
            const uint32_t ro = __shfl_sync(0xffffffffu, rows, 8 + g);
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const bool v = (ro != 0xffffffffu) && (cno[j][c] != 0xffffffffu);
                    const double* src = v ? C + (size_t)ro + cno[j][c] : C;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(
                                     st_u + (8 * AP + g * CP + j * 8 + 2 * t + c) * 8),
                                 "l"(src), "r"(v ? 8 : 0));
                }
        }
    };
    (void)c_mode;
    (void)CIT;

    uint32_t rows_next = fetch_rows(first);
#pragma unroll
    for (int s = 0; s < SK_STAGES - 1; ++s) {
        const uint32_t rows = rows_next;
        rows_next = fetch_rows(first + (s + 1) * stride);
        issue(first + s * stride, s, rows);
        cp_async_commit();
    }
    int it = 0;
    for (int mt = first; mt < ntiles; mt += stride, ++it) {
        cp_async_wait<SK_STAGES - 2>();
        __syncwarp();
        {
            const uint32_t rows = rows_next;
            rows_next = fetch_rows(mt + SK_STAGES * stride);
            issue(mt + (SK_STAGES - 1) * stride, (it + SK_STAGES - 1) % SK_STAGES, rows);
            cp_async_commit();
        }
        const double* st = wst + (it % SK_STAGES) * Cfg::STAGE;
        const double* as = st + g * AP + t;
        const double* cs = st + 8 * AP + g * CP + 2 * t;
        const uint32_t* roffs = reinterpret_cast<const uint32_t*>(st + 8 * AP + 8 * CP);
        double acc[NI][2];
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < KT; ++k4) {
            const double af = as[k4 * 4];
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc[j][0], acc[j][1], af, bf[k4][j]);
        }
        const uint32_t ro = roffs[g];
        if (ro != 0xffffffffu) {
            double* Cr = C + ro;
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (cno[j][c] != 0xffffffffu) {
                        double v = p.alpha * acc[j][c];
                        if (rmw) v += p.beta * cs[j * 8 + c];
                        Cr[cno[j][c]] = v;
                    }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

// deterministic split-K reduction + scatter
__global__ void splitk_reduce_kernel(const GemmParams p, int batch) {
    size_t mn = (size_t)p.M * p.N;
    size_t total = mn * batch;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(idx / mn);
        size_t r = idx - (size_t)b * mn;
        int row = (int)(r / p.N);
        int col = (int)(r - (size_t)row * p.N);
        const double* P = p.partial + (size_t)b * p.splitk * mn + r;
        double s = 0.0;
        for (int ks = 0; ks < p.splitk; ++ks) s += P[(size_t)ks * mn];
        double* dst = p.C + (long long)b * p.bsC + p.cm[row] + p.cn[col];
        double v = p.alpha * s;
        if (p.beta != 0.0) v += p.beta * (*dst);
        *dst = v;
    }
}

// the same for long split-K chains over a tiny output (the K = m^3 contractions into m x m
// blocks, split-K of a few hundred): 32 consecutive outputs x 16 slices of the chain per CTA,
// fixed summation order (slice-strided partial sums, then the 16 slices in order)
__global__ void __launch_bounds__(512) splitk_reduce_wide_kernel(const GemmParams p, int batch) {
    __shared__ double sh[16][33];
    const size_t mn = (size_t)p.M * p.N;
    const size_t total = mn * batch;
    const size_t idx = (size_t)blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    int b = 0;
    size_t r = 0;
    if (idx < total) {
        b = (int)(idx / mn);
        r = idx - (size_t)b * mn;
        const double* P = p.partial + (size_t)b * p.splitk * mn + r;
        for (int ks = threadIdx.y; ks < p.splitk; ks += 16) s += P[(size_t)ks * mn];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && idx < total) {
#pragma unroll
        for (int k = 1; k < 16; ++k) s += sh[k][threadIdx.x];
        const int row = (int)(r / p.N);
        const int col = (int)(r - (size_t)row * p.N);
        double* dst = p.C + (long long)b * p.bsC + p.cm[row] + p.cn[col];
        double v = p.alpha * s;
        if (p.beta != 0.0) v += p.beta * (*dst);
        *dst = v;
    }
}

// kind 2: rank-K update with small K and N (<= 64): the n^5 "dressing" terms
//   C[m, n] (+)= alpha * sum_k A[m, k] * B[k, n],   M ~ n^3 rows, K, N ~ n.
// These move 24 bytes per C element for ~2K flops: HBM-bound.  One thread owns one
// row: its A row lives in registers, B (K x N, a few KB) is broadcast from shared
// memory, C is read-modify-written once, coalesced across the warp when the row
// index is the contiguous one (it is for every term of the residual).
template <int KMAX>
__global__ void __launch_bounds__(256, 2) rankk_kernel(const GemmParams p, int blocks_per_cta) {
    __shared__ __align__(16) double Bs[KMAX][64];
    __shared__ uint32_t aks[KMAX];
    __shared__ uint32_t cns[64];
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int npad = (p.N + 3) & ~3;
    const double* A = p.A + (long long)b * p.bsA;
    const double* B = p.B + (long long)b * p.bsB;
    double* C = p.C + (long long)b * p.bsC;
    // stage B (K x N) and the k / n offset tables once per CTA
    for (int idx = tid; idx < KMAX * 64; idx += 256) {
        int k = idx >> 6, n = idx & 63;
        Bs[k][n] = (k < p.K && n < p.N) ? B[(size_t)p.bk[k] + p.bn[n]] : 0.0;
    }
    if (tid < KMAX) aks[tid] = (tid < p.K) ? p.ak[tid] : 0u;
    if (tid < 64) cns[tid] = (tid < p.N) ? p.cn[tid] : 0u;
    __syncthreads();
    const bool rmw = p.beta != 0.0;
    for (int it = 0; it < blocks_per_cta; ++it) {
        const int m = (blockIdx.x * blocks_per_cta + it) * 256 + tid;
        if (m >= p.M) break;
        const double* Ar = A + p.am[m];
        double* Cr = C + p.cm[m];
        double a[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) a[k] = (k < p.K) ? Ar[aks[k]] : 0.0;
        for (int n0 = 0; n0 < npad; n0 += 8) {
            double old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                old[j] = (rmw && n0 + j < p.N) ? Cr[cns[n0 + j]] : 0.0;
            double acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k][n0]);
                const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k][n0 + 2]);
                const double2 b45 = *reinterpret_cast<const double2*>(&Bs[k][n0 + 4]);
                const double2 b67 = *reinterpret_cast<const double2*>(&Bs[k][n0 + 6]);
                acc[0] = fma(a[k], b01.x, acc[0]);
                acc[1] = fma(a[k], b01.y, acc[1]);
                acc[2] = fma(a[k], b23.x, acc[2]);
                acc[3] = fma(a[k], b23.y, acc[3]);
                acc[4] = fma(a[k], b45.x, acc[4]);
                acc[5] = fma(a[k], b45.y, acc[5]);
                acc[6] = fma(a[k], b67.x, acc[6]);
                acc[7] = fma(a[k], b67.y, acc[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (n0 + j < p.N)
                    Cr[cns[n0 + j]] = p.alpha * acc[j] + (rmw ? p.beta * old[j] : 0.0);
        }
    }
}

// kind 3: fused index-permuted sums and outer products (no contracted index)
//   C[cm[m]+cn[n]] = beta*C + sum_t alpha_t * X_t[xm_t[m]+xn_t[n]] * (Y_t ? Y_t[ym_t[m]+yn_t[n]] : 1)
// e.g. tau = t2 + t1 x t1 - t1 x t1, the four antisymmetric images rg[abij] - rg[baij] - ...,
// the drivers.  One pass over C instead of one read-modify-write pass per term.  n runs over
// the last (contiguous) one or two indices of C, m over the others, one CTA per 32 x 32 tile.
// A term whose big operand is contiguous along the fastest m letter instead is read m-fast
// and transposed through shared memory.
constexpr int EW_MAX_TERMS = 6;
struct EwTerm {
    const double* X;
    const double* Y;                           // nullptr: no second factor
    const uint32_t* xm;
    const uint32_t* xn;
    const uint32_t* ym;
    const uint32_t* yn;
    long long bsX, bsY;
    double alpha;
    int transposed;
};
struct EwParams {
    EwTerm t[EW_MAX_TERMS];
    int nterms;
    double* C;
    const uint32_t* cm;
    const uint32_t* cn;
    long long bsC;
    double beta;
    int M, N;
};

__global__ void __launch_bounds__(256) fused_ew_kernel(const __grid_constant__ EwParams p) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31;
    const int ty = threadIdx.x >> 5;   // 0..7
    const int m0 = blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    const int b = blockIdx.z;
    const int n = n0 + tx;             // n-fast role: this thread's column, rows ty + 8 i
    const int mt = m0 + tx;            // m-fast role: this thread's row, columns ty + 8 i
    const bool nv = n < p.N, mv = mt < p.M;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < p.nterms; ++t) {
        const EwTerm& q = p.t[t];
        const double* X = q.X + (long long)b * q.bsX;
        const double* Y = q.Y ? q.Y + (long long)b * q.bsY : nullptr;
        if (!q.transposed) {
            const uint32_t xo = nv ? __ldg(q.xn + n) : 0u;
            const uint32_t yo = (nv && Y) ? __ldg(q.yn + n) : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + ty + 8 * i;
                if (nv && m < p.M) {
                    double v = q.alpha * X[(size_t)__ldg(q.xm + m) + xo];
                    if (Y) v *= Y[(size_t)__ldg(q.ym + m) + yo];
                    acc[i] += v;
                }
            }
        } else {
            const uint32_t xo = mv ? __ldg(q.xm + mt) : 0u;
            const uint32_t yo = (mv && Y) ? __ldg(q.ym + mt) : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nl = ty + 8 * i;
                const int nn = n0 + nl;
                double v = 0.0;
                if (mv && nn < p.N) {
                    v = q.alpha * X[(size_t)xo + __ldg(q.xn + nn)];
                    if (Y) v *= Y[(size_t)yo + __ldg(q.yn + nn)];
                }
                tile[tx][nl] = v;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += tile[ty + 8 * i][tx];
            __syncthreads();
        }
    }
    if (nv) {
        double* C = p.C + (long long)b * p.bsC + __ldg(p.cn + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty + 8 * i;
            if (m < p.M) {
                double* dst = C + __ldg(p.cm + m);
                double v = acc[i];
                if (p.beta != 0.0) v += p.beta * (*dst);
                *dst = v;
            }
        }
    }
}

// kind 1: C[b][cm[m]+cn[n]] = beta*C + alpha*A[b][am[m]+ak[n]]  (32x32 smem transpose tile)
struct PermParams {
    const double* A;
    double* C;
    const uint32_t* am;
    const uint32_t* an;
    const uint32_t* cm;
    const uint32_t* cn;
    int M, N;
    long long bsA, bsC;
    double alpha, beta;
    int a_mode, c_mode;   // 0: contiguous along n, 1: contiguous along m
};

__global__ void __launch_bounds__(256) permute_axpby_kernel(const PermParams p) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31;
    const int ty = threadIdx.x >> 5;   // 0..7
    const int m0 = blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    const int b = blockIdx.z;
    const double* A = p.A + (long long)b * p.bsA;
    double* C = p.C + (long long)b * p.bsC;
    if (p.a_mode == 0) {
        int n = n0 + tx;
        uint32_t no = (n < p.N) ? p.an[n] : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int ml = ty + 8 * i;
            int m = m0 + ml;
            tile[ml][tx] = (m < p.M && n < p.N) ? A[(size_t)p.am[m] + no] : 0.0;
        }
    } else {
        int m = m0 + tx;
        uint32_t mo = (m < p.M) ? p.am[m] : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int nl = ty + 8 * i;
            int n = n0 + nl;
            tile[tx][nl] = (m < p.M && n < p.N) ? A[(size_t)mo + p.an[n]] : 0.0;
        }
    }
    __syncthreads();
    if (p.c_mode == 0) {
        int n = n0 + tx;
        uint32_t no = (n < p.N) ? p.cn[n] : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int ml = ty + 8 * i;
            int m = m0 + ml;
            if (m < p.M && n < p.N) {
                double* dst = C + (size_t)p.cm[m] + no;
                double v = p.alpha * tile[ml][tx];
                if (p.beta != 0.0) v += p.beta * (*dst);
                *dst = v;
            }
        }
    } else {
        int m = m0 + tx;
        uint32_t mo = (m < p.M) ? p.cm[m] : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int nl = ty + 8 * i;
            int n = n0 + nl;
            if (m < p.M && n < p.N) {
                double* dst = C + (size_t)mo + p.cn[n];
                double v = p.alpha * tile[tx][nl];
                if (p.beta != 0.0) v += p.beta * (*dst);
                *dst = v;
            }
        }
    }
}

}  // namespace kb200
