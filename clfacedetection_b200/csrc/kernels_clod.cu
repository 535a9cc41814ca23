// kernels_clod.cu -- sliding-window Haar-cascade evaluation for sm_100a.  Replaces, from
// scratch, the reference's per-stage runStage kernel (clod.cl:32-93) and its host loop
// with a blocking round trip per stage per scale (clod.cpp:1270-1302, 789-818), with the
// window semantics of cvRunHaarClassifierCascadeSum (tempcv.cpp:795-972) on the pyramid
// grid of HaarDetectObjects_ScaleImage_Invoker (tempcv.cpp:1011-1103).
//
// Two kernels per cascade per batch, no host round trip in between:
//   k_cascade_tiles : one CTA per 64x32-window tile of one level of one frame.  The int32
//       integral tile is staged into shared memory with TMA bulk row copies
//       (cp.async.bulk + mbarrier), sigma is computed once per window in FP64, and the
//       leading ("dense") stages are evaluated one thread per window with a warp-ballot
//       stream compaction of the surviving window list after EVERY stage, so warps stay
//       full despite the steep early-exit profile.  Stumps come from the constant bank
//       (the packed cascade is a __grid_constant__ kernel parameter, <= 32 KB).
//   k_cascade_deep  : survivors of the dense prefix from all tiles / levels / frames are
//       pooled in one global queue and evaluated one WARP per window, lanes striding over
//       the trees of a stage (handles multi-node trees, tilted features and the alt_tree
//       stage tree).  The stage sum is reduced with shuffles when the packer proved the
//       alpha sum exact in any order, otherwise accumulated in tree order.
//
// Arithmetic is bit-identical to the reference's C expressions: integer rect sums; FP64
// variance with separately rounded mul/sub/sqrt; two_rects stages multiply in double
// (products of a <2^24 integer and a 24-bit float are exact, so one FMA equals mul+add);
// other stages multiply in FLOAT and accumulate in double (tempcv.cpp:782-786,907-910).
// Tensor cores are not used: this is gather + compare work, not a contraction.
#include <cstdint>

#include "clfd_pack.h"
#include "kernels.h"

namespace clfd {

typedef unsigned long long ull;

// ------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + TMA bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double window_sigma(int s4, ull q4, double inv_area) {
    // tempcv.cpp:824-832, every operation rounded separately
    const double mean = __dmul_rn((double)s4, inv_area);
    const double v = __dsub_rn(__dmul_rn((double)q4, inv_area), __dmul_rn(mean, mean));
    return v >= 0. ? sqrt(v) : 1.;
}

__device__ __forceinline__ void emit_rect(const CascadeArgs &a, const CasLevel &CL, int frame, int px, int py) {
    const ull slot = atomicAdd(a.counters + 0, 1ull);
    if (slot < a.rect_cap) {
        DevRect r;
        r.x = __double2int_rn(__dmul_rn((double)px, CL.factor));  // cvRound(x*factor), tempcv.cpp:1099
        r.y = __double2int_rn(__dmul_rn((double)py, CL.factor));
        r.w = CL.win_w; r.h = CL.win_h; r.frame = frame; r.cascade = a.cascade_index;
        a.rects[slot] = r;
    } else {
        atomicAdd(a.counters + 2, 1ull);
    }
}

// ------------------------------------------------------------------------------------
// dense tile kernel
//
// One CTA (128 threads) per 64x16-window tile.  Shared memory:
//   tile    int32 [rows][S]   ystep-1 levels: natural layout, staged by TMA bulk row copies;
//                             ystep-2 levels: even columns in [0,S/2), odd columns in [S/2,S)
//                             of each row (LDG.128 + 2 x STS.64), so that in both layouts a
//                             window's base word is f(wy)*S + wx and the 32 lanes of a warp
//                             (consecutive wx) hit 32 different banks on every corner load.
//   sigma   double [1024]     per-window variance normaliser (FP64, exact)
//   list    u16 [2][1024]     compacted survivor lists (ping-pong)
//
// Phase 1, "fixed geometry" (stages 0 .. n_fixed-1, where most windows are still alive):
//   thread t owns the column of 8 windows (wx = t & 63, wy = (t >> 6) + 2k).  Their tile
//   addresses differ by a compile-time constant, so a corner address is computed ONCE per
//   stump and the 4 windows of a chunk are read with immediate offsets (LDS [R + k*ROWSTEP]):
//   no per-window address arithmetic, no compaction traffic, conflict-free banks.  Dead
//   windows ride along; a chunk whose 4 x 32 windows are all dead is skipped (warp-uniform).
// Phase 2, compacted (remaining dense stages): survivors are stream-compacted with a warp
//   ballot after every stage; a warp takes rows of 32 list entries, up to 4 rows per pass.
// Survivors of the last dense stage are appended to the global queue of the deep kernel.
//
// Stage arithmetic: an FP32 filter decides each stump; whenever |s32 - t32| is inside a
// guard band (2^-20 |t32| plus the cancellation terms) the window's whole stage is redone by
// dense_eval_stage_exact(), which reproduces the reference's C expressions bit for bit.
// Outside the band both agree by the error analysis in DESIGN.md, so results are identical
// to the all-FP64 evaluation (tests also run with force_exact = 1 and compare).
// ------------------------------------------------------------------------------------
#define TILE_LD(base, off) (*reinterpret_cast<const int *>((base) + (off)))

struct DenseSmemPlan {
    size_t tile, sigma, list, blist, ctl, bar, total;
};
// control block (ints): [0..31] bucket counts, [32..63] snapshot of the counts, [64..95] round
// masks, [96..127] round prefixes, [128] survivors, [130..131] queue base
constexpr int kCtlBcnt = 0, kCtlBcopy = 32, kCtlRmask = 64, kCtlRpref = 96, kCtlAlive = 128, kCtlQueue = 130, kCtlInts = 136;
__host__ __device__ inline DenseSmemPlan dense_smem_plan(const DenseParams &P) {
    DenseSmemPlan p;
    const size_t rows = (size_t)(kTileH - 1) * P.ystep + P.win_h + 1;
    p.tile = 0;
    p.sigma = (rows * P.tile_stride * 4 + 127) & ~(size_t)127;
    p.list = p.sigma + kTileWindows * sizeof(double);
    p.blist = p.list + kTileWindows * sizeof(uint16_t);
    p.ctl = p.blist + kTileWindows * sizeof(uint16_t);
    p.bar = p.ctl + kCtlInts * sizeof(int);
    p.total = p.bar + 16;
    return p;
}
size_t dense_smem_bytes(const DenseParams &P) { return dense_smem_plan(P).total; }

// Exact evaluation of one dense stage for one window (the reference's arithmetic).
__device__ __noinline__ bool dense_eval_stage_exact(const DenseParams &P, int s, const unsigned char *base, double sigma) {
    const int first = P.stage[s].first, count = P.stage[s].count;
    const bool dbl = P.stage[s].flags & 1u;
    double S = 0.0;
    for (int j = 0; j < count; j++) {
        const DenseStump &q = P.stump[first + j];
        const int r0 = TILE_LD(base, q.off[0]) - TILE_LD(base, q.off[1]) - TILE_LD(base, q.off[2]) + TILE_LD(base, q.off[3]);
        const int r1 = TILE_LD(base, q.off[4]) - TILE_LD(base, q.off[5]) - TILE_LD(base, q.off[6]) + TILE_LD(base, q.off[7]);
        const double t = __dmul_rn((double)q.thr, sigma);
        double sum;
        if (dbl) {  // tempcv.cpp:872-898; both products are exact in double, so fma == mul, mul, add
            sum = __fma_rn((double)r1, (double)q.w[1], __dmul_rn((double)r0, (double)q.w[0]));
        } else {    // tempcv.cpp:899-930 / 782-786: float products, double accumulation
            sum = __dadd_rn((double)__fmul_rn(__int2float_rn(r0), q.w[0]), (double)__fmul_rn(__int2float_rn(r1), q.w[1]));
            if (q.off[11] != 0) {
                const int r2 = TILE_LD(base, q.off[8]) - TILE_LD(base, q.off[9]) - TILE_LD(base, q.off[10]) + TILE_LD(base, q.off[11]);
                sum = __dadd_rn(sum, (double)__fmul_rn(__int2float_rn(r2), q.w[2]));
            }
        }
        S = __dadd_rn(S, sum >= t ? q.a1 : q.a0);
    }
    return S >= (double)P.stage[s].thr;
}

// FP32-filtered evaluation of stage s for K windows of this thread.
//   FIXED = true : window k lives at base0 + k*ROWSTEP (compile-time) -> immediate offsets
//   FIXED = false: window k lives at base[k]
template <int K, bool DBL, bool HAS3, bool FIXED, int ROWSTEP>
__device__ __forceinline__ void dense_filter_stage(const DenseParams &P, int s, const unsigned char *const (&base)[K],
                                                   const float (&sg)[K], double (&S)[K], bool (&near)[K]) {
    const int first = P.stage[s].first, count = P.stage[s].count;
    const float eps = P.filter_eps, eps4 = eps * 0.25f;
#pragma unroll 1
    for (int j = 0; j < count; j++) {
        const DenseStump &q = P.stump[first + j];
        const float w0 = q.w[0], w1 = q.w[1], thr = q.thr;
        const double al0 = q.a0, al1 = q.a1;
        const bool three = HAS3 && (q.off[11] != 0);   // warp-uniform
        // corner pointers of window 0 (shared by all K windows in fixed geometry)
        const unsigned char *c00 = base[0] + q.off[0], *c01 = base[0] + q.off[1], *c02 = base[0] + q.off[2], *c03 = base[0] + q.off[3];
        const unsigned char *c10 = base[0] + q.off[4], *c11 = base[0] + q.off[5], *c12 = base[0] + q.off[6], *c13 = base[0] + q.off[7];
#pragma unroll
        for (int k = 0; k < K; k++) {
            int r0, r1;
            if (FIXED) {
                r0 = TILE_LD(c00, k * ROWSTEP) - TILE_LD(c01, k * ROWSTEP) - TILE_LD(c02, k * ROWSTEP) + TILE_LD(c03, k * ROWSTEP);
                r1 = TILE_LD(c10, k * ROWSTEP) - TILE_LD(c11, k * ROWSTEP) - TILE_LD(c12, k * ROWSTEP) + TILE_LD(c13, k * ROWSTEP);
            } else {
                const unsigned char *b = base[k];
                r0 = TILE_LD(b, q.off[0]) - TILE_LD(b, q.off[1]) - TILE_LD(b, q.off[2]) + TILE_LD(b, q.off[3]);
                r1 = TILE_LD(b, q.off[4]) - TILE_LD(b, q.off[5]) - TILE_LD(b, q.off[6]) + TILE_LD(b, q.off[7]);
            }
            const float p0 = __fmul_rn(__int2float_rn(r0), w0), p1 = __fmul_rn(__int2float_rn(r1), w1);
            float s32 = __fadd_rn(p0, p1);
            const float t32 = __fmul_rn(thr, sg[k]);
            float m = __fmul_rn(fabsf(t32), eps);
            if (DBL) {
                // reference adds the two EXACT products: fp32 product errors do not cancel
                m = __fadd_rn(m, __fmul_rn(__fadd_rn(fabsf(p0), fabsf(p1)), eps4));
            }
            if (HAS3) {
                if (three) {
                    const unsigned char *b = FIXED ? base[0] + k * ROWSTEP : base[k];
                    const int r2 = TILE_LD(b, q.off[8]) - TILE_LD(b, q.off[9]) - TILE_LD(b, q.off[10]) + TILE_LD(b, q.off[11]);
                    m = __fadd_rn(m, __fmul_rn(fabsf(s32), eps4));   // rounding of the first add
                    s32 = __fadd_rn(s32, __fmul_rn(__int2float_rn(r2), q.w[2]));
                }
            }
            const float d = __fadd_rn(s32, -t32);
            near[k] = near[k] || (fabsf(d) <= m);
            S[k] = __dadd_rn(S[k], d >= 0.f ? al1 : al0);
        }
    }
}

template <int K, bool FIXED, int ROWSTEP>
__device__ __forceinline__ void dense_dispatch_stage(const DenseParams &P, int s, const unsigned char *const (&base)[K],
                                                     const float (&sg)[K], double (&S)[K], bool (&near)[K]) {
    const uint32_t flags = P.stage[s].flags;
    if (flags & 1u) dense_filter_stage<K, true, false, FIXED, ROWSTEP>(P, s, base, sg, S, near);
    else if (flags & 2u) dense_filter_stage<K, false, true, FIXED, ROWSTEP>(P, s, base, sg, S, near);
    else dense_filter_stage<K, false, false, FIXED, ROWSTEP>(P, s, base, sg, S, near);
}

struct DenseCtx {
    unsigned char *tile;
    double *sigma;
    int16_t *codes;   // this frame + level, or nullptr
    int row_mul;      // bytes between window rows in the tile
    int tx, ty, nx;
    int code_mul;
};

__device__ __forceinline__ void dense_write_code(const DenseCtx &c, int wid, int code) {
    const int wx = wid & (kTileW - 1), wy = wid / kTileW;
    c.codes[(size_t)(c.ty * kTileH + wy) * c.nx + c.tx * kTileW + wx] = (int16_t)code;
}

// survivors go to the bucket of their bank residue (wx mod 32): blist[slot][bucket]
__device__ __forceinline__ void dense_append(uint16_t *blist, int *bcnt, int wid) {
    const int b = wid & 31;
    const int slot = atomicAdd(&bcnt[b], 1);
    blist[slot * 32 + b] = (uint16_t)wid;
}

// compacted phase: evaluate stage s for K list entries of this thread
template <int K>
__device__ __forceinline__ void dense_run_rows(const DenseParams &P, const DenseCtx &c, int s, const int (&wid)[K],
                                               const bool (&act)[K], uint16_t *blist, int *bcnt) {
    const unsigned char *base[K];
    float sg[K];
    double S[K];
    bool near[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int wx = wid[k] & (kTileW - 1), wy = wid[k] / kTileW;
        base[k] = c.tile + wy * c.row_mul + wx * 4;
        sg[k] = (float)c.sigma[wid[k]];
        S[k] = 0.0;
        near[k] = P.force_exact != 0;
    }
    dense_dispatch_stage<K, false, 0>(P, s, base, sg, S, near);
    const double sthr = (double)P.stage[s].thr;
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!act[k]) continue;
        bool pass = S[k] >= sthr;
        if (near[k]) pass = dense_eval_stage_exact(P, s, base[k], c.sigma[wid[k]]);
        if (pass) dense_append(blist, bcnt, wid[k]);
        else if (c.codes) dense_write_code(c, wid[k], s * c.code_mul);
    }
}

// Turn the bucketed survivors into a linear list in ROUND-ROBIN bucket order (slot 0 of every
// non-empty bucket, then slot 1, ...): a row of 32 consecutive entries then holds (nearly)
// distinct bank residues, so phase-2 corner loads are (nearly) conflict free while rows stay
// full.  Returns the number of survivors.  Three barriers; all threads must call it.
__device__ __forceinline__ int dense_reorder(uint16_t *list, const uint16_t *blist, int *ctl, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    __syncthreads();   // all appends done
    if (warp == 0) {
        const int cl = ctl[kCtlBcnt + lane];
        const int kmax = __reduce_max_sync(0xffffffffu, cl);
        unsigned mymask = 0;   // lane r keeps the mask of buckets that have a slot r
        for (int r = 0; r < kmax; r++) {
            const unsigned m = __ballot_sync(0xffffffffu, cl > r);
            if (lane == r) mymask = m;
        }
        const int n = __popc(mymask);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        ctl[kCtlRmask + lane] = (int)mymask;
        ctl[kCtlRpref + lane] = incl - n;
        ctl[kCtlBcopy + lane] = cl;
        ctl[kCtlBcnt + lane] = 0;
        if (lane == 31) ctl[kCtlAlive] = incl;
    }
    __syncthreads();
    {
        const int mine = ctl[kCtlBcopy + lane];
        for (int r = warp; r < mine; r += kDenseWarps) {
            const unsigned m = (unsigned)ctl[kCtlRmask + r];
            list[ctl[kCtlRpref + r] + __popc(m & ((1u << lane) - 1u))] = blist[r * 32 + lane];
        }
    }
    __syncthreads();
    return ctl[kCtlAlive];
}

// ROWSTEP_T: compile-time byte distance between a thread's consecutive windows in phase 1
// (= 2 window rows), or 0 to use the runtime value (generic window sizes).
template <int ROWSTEP_T>
__global__ void __launch_bounds__(kDenseThreads)
k_cascade_tiles(const __grid_constant__ DenseParams P, const __grid_constant__ CascadeArgs a, const int tile0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DenseSmemPlan plan = dense_smem_plan(P);
    unsigned char *tile = smem_raw + plan.tile;
    double *sigma = reinterpret_cast<double *>(smem_raw + plan.sigma);
    uint16_t *list = reinterpret_cast<uint16_t *>(smem_raw + plan.list);
    uint16_t *blist = reinterpret_cast<uint16_t *>(smem_raw + plan.blist);
    int *ctl = reinterpret_cast<int *>(smem_raw + plan.ctl);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + plan.bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const int tile_id = tile0 + blockIdx.x;

    int cl = 0;   // which level does this tile belong to?
    {
        int lo = 0, hi = a.n_cas_levels - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(&a.cas_levels[mid].tile_base) <= tile_id) lo = mid; else hi = mid - 1;
        }
        cl = lo;
    }
    const CasLevel CL = a.cas_levels[cl];
    const PyrLevel L = a.levels[CL.pyr_level];
    const int local = tile_id - CL.tile_base;
    const int tx = local % CL.tiles_x, ty = local / CL.tiles_x;
    const int ystep = P.ystep;   // == CL.ystep by construction of the launch
    const int px0 = tx * kTileW * ystep, py0 = ty * kTileH * ystep;   // tile origin in the integral image
    const int n_wx = min(kTileW, CL.nx - tx * kTileW), n_wy = min(kTileH, CL.ny - ty * kTileH);
    const int rows = min((kTileH - 1) * ystep + P.win_h + 1, L.h + 1 - py0);
    const int cols = ((kTileW - 1) * ystep + P.win_w + 1 + 3) & ~3;
    const int S = P.tile_stride;

    const size_t frame_off = (size_t)frame * a.sum_frame_stride + L.sum_off;
    const int32_t *__restrict__ gsum = a.sum + frame_off + (size_t)py0 * L.sum_pitch + px0;
    const ull *__restrict__ gsq = a.sq + frame_off + (size_t)py0 * L.sum_pitch + px0;

    // ---- stage the integral tile ----
    for (int i = tid; i < kCtlInts; i += kDenseThreads) ctl[i] = 0;
    if (ystep == 1) {   // natural layout: one TMA bulk copy per row
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(rows * cols * 4));
        for (int r = tid; r < rows; r += kDenseThreads)
            tma_bulk_g2s(tile + (size_t)r * S * 4, gsum + (size_t)r * L.sum_pitch, (uint32_t)(cols * 4), bar);
        mbar_wait(bar, 0);
    } else {            // de-interleave columns: x -> (x&1)*S/2 + (x>>1)
        const int c4 = cols >> 2, total = rows * c4;
        for (int i = tid; i < total; i += kDenseThreads) {
            const int r = i / c4, cq = i - r * c4;
            const int4 v = __ldg(reinterpret_cast<const int4 *>(gsum + (size_t)r * L.sum_pitch) + cq);
            int *row = reinterpret_cast<int *>(tile) + r * S;
            *reinterpret_cast<int2 *>(row + 2 * cq) = make_int2(v.x, v.z);
            *reinterpret_cast<int2 *>(row + (S >> 1) + 2 * cq) = make_int2(v.y, v.w);
        }
        __syncthreads();
    }

    DenseCtx c;
    c.tile = tile; c.sigma = sigma;
    c.codes = a.codes ? a.codes + (size_t)frame * a.windows_per_frame + CL.win_base : nullptr;
    c.row_mul = ystep * S * 4;
    c.tx = tx; c.ty = ty; c.nx = CL.nx;
    c.code_mul = P.is_tree ? 2 : 1;

    // ---- phase 1: sigma, then the fixed-geometry stages ----
    constexpr int kRowsPerSlot = kDenseThreads / kTileW;   // window rows between a thread's slots (2)
    const int rowstep = ROWSTEP_T ? ROWSTEP_T : kRowsPerSlot * c.row_mul;
    const int wx = tid & (kTileW - 1), wy0 = tid / kTileW;
    uint32_t alive = 0;   // bit k: window (wx, wy0 + 2k) still alive
    {
        const int eq_w = P.win_w - 2, eq_h = P.win_h - 2;
        int e[4];   // equRect corners (1,1),(1,1+eq_w),(1+eq_h,1),(1+eq_h,1+eq_w) as tile byte offsets
        {
            const int ys[4] = {1, 1, 1 + eq_h, 1 + eq_h}, xs[4] = {1, 1 + eq_w, 1, 1 + eq_w};
#pragma unroll
            for (int q = 0; q < 4; q++)
                e[q] = 4 * (ystep == 1 ? ys[q] * S + xs[q] : ys[q] * S + (xs[q] & 1) * (S >> 1) + (xs[q] >> 1));
        }
        const int g0 = L.sum_pitch + 1, g1 = g0 + eq_w, g2 = (1 + eq_h) * L.sum_pitch + 1, g3 = g2 + eq_w;
#pragma unroll
        for (int k = 0; k < kDenseSlots; k++) {
            const int wy = wy0 + k * kRowsPerSlot;
            if (wx < n_wx && wy < n_wy) {
                alive |= 1u << k;
                const unsigned char *base = tile + wy * c.row_mul + wx * 4;
                const int s4 = TILE_LD(base, e[0]) - TILE_LD(base, e[1]) - TILE_LD(base, e[2]) + TILE_LD(base, e[3]);
                const ull *q = gsq + (size_t)(wy * ystep) * L.sum_pitch + wx * ystep;
                const ull q4 = __ldg(q + g0) - __ldg(q + g1) - __ldg(q + g2) + __ldg(q + g3);
                sigma[wy * kTileW + wx] = window_sigma(s4, q4, P.inv_area);
            } else {
                sigma[wy * kTileW + wx] = 1.0;
            }
        }
    }
    // a thread reads back only sigmas it wrote itself: no barrier needed in phase 1
    int s = 0;
    for (; s < P.n_fixed; s++) {
#pragma unroll 1
        for (int k0 = 0; k0 < kDenseSlots; k0 += kDenseChunk) {
            const uint32_t m4 = (alive >> k0) & ((1u << kDenseChunk) - 1u);
            if (!__any_sync(0xffffffffu, m4 != 0)) continue;   // whole 4 x 32 block is dead
            const unsigned char *base[kDenseChunk];
            float sg[kDenseChunk];
            double Ssum[kDenseChunk];
            bool near[kDenseChunk];
#pragma unroll
            for (int k = 0; k < kDenseChunk; k++) {
                const int wy = wy0 + (k0 + k) * kRowsPerSlot;
                base[k] = tile + wy * c.row_mul + wx * 4;   // == base[0] + k * rowstep
                sg[k] = (float)sigma[wy * kTileW + wx];
                Ssum[k] = 0.0;
                near[k] = P.force_exact != 0;
            }
            if (ROWSTEP_T) dense_dispatch_stage<kDenseChunk, true, ROWSTEP_T>(P, s, base, sg, Ssum, near);
            else dense_dispatch_stage<kDenseChunk, false, 0>(P, s, base, sg, Ssum, near);
            const double sthr = (double)P.stage[s].thr;
#pragma unroll
            for (int k = 0; k < kDenseChunk; k++) {
                if (!((m4 >> k) & 1u)) continue;
                const int wid = (wy0 + (k0 + k) * kRowsPerSlot) * kTileW + wx;
                bool pass = Ssum[k] >= sthr;
                if (near[k]) pass = dense_eval_stage_exact(P, s, base[k], sigma[wid]);
                if (!pass) {
                    alive &= ~(1u << (k0 + k));
                    if (c.codes) dense_write_code(c, wid, s * c.code_mul);
                }
            }
        }
    }
    (void)rowstep;

    // ---- phase-1 survivors -> buckets -> bank-friendly linear list ----
#pragma unroll
    for (int k = 0; k < kDenseSlots; k++)
        if ((alive >> k) & 1u) dense_append(blist, ctl + kCtlBcnt, (wy0 + k * kRowsPerSlot) * kTileW + wx);
    int n_alive = dense_reorder(list, blist, ctl, tid);

    // ---- phase 2: remaining dense stages on compacted lists ----
    for (;;) {
        if (n_alive == 0) return;
        if (s >= P.n_stages || (n_alive <= kHandoffWindows && P.n_stages < P.total_stages)) break;
        const int n_rows = (n_alive + 31) >> 5;
        for (int r0 = warp; r0 < n_rows; r0 += kDenseWarps * kDenseChunk) {
            int wid[kDenseChunk];
            bool act[kDenseChunk];
            int K = 0;
#pragma unroll
            for (int k = 0; k < kDenseChunk; k++) {
                const int row = r0 + k * kDenseWarps;
                const int i = row * 32 + lane;
                act[k] = row < n_rows && i < n_alive;
                wid[k] = act[k] ? list[i] : lane;
                K += row < n_rows;
            }
            if (K == 1) {
                const int w1[1] = {wid[0]}; const bool a1[1] = {act[0]};
                dense_run_rows<1>(P, c, s, w1, a1, blist, ctl + kCtlBcnt);
            } else if (K == 2) {
                const int w2[2] = {wid[0], wid[1]}; const bool a2[2] = {act[0], act[1]};
                dense_run_rows<2>(P, c, s, w2, a2, blist, ctl + kCtlBcnt);
            } else if (K == 3) {
                const int w3[3] = {wid[0], wid[1], wid[2]}; const bool a3[3] = {act[0], act[1], act[2]};
                dense_run_rows<3>(P, c, s, w3, a3, blist, ctl + kCtlBcnt);
            } else {
                dense_run_rows<4>(P, c, s, wid, act, blist, ctl + kCtlBcnt);
            }
        }
        n_alive = dense_reorder(list, blist, ctl, tid);
        s++;
    }

    // ---- survivors: accepted (whole cascade was dense) or handed to the deep kernel ----
    const uint16_t *lin = list;
    if (s >= P.total_stages) {
        for (int i = tid; i < n_alive; i += kDenseThreads) {
            const int w = lin[i];
            emit_rect(a, CL, frame, px0 + (w & (kTileW - 1)) * ystep, py0 + (w / kTileW) * ystep);
            if (c.codes) dense_write_code(c, w, P.total_stages);
        }
    } else {
        if (tid == 0) *reinterpret_cast<ull *>(ctl + kCtlQueue) = atomicAdd(a.counters + 1, (ull)n_alive);
        __syncthreads();
        const ull qb = *reinterpret_cast<const ull *>(ctl + kCtlQueue);
        for (int i = tid; i < n_alive; i += kDenseThreads) {
            const int w = lin[i];
            if (qb + i < a.queue_cap) {
                QueueItem it;
                it.key = ((uint32_t)frame << 16) | ((uint32_t)cl << 8) | (uint32_t)s;
                it.xy = ((uint32_t)(py0 + (w / kTileW) * ystep) << 16) | (uint32_t)(px0 + (w & (kTileW - 1)) * ystep);
                a.queue[qb + i] = it;
            } else {
                atomicAdd(a.counters + 3, 1ull);
            }
        }
    }
}

template <int ROWSTEP_T>
static cudaError_t launch_tiles_t(const DenseParams &P, const CascadeArgs &a, int tile0, int n_tiles, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_cascade_tiles<ROWSTEP_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    k_cascade_tiles<ROWSTEP_T><<<dim3(n_tiles, a.n_frames), kDenseThreads, smem, stream>>>(P, a, tile0);
    return cudaGetLastError();
}

cudaError_t launch_cascade_tiles(const DenseParams &P, const CascadeArgs &a, int tile0, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0 || a.n_frames == 0) return cudaSuccess;
    const size_t smem = dense_smem_bytes(P);
    // byte distance between a thread's consecutive phase-1 windows: 2 window rows
    const int rowstep = (kDenseThreads / kTileW) * P.ystep * P.tile_stride * 4;
    switch (rowstep) {   // common window sizes get immediate-offset code (20x20 / 24x24 cascades)
        case 2 * 1 * 96 * 4:  return launch_tiles_t<2 * 1 * 96 * 4>(P, a, tile0, n_tiles, smem, stream);
        case 2 * 2 * 160 * 4: return launch_tiles_t<2 * 2 * 160 * 4>(P, a, tile0, n_tiles, smem, stream);
        default:              return launch_tiles_t<0>(P, a, tile0, n_tiles, smem, stream);
    }
}

// ------------------------------------------------------------------------------------
// enqueue-all: cascades without a dense prefix start every window in the deep kernel
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_enqueue_all(const CascadeArgs a) {
    const long long w = (long long)blockIdx.x * 256 + threadIdx.x;
    const int frame = blockIdx.y;
    if (w == 0 && frame == 0) a.counters[1] = (ull)a.windows_per_frame * a.n_frames;
    if (w >= a.windows_per_frame) return;
    int lo = 0, hi = a.n_cas_levels - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&a.cas_levels[mid].win_base) <= w) lo = mid; else hi = mid - 1;
    }
    const int nx = __ldg(&a.cas_levels[lo].nx), ystep = __ldg(&a.cas_levels[lo].ystep);
    const int local = (int)(w - __ldg(&a.cas_levels[lo].win_base));
    const int iy = local / nx, ix = local - iy * nx;
    QueueItem it;
    it.key = ((uint32_t)frame << 16) | ((uint32_t)lo << 8);
    it.xy = ((uint32_t)(iy * ystep) << 16) | (uint32_t)(ix * ystep);
    a.queue[(size_t)frame * a.windows_per_frame + w] = it;
}

cudaError_t launch_enqueue_all(const CascadeArgs &a, cudaStream_t stream) {
    if (a.windows_per_frame == 0 || a.n_frames == 0) return cudaSuccess;
    k_enqueue_all<<<dim3((unsigned)((a.windows_per_frame + 255) / 256), a.n_frames), 256, 0, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// deep kernel: one warp per queued window
// ------------------------------------------------------------------------------------
constexpr int kDeepThreads = 256;

__device__ __forceinline__ int rect_sum_g(const int32_t *__restrict__ base, int pitch, uint32_t dxw, uint32_t dyw) {
    // dxw / dyw hold the 4 corner coordinates of one rectangle, one byte each
    const int o0 = (int)(dyw & 255u) * pitch + (int)(dxw & 255u);
    const int o1 = (int)((dyw >> 8) & 255u) * pitch + (int)((dxw >> 8) & 255u);
    const int o2 = (int)((dyw >> 16) & 255u) * pitch + (int)((dxw >> 16) & 255u);
    const int o3 = (int)(dyw >> 24) * pitch + (int)(dxw >> 24);
    return __ldg(base + o0) - __ldg(base + o1) - __ldg(base + o2) + __ldg(base + o3);
}

__global__ void __launch_bounds__(kDeepThreads) k_cascade_deep(const __grid_constant__ CascadeArgs a) {
    const int lane = threadIdx.x & 31;
    const ull warp0 = ((ull)blockIdx.x * kDeepThreads + threadIdx.x) >> 5;
    const ull nwarps = ((ull)gridDim.x * kDeepThreads) >> 5;
    ull n = a.counters[1];
    if (n > a.queue_cap) n = a.queue_cap;
    const DeepCascadeDev &D = a.deep;

    for (ull item = warp0; item < n; item += nwarps) {
        const QueueItem q = a.queue[item];
        const int frame = q.key >> 16, cl = (q.key >> 8) & 255, stage0 = q.key & 255;
        const int x = q.xy & 0xffff, y = q.xy >> 16;
        const CasLevel CL = a.cas_levels[cl];
        const PyrLevel L = a.levels[CL.pyr_level];
        const int pitch = L.sum_pitch;
        const size_t off = (size_t)frame * a.sum_frame_stride + L.sum_off + (size_t)y * pitch + x;
        const int32_t *__restrict__ sum = a.sum + off;
        const int32_t *__restrict__ til = a.tilted ? a.tilted + off : sum;
        const ull *__restrict__ sq = a.sq + off;

        const int eq_w = D.win_w - 2, eq_h = D.win_h - 2;
        const int g0 = pitch + 1, g1 = g0 + eq_w, g2 = (1 + eq_h) * pitch + 1, g3 = g2 + eq_w;
        const int s4 = __ldg(sum + g0) - __ldg(sum + g1) - __ldg(sum + g2) + __ldg(sum + g3);
        const ull q4 = __ldg(sq + g0) - __ldg(sq + g1) - __ldg(sq + g2) + __ldg(sq + g3);
        const double sigma = window_sigma(s4, q4, D.inv_area);

        int ptr = stage0, last = stage0, accepted = 0;
        for (;;) {
            const DeepStage st = D.stages[ptr];
            last = ptr;
            const bool dbl = st.flags & 1, order_free = st.flags & 2;
            double S = 0.0, part = 0.0;
            for (int j0 = 0; j0 < st.ntrees; j0 += 32) {
                const int j = j0 + lane;
                float av = 0.f;
                if (j < st.ntrees) {
                    const int tree = st.first_tree + j;
                    const int n0 = __ldg(D.tree_first_node + tree);
                    int idx = 0;
                    do {
                        const uint4 *nd = reinterpret_cast<const uint4 *>(D.nodes + n0 + idx);
                        const uint4 c0 = __ldg(nd), c1 = __ldg(nd + 1), c2 = __ldg(nd + 2);
                        const int flags = __ldg(reinterpret_cast<const int *>(nd + 3));
                        // c0 = dx[0..11], dy[0..3]; c1 = dy[4..11], w0, w1; c2 = w2, thr, left, right
                        const int32_t *__restrict__ base = (flags & 1) ? til : sum;
                        const int r0 = rect_sum_g(base, pitch, c0.x, c0.w);
                        const int r1 = rect_sum_g(base, pitch, c0.y, c1.x);
                        const float w0 = __uint_as_float(c1.z), w1 = __uint_as_float(c1.w);
                        const float thr = __uint_as_float(c2.y);
                        const double t = __dmul_rn((double)thr, sigma);
                        double sv;
                        if (dbl) {
                            sv = __fma_rn((double)r1, (double)w1, __dmul_rn((double)r0, (double)w0));
                        } else {
                            sv = __dadd_rn((double)__fmul_rn(__int2float_rn(r0), w0), (double)__fmul_rn(__int2float_rn(r1), w1));
                            if ((flags >> 8) == 3) {
                                const int r2 = rect_sum_g(base, pitch, c0.z, c1.y);
                                sv = __dadd_rn(sv, (double)__fmul_rn(__int2float_rn(r2), __uint_as_float(c2.x)));
                            }
                        }
                        idx = sv < t ? (int)c2.z : (int)c2.w;
                    } while (idx > 0);
                    av = __ldg(D.alpha + n0 + tree - idx);
                }
                if (order_free) {
                    part = __dadd_rn(part, (double)av);
                } else {
                    const int cnt = min(32, st.ntrees - j0);
                    for (int k = 0; k < cnt; k++) S = __dadd_rn(S, (double)__shfl_sync(0xffffffffu, av, k));
                }
            }
            if (order_free) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, d));
                S = part;
            }
            const bool pass = S >= (double)st.thr;
            if (D.is_tree) {  // tempcv.cpp:849-859
                if (pass) {
                    ptr = st.child;
                    if (ptr < 0) { accepted = 1; break; }
                } else {
                    int p = ptr;
                    while (p >= 0 && __ldg(&D.stages[p].next) < 0) p = __ldg(&D.stages[p].parent);
                    if (p < 0) break;
                    ptr = __ldg(&D.stages[p].next);
                }
            } else {
                if (!pass) break;
                if (++ptr >= D.n_stages) { accepted = 1; break; }
            }
        }
        if (lane == 0) {
            if (accepted) emit_rect(a, CL, frame, x, y);
            if (a.codes) {
                const int code = D.is_tree ? 2 * last + accepted : (accepted ? D.n_stages : last);
                a.codes[(size_t)frame * a.windows_per_frame + CL.win_base + (size_t)(y / CL.ystep) * CL.nx + x / CL.ystep] = (int16_t)code;
            }
        }
    }
}

cudaError_t launch_cascade_deep(const CascadeArgs &a, int n_sms, cudaStream_t stream) {
    if (a.n_frames == 0) return cudaSuccess;
    k_cascade_deep<<<n_sms * 8, kDeepThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace clfd
