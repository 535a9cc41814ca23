// kernels_clod.cu -- sliding-window Haar-cascade evaluation for sm_100a.  Replaces, from
// scratch, the reference's per-stage runStage kernel (clod.cl:32-93) and its host loop
// with a blocking round trip per stage per scale (clod.cpp:1270-1302, 789-818), with the
// window semantics of cvRunHaarClassifierCascadeSum (tempcv.cpp:795-972) on the pyramid
// grid of HaarDetectObjects_ScaleImage_Invoker (tempcv.cpp:1011-1103).
//
// One or two kernels per cascade per batch, no host round trip in between:
//   k_cascade_tiles : one CTA per 64x16-window tile of one level of one frame.  The int32
//       integral tile is staged into shared memory (TMA bulk row copies, cp.async.bulk +
//       mbarrier, on ystep-1 levels), sigma is computed once per window in FP64, and the
//       cascade is evaluated in two phases: fixed geometry (thread per window column, no
//       compaction, stumps from the constant bank: the packed cascade is a __grid_constant__
//       kernel parameter, <= 32 KB) while most windows are alive, then warp-autonomous: a
//       warp carries 32 survivors through all remaining stages, re-dividing its lanes between
//       windows and stumps as windows die (stump records from global memory).
//       Stump-based upright cascades (frontalface_alt / _default, eye, profileface) are
//       finished inside this kernel.
//   k_cascade_deep  : for cascades the tile kernel cannot finish (multi-node trees, tilted
//       features, the alt_tree stage tree) the survivors of the dense prefix from all tiles /
//       levels / frames are pooled in one global queue and evaluated one WARP per window,
//       lanes striding over the trees of a stage.  The stage sum is reduced with shuffles
//       when the packer proved the alpha sum exact in any order, otherwise accumulated in
//       tree order.
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
//                             of each row (LDG.128 + 2 x STS.64).  In both layouts a window's
//                             base word is ystep*wy*S + wx with ystep*S = 8 (mod 32), so its
//                             bank class is (wx + 8*wy) mod 32: the 32 lanes of a warp with
//                             consecutive wx never conflict, and neither do compacted rows
//                             whose windows have distinct classes.
//   sgf     float [1024]      per-window sigma rounded to FP32 for the phase-1 filter (the FP64
//                             value is recomputed where exact arithmetic needs it)
//   list    u16 [1024]        survivors of phase 1
//   scr     384 B per warp    re-packing scratch of phase 2
//
// Phase 1, "fixed geometry" (stages 0 .. n_fixed-1, where most windows are still alive):
//   thread t owns the column of 8 windows (wx = t & 63, wy = (t >> 6) + 2k).  Their tile
//   addresses differ by a compile-time constant, so a corner address is computed ONCE per
//   stump and the 4 windows of a chunk are read with immediate offsets (LDS [R + k*ROWSTEP]):
//   no per-window address arithmetic, no compaction traffic, conflict-free banks.  Stumps
//   come from the constant bank (the packed cascade is a kernel parameter).  An FP32 filter
//   decides each stump; whenever |s32 - t32| is inside a guard band (2^-20 |t32| plus the
//   cancellation terms) the window's whole stage is redone by dense_stage_exact(), which
//   reproduces the reference's C expressions bit for bit.  Outside the band both agree by the
//   error analysis in DESIGN.md, so results are identical to the all-FP64 evaluation (tests
//   also run with force_exact = 1 and compare).
// Phase 2, warp-autonomous (all remaining stages the tile kernel knows): see
//   dense_warp_finish().  Stump-based upright cascades are finished here; the others hand
//   the survivors of their dense prefix to the queue of k_cascade_deep.
// ------------------------------------------------------------------------------------
#define TILE_LD(base, off) (*reinterpret_cast<const int *>((base) + (off)))

struct DenseSmemPlan {
    size_t tile, sgf, list, scr, ctl, bar, total;
};
// control block (ints): [0] survivors of phase 1
constexpr int kCtlAlive = 0, kCtlInts = 8;
constexpr int kScrBytesPerWarp = 32 * (sizeof(double) + sizeof(int));
__host__ __device__ inline DenseSmemPlan dense_smem_plan(const DenseParams &P) {
    DenseSmemPlan p;
    const size_t rows = (size_t)(kTileH - 1) * P.ystep + P.win_h + 1;
    p.tile = 0;
    p.sgf = (rows * P.tile_stride * 4 + 127) & ~(size_t)127;
    p.list = p.sgf + kTileWindows * sizeof(float);
    p.scr = p.list + kTileWindows * sizeof(uint16_t);
    p.ctl = p.scr + (size_t)kDenseWarps * kScrBytesPerWarp;
    p.bar = p.ctl + kCtlInts * sizeof(int);
    p.total = p.bar + 16;
    return p;
}
size_t dense_smem_bytes(const DenseParams &P) { return dense_smem_plan(P).total; }

struct DenseCtx {
    unsigned char *tile;
    float *sgf;
    const ull *gsq;   // squared integral at the tile origin
    int16_t *codes;   // this frame + level, or nullptr
    int sq_pitch;     // elements
    int row_mul;      // bytes between window rows in the tile
    int ystep, S;
    int tx, ty, nx;
    int code_mul;
};

__device__ __forceinline__ const unsigned char *dense_base(const DenseCtx &c, int wid) {
    return c.tile + (wid / kTileW) * c.row_mul + (wid & (kTileW - 1)) * 4;
}
__device__ __forceinline__ int dense_tile_off(const DenseCtx &c, int y, int x) {   // byte offset of integral (y, x) from a window base
    return 4 * (c.ystep == 1 ? y * c.S + x : y * c.S + (x & 1) * (c.S >> 1) + (x >> 1));
}
__device__ __forceinline__ void dense_write_code(const DenseCtx &c, int wid, int code) {
    const int wx = wid & (kTileW - 1), wy = wid / kTileW;
    c.codes[(size_t)(c.ty * kTileH + wy) * c.nx + c.tx * kTileW + wx] = (int16_t)code;
}

// FP64 sigma of window `wid` (tempcv.cpp:824-832): int32 corners from the tile, uint64 from global
__device__ __forceinline__ double dense_sigma(const DenseParams &P, const DenseCtx &c, int wid) {
    const int wx = wid & (kTileW - 1), wy = wid / kTileW;
    const int eq_w = P.win_w - 2, eq_h = P.win_h - 2;
    const unsigned char *base = c.tile + wy * c.row_mul + wx * 4;
    const int s4 = TILE_LD(base, dense_tile_off(c, 1, 1)) - TILE_LD(base, dense_tile_off(c, 1, 1 + eq_w)) -
                   TILE_LD(base, dense_tile_off(c, 1 + eq_h, 1)) + TILE_LD(base, dense_tile_off(c, 1 + eq_h, 1 + eq_w));
    const ull *q = c.gsq + (size_t)(wy * c.ystep) * c.sq_pitch + wx * c.ystep;
    const int g0 = c.sq_pitch + 1, g1 = g0 + eq_w, g2 = (1 + eq_h) * c.sq_pitch + 1, g3 = g2 + eq_w;
    const ull q4 = __ldg(q + g0) - __ldg(q + g1) - __ldg(q + g2) + __ldg(q + g3);
    return window_sigma(s4, q4, P.inv_area);
}

// Exact evaluation of one parameter-resident stage for one window (the reference's arithmetic).
__device__ __noinline__ bool dense_stage_exact(const DenseParams &P, const DenseCtx &c, int s, int wid) {
    const unsigned char *base = dense_base(c, wid);
    const double sigma = dense_sigma(P, c, wid);
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

// FP32-filtered evaluation of stumps j0, j0 + jstep, ... of stage s for K windows of this thread.
//   FIXED = true : window k lives at base0 + k*ROWSTEP (compile-time) -> immediate offsets
//   FIXED = false: window k lives at base[k]
template <int K, bool DBL, bool HAS3, bool FIXED, int ROWSTEP>
__device__ __forceinline__ void dense_filter_stage(const DenseParams &P, int s, int j0, int jstep,
                                                   const unsigned char *const (&base)[K], const float (&sg)[K],
                                                   double (&S)[K], bool (&near)[K]) {
    const int first = P.stage[s].first, count = P.stage[s].count;
    const float eps = P.filter_eps, eps4 = eps * 0.25f;
#pragma unroll 1
    for (int j = j0; j < count; j += jstep) {
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
__device__ __forceinline__ void dense_dispatch_stage(const DenseParams &P, int s, int j0, int jstep,
                                                     const unsigned char *const (&base)[K], const float (&sg)[K],
                                                     double (&S)[K], bool (&near)[K]) {
    const uint32_t flags = P.stage[s].flags;
    if (flags & 1u) dense_filter_stage<K, true, false, FIXED, ROWSTEP>(P, s, j0, jstep, base, sg, S, near);
    else if (flags & 2u) dense_filter_stage<K, false, true, FIXED, ROWSTEP>(P, s, j0, jstep, base, sg, S, near);
    else dense_filter_stage<K, false, false, FIXED, ROWSTEP>(P, s, j0, jstep, base, sg, S, near);
}

// Phase 2: a warp takes up to 32 surviving windows (lane k holds window k) through ALL the
// remaining stages on its own -- no block barrier, no shared survivor list.  For every stage
// the 32 lanes are arranged as  w window slots x G stump groups  (w = smallest power of two
// >= live windows, G = 32 / w): lane (slot, grp) evaluates stumps grp, grp + G, ... of the
// stage for window `slot`, and the G partial sums of a window meet through xor-shuffles.
// With 32 windows this is thread-per-window; as windows die the lanes they free take over a
// share of the stumps (one window left: 32 stumps per pass), so lanes stay busy without
// re-compaction across warps.  Survivors are re-packed to the low lanes through a 384-byte
// scratch.  Stump records come from global memory (TailStump, 3 x LDG.128 per lane); the
// arithmetic is the reference's exact one (no filter).  Summing partial sums out of order is
// bit-exact because the packer proved the stage's alpha sum exact in any order (flags bit2);
// a stage without the proof keeps w = 32, G = 1, i.e. tree order.
__device__ __forceinline__ void dense_warp_finish(const DenseParams &P, const DenseCtx &c, const CascadeArgs &a,
                                                  const CasLevel &CL, int frame, int cl, int px0, int py0, int s0, int wid,
                                                  double *scr_sigma, int *scr_wid, int lane) {
    const TailStump *__restrict__ tail = P.tail;
    int nw = __popc(__ballot_sync(0xffffffffu, wid >= 0));   // the windows sit in lanes 0 .. nw-1
    double sigma = wid >= 0 ? dense_sigma(P, c, wid) : 1.0;
    int ss = s0;
    while (nw > 0 && ss < P.tail_stages) {
        const DenseStage st = P.stage[ss];
        const bool dbl = st.flags & 1u;
        int lw = 5;
        if (st.flags & 4u) while (lw > 0 && (1 << (lw - 1)) >= nw) lw--;
        const int w = 1 << lw, G = 32 >> lw;
        const int slot = lane & (w - 1), grp = lane >> lw;
        const int swid = __shfl_sync(0xffffffffu, wid, slot);
        const double ssig = __shfl_sync(0xffffffffu, sigma, slot);
        double acc = 0.0;
        if (slot < nw) {
            const unsigned char *base = dense_base(c, swid);
            const uint4 *rec = reinterpret_cast<const uint4 *>(tail + st.tail_first + grp);
            for (int j = grp; j < st.count; j += G, rec += 3 * G) {
                const uint4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
                // q0 = off[0..7]; q1 = off[8..11], w0, w1; q2 = w2, thr, a0, a1
                const int r0 = TILE_LD(base, q0.x & 0xffffu) - TILE_LD(base, q0.x >> 16) - TILE_LD(base, q0.y & 0xffffu) + TILE_LD(base, q0.y >> 16);
                const int r1 = TILE_LD(base, q0.z & 0xffffu) - TILE_LD(base, q0.z >> 16) - TILE_LD(base, q0.w & 0xffffu) + TILE_LD(base, q0.w >> 16);
                const float w0 = __uint_as_float(q1.z), w1 = __uint_as_float(q1.w);
                const double t = __dmul_rn((double)__uint_as_float(q2.y), ssig);
                double sv;
                if (dbl) {  // tempcv.cpp:872-898; both products are exact in double, so fma == mul, mul, add
                    sv = __fma_rn((double)r1, (double)w1, __dmul_rn((double)r0, (double)w0));
                } else {    // tempcv.cpp:899-930 / 782-786: float products, double accumulation
                    sv = __dadd_rn((double)__fmul_rn(__int2float_rn(r0), w0), (double)__fmul_rn(__int2float_rn(r1), w1));
                    if ((q1.y >> 16) != 0) {
                        const int r2 = TILE_LD(base, q1.x & 0xffffu) - TILE_LD(base, q1.x >> 16) - TILE_LD(base, q1.y & 0xffffu) + TILE_LD(base, q1.y >> 16);
                        sv = __dadd_rn(sv, (double)__fmul_rn(__int2float_rn(r2), __uint_as_float(q2.x)));
                    }
                }
                acc = __dadd_rn(acc, (double)__uint_as_float(sv >= t ? q2.w : q2.z));
            }
        }
        for (int d = w; d < 32; d <<= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, d));
        const bool mine = lane < nw;   // then slot == lane: acc is this lane's own window
        const bool pass = acc >= (double)st.thr;
        const unsigned surv = __ballot_sync(0xffffffffu, mine && pass);
        if (mine && !pass && c.codes) dense_write_code(c, wid, ss * c.code_mul);
        ss++;
        const int n2 = __popc(surv);
        if (n2 != nw) {   // re-pack the survivors into lanes 0 .. n2-1
            if (mine && pass) {
                const int d = __popc(surv & ((1u << lane) - 1u));
                scr_wid[d] = wid;
                scr_sigma[d] = sigma;
            }
            __syncwarp();
            nw = n2;
            if (lane < nw) { wid = scr_wid[lane]; sigma = scr_sigma[lane]; } else wid = -1;
            __syncwarp();
        }
    }
    if (nw == 0 || lane >= nw) return;
    const int x = px0 + (wid & (kTileW - 1)) * c.ystep, y = py0 + (wid / kTileW) * c.ystep;
    if (ss >= P.total_stages) {          // passed every stage: a detection
        emit_rect(a, CL, frame, x, y);
        if (c.codes) dense_write_code(c, wid, P.total_stages);
    } else {                             // the rest of the cascade belongs to the deep kernel
        const ull slot = atomicAdd(a.counters + 1, 1ull);
        if (slot < a.queue_cap) {
            QueueItem it;
            it.key = ((uint32_t)frame << 16) | ((uint32_t)cl << 8) | (uint32_t)ss;
            it.xy = ((uint32_t)y << 16) | (uint32_t)x;
            a.queue[slot] = it;
        } else {
            atomicAdd(a.counters + 3, 1ull);
        }
    }
}

// ROWSTEP_T: compile-time byte distance between a thread's consecutive windows in phase 1
// (= 2 window rows), or 0 to use the runtime value (generic window sizes).
template <int ROWSTEP_T>
__global__ void __launch_bounds__(kDenseThreads)
k_cascade_tiles(const __grid_constant__ DenseParams P, const __grid_constant__ CascadeArgs a, const int tile0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DenseSmemPlan plan = dense_smem_plan(P);
    unsigned char *tile = smem_raw + plan.tile;
    float *sgf = reinterpret_cast<float *>(smem_raw + plan.sgf);
    uint16_t *list = reinterpret_cast<uint16_t *>(smem_raw + plan.list);
    unsigned char *scr = smem_raw + plan.scr + (size_t)(threadIdx.x >> 5) * kScrBytesPerWarp;
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
    c.tile = tile; c.sgf = sgf; c.gsq = gsq;
    c.codes = a.codes ? a.codes + (size_t)frame * a.windows_per_frame + CL.win_base : nullptr;
    c.sq_pitch = L.sum_pitch;
    c.row_mul = ystep * S * 4;
    c.ystep = ystep; c.S = S;
    c.tx = tx; c.ty = ty; c.nx = CL.nx;
    c.code_mul = P.is_tree ? 2 : 1;

    // ---- phase 1: sigma, then the fixed-geometry stages ----
    constexpr int kRowsPerSlot = kDenseThreads / kTileW;   // window rows between a thread's slots (2)
    const int wx = tid & (kTileW - 1), wy0 = tid / kTileW;
    uint32_t alive = 0;   // bit k: window (wx, wy0 + 2k) still alive
#pragma unroll
    for (int k = 0; k < kDenseSlots; k++) {
        const int wy = wy0 + k * kRowsPerSlot;
        const int wid = wy * kTileW + wx;
        if (wx < n_wx && wy < n_wy) {
            alive |= 1u << k;
            sgf[wid] = (float)dense_sigma(P, c, wid);
        } else {
            sgf[wid] = 1.0f;
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
                sg[k] = sgf[wy * kTileW + wx];
                Ssum[k] = 0.0;
                near[k] = P.force_exact != 0;
            }
            if (ROWSTEP_T) dense_dispatch_stage<kDenseChunk, true, ROWSTEP_T>(P, s, 0, 1, base, sg, Ssum, near);
            else dense_dispatch_stage<kDenseChunk, false, 0>(P, s, 0, 1, base, sg, Ssum, near);
            const double sthr = (double)P.stage[s].thr;
#pragma unroll
            for (int k = 0; k < kDenseChunk; k++) {
                if (!((m4 >> k) & 1u)) continue;
                const int wid = (wy0 + (k0 + k) * kRowsPerSlot) * kTileW + wx;
                bool pass = Ssum[k] >= sthr;
                if (near[k]) pass = dense_stage_exact(P, c, s, wid);
                if (!pass) {
                    alive &= ~(1u << (k0 + k));
                    if (c.codes) dense_write_code(c, wid, s * c.code_mul);
                }
            }
        }
    }

    // ---- phase-1 survivors -> one list (warp-aggregated append; a warp's entries keep
    //      consecutive wx, i.e. distinct banks) ----
#pragma unroll
    for (int k = 0; k < kDenseSlots; k++) {
        const bool al = (alive >> k) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, al);
        if (m) {
            int b = 0;
            if (lane == 0) b = atomicAdd(ctl + kCtlAlive, __popc(m));
            b = __shfl_sync(0xffffffffu, b, 0);
            if (al) list[b + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((wy0 + k * kRowsPerSlot) * kTileW + wx);
        }
    }
    __syncthreads();
    const int n_alive = ctl[kCtlAlive];
    if (n_alive == 0) return;

    // ---- phase 2: every warp finishes an equal share of the survivors on its own ----
    const int share = (n_alive + kDenseWarps - 1) / kDenseWarps;
    const int lo = warp * share, hi = min(n_alive, lo + share);
    double *scr_sigma = reinterpret_cast<double *>(scr);
    int *scr_wid = reinterpret_cast<int *>(scr + 32 * sizeof(double));
    for (int b0 = lo; b0 < hi; b0 += 32) {
        const int wid = b0 + lane < hi ? (int)list[b0 + lane] : -1;
        dense_warp_finish(P, c, a, CL, frame, cl, px0, py0, s, wid, scr_sigma, scr_wid, lane);
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

// row steps of the stock window sizes (bytes between a thread's consecutive phase-1 windows)
constexpr int dense_stride_ce(int win_w, int ystep) {
    int cols = ((kTileW - 1) * ystep + win_w + 1 + 3) & ~3;
    int s = ystep == 1 ? cols : 2 * ((cols + 1) / 2);
    s = (s + 3) & ~3;
    while ((ystep * s) % 32 != 8) s += 4;
    return s;
}
constexpr int dense_rowstep_ce(int win_w, int ystep) { return (kDenseThreads / kTileW) * ystep * dense_stride_ce(win_w, ystep) * 4; }

cudaError_t launch_cascade_tiles(const DenseParams &P, const CascadeArgs &a, int tile0, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0 || a.n_frames == 0) return cudaSuccess;
    const size_t smem = dense_smem_bytes(P);
    // byte distance between a thread's consecutive phase-1 windows: 2 window rows
    const int rowstep = (kDenseThreads / kTileW) * P.ystep * P.tile_stride * 4;
    // the common window widths (20 and 24 pixels) get immediate-offset code
    constexpr int r20_1 = dense_rowstep_ce(20, 1), r20_2 = dense_rowstep_ce(20, 2);
    constexpr int r24_1 = dense_rowstep_ce(24, 1), r24_2 = dense_rowstep_ce(24, 2);
    static_assert(r20_1 == r24_1 && r20_2 != r24_2 && r20_2 != r20_1 && r24_2 != r20_1, "row steps must be distinct switch labels");
    switch (rowstep) {
        case r20_1: return launch_tiles_t<r20_1>(P, a, tile0, n_tiles, smem, stream);
        case r20_2: return launch_tiles_t<r20_2>(P, a, tile0, n_tiles, smem, stream);
        case r24_2: return launch_tiles_t<r24_2>(P, a, tile0, n_tiles, smem, stream);
        default:    return launch_tiles_t<0>(P, a, tile0, n_tiles, smem, stream);
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
