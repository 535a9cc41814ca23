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
// ------------------------------------------------------------------------------------
struct DenseSmem {
    // [tile ints][sigma doubles][list0][list1][ctl]
    static __host__ __device__ size_t tile_bytes(const DenseParams &P) {
        return (size_t)((kTileH - 1) * 2 + P.win_h + 1) * P.tile_stride * 4;
    }
};

size_t dense_smem_bytes(const DenseParams &P) {
    size_t tile = (DenseSmem::tile_bytes(P) + 127) & ~(size_t)127;
    return tile + kTileWindows * sizeof(double) + 2 * kTileWindows * sizeof(uint16_t) + 64;
}

#define TILE_LD(base, off) (*reinterpret_cast<const int *>((base) + (off)))

__device__ __forceinline__ bool dense_eval_stage(const DenseParams &P, int s, const unsigned char *base, double sigma) {
    const int first = P.stage[s].first, count = P.stage[s].count;
    const uint32_t flags = P.stage[s].flags;
    double S = 0.0;
    if (flags & 1u) {  // two_rects stage of a stump cascade: double products (tempcv.cpp:872-898)
        for (int j = 0; j < count; j++) {
            const DenseStump &q = P.stump[first + j];
            const int r0 = TILE_LD(base, q.off[0]) - TILE_LD(base, q.off[1]) - TILE_LD(base, q.off[2]) + TILE_LD(base, q.off[3]);
            const int r1 = TILE_LD(base, q.off[4]) - TILE_LD(base, q.off[5]) - TILE_LD(base, q.off[6]) + TILE_LD(base, q.off[7]);
            const double t = __dmul_rn((double)q.thr, sigma);
            // both products are exact in double (|r| < 2^24, 24-bit weights) => fma == mul, mul, add
            const double sum = __fma_rn((double)r1, (double)q.w[1], __dmul_rn((double)r0, (double)q.w[0]));
            S = __dadd_rn(S, (double)(sum >= t ? q.a1 : q.a0));
        }
    } else {  // float products, double accumulation (tempcv.cpp:899-930, 782-786)
        for (int j = 0; j < count; j++) {
            const DenseStump &q = P.stump[first + j];
            const int r0 = TILE_LD(base, q.off[0]) - TILE_LD(base, q.off[1]) - TILE_LD(base, q.off[2]) + TILE_LD(base, q.off[3]);
            const int r1 = TILE_LD(base, q.off[4]) - TILE_LD(base, q.off[5]) - TILE_LD(base, q.off[6]) + TILE_LD(base, q.off[7]);
            const double t = __dmul_rn((double)q.thr, sigma);
            double sum = __dadd_rn((double)__fmul_rn(__int2float_rn(r0), q.w[0]), (double)__fmul_rn(__int2float_rn(r1), q.w[1]));
            if (q.off[11] != 0) {  // warp-uniform: third rectangle present
                const int r2 = TILE_LD(base, q.off[8]) - TILE_LD(base, q.off[9]) - TILE_LD(base, q.off[10]) + TILE_LD(base, q.off[11]);
                sum = __dadd_rn(sum, (double)__fmul_rn(__int2float_rn(r2), q.w[2]));
            }
            S = __dadd_rn(S, (double)(sum >= t ? q.a1 : q.a0));
        }
    }
    return S >= (double)P.stage[s].thr;
}

__global__ void __launch_bounds__(kDenseThreads)
k_cascade_tiles(const __grid_constant__ DenseParams P, const __grid_constant__ CascadeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t tile_sz = (DenseSmem::tile_bytes(P) + 127) & ~(size_t)127;
    unsigned char *tile = smem_raw;
    double *sigma = reinterpret_cast<double *>(smem_raw + tile_sz);
    uint16_t *list0 = reinterpret_cast<uint16_t *>(sigma + kTileWindows);
    uint16_t *list1 = list0 + kTileWindows;
    int *ctl = reinterpret_cast<int *>(list1 + kTileWindows);   // [0],[1] list counters, [2] queue base
    uint64_t *bar = reinterpret_cast<uint64_t *>(ctl + 4);

    const int tid = threadIdx.x, lane = tid & 31;
    const int frame = blockIdx.y;

    // which level does this tile belong to?
    int cl = 0;
    {
        int lo = 0, hi = a.n_cas_levels - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(&a.cas_levels[mid].tile_base) <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        cl = lo;
    }
    const CasLevel CL = a.cas_levels[cl];
    const PyrLevel L = a.levels[CL.pyr_level];
    const int local = blockIdx.x - CL.tile_base;
    const int tx = local % CL.tiles_x, ty = local / CL.tiles_x;
    const int ystep = CL.ystep;
    const int px0 = tx * kTileW * ystep, py0 = ty * kTileH * ystep;   // tile origin in the integral image
    const int n_wx = min(kTileW, CL.nx - tx * kTileW), n_wy = min(kTileH, CL.ny - ty * kTileH);
    const int rows = min((kTileH - 1) * ystep + P.win_h + 1, L.h + 1 - py0);
    const int cols = ((kTileW - 1) * ystep + P.win_w + 1 + 3) & ~3;
    const int stride_b = P.tile_stride * 4;

    const size_t frame_off = (size_t)frame * a.sum_frame_stride + L.sum_off;
    const int32_t *__restrict__ gsum = a.sum + frame_off + (size_t)py0 * L.sum_pitch + px0;
    const ull *__restrict__ gsq = a.sq + frame_off + (size_t)py0 * L.sum_pitch + px0;

    // ---- stage the integral tile: one TMA bulk copy per row ----
    if (tid == 0) {
        ctl[0] = 0; ctl[1] = 0; ctl[2] = 0;
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)(rows * cols * 4));
    for (int r = tid; r < rows; r += kDenseThreads)
        tma_bulk_g2s(tile + (size_t)r * stride_b, gsum + (size_t)r * L.sum_pitch, (uint32_t)(cols * 4), bar);
    mbar_wait(bar, 0);

    const int eq_w = P.win_w - 2, eq_h = P.win_h - 2;
    const int e0 = (1 * P.tile_stride + 1) * 4, e1 = e0 + eq_w * 4;
    const int e2 = ((1 + eq_h) * P.tile_stride + 1) * 4, e3 = e2 + eq_w * 4;
    const int g0 = L.sum_pitch + 1, g1 = g0 + eq_w, g2 = (1 + eq_h) * L.sum_pitch + 1, g3 = g2 + eq_w;
    const int code_mul = P.pad[0] ? 2 : 1;   // stage-tree cascades report 2*last_stage (+accept)
    int16_t *__restrict__ codes = a.codes ? a.codes + (size_t)frame * a.windows_per_frame + CL.win_base : nullptr;

    // ---- pass A: sigma + stage 0 for every window of the tile ----
    for (int w0 = 0; w0 < kTileWindows; w0 += kDenseThreads) {
        const int wid = w0 + tid;
        const int wx = wid & (kTileW - 1), wy = wid / kTileW;
        const bool valid = wx < n_wx && wy < n_wy;
        bool pass = false;
        if (valid) {
            const unsigned char *base = tile + ((size_t)(wy * ystep) * P.tile_stride + wx * ystep) * 4;
            const int s4 = TILE_LD(base, e0) - TILE_LD(base, e1) - TILE_LD(base, e2) + TILE_LD(base, e3);
            const ull *q = gsq + (size_t)(wy * ystep) * L.sum_pitch + wx * ystep;
            const ull q4 = __ldg(q + g0) - __ldg(q + g1) - __ldg(q + g2) + __ldg(q + g3);
            const double sg = window_sigma(s4, q4, P.inv_area);
            sigma[wid] = sg;
            pass = dense_eval_stage(P, 0, base, sg);
            if (!pass && codes) codes[(size_t)(ty * kTileH + wy) * CL.nx + tx * kTileW + wx] = 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&ctl[0], __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (pass) list0[wbase + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)wid;
        }
    }
    __syncthreads();

    // ---- remaining dense stages, compacting after each ----
    uint16_t *lin = list0, *lout = list1;
    int cur = 0;            // ctl index of the input list counter
    int s = 1;
    int n = ctl[0];
    while (n > 0 && s < P.n_stages && !(n <= kHandoffWindows && P.n_stages < P.total_stages)) {
        if (tid == 0) ctl[cur ^ 1] = 0;
        __syncthreads();
        for (int i0 = (tid & ~31); i0 < n; i0 += kDenseThreads) {
            const int i = i0 + lane;
            bool pass = false;
            int wid = 0;
            if (i < n) {
                wid = lin[i];
                const int wx = wid & (kTileW - 1), wy = wid / kTileW;
                const unsigned char *base = tile + ((size_t)(wy * ystep) * P.tile_stride + wx * ystep) * 4;
                pass = dense_eval_stage(P, s, base, sigma[wid]);
                if (!pass && codes)
                    codes[(size_t)(ty * kTileH + wy) * CL.nx + tx * kTileW + wx] = (int16_t)(s * code_mul);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&ctl[cur ^ 1], __popc(bal));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (pass) lout[wbase + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)wid;
            }
        }
        __syncthreads();
        cur ^= 1;
        n = ctl[cur];
        uint16_t *tmp = lin; lin = lout; lout = tmp;
        s++;
    }
    if (n == 0) return;

    // ---- survivors: accepted (whole cascade was dense) or handed to the deep kernel ----
    if (s >= P.total_stages) {
        for (int i = tid; i < n; i += kDenseThreads) {
            const int wid = lin[i];
            const int wx = wid & (kTileW - 1), wy = wid / kTileW;
            emit_rect(a, CL, frame, px0 + wx * ystep, py0 + wy * ystep);
            if (codes) codes[(size_t)(ty * kTileH + wy) * CL.nx + tx * kTileW + wx] = (int16_t)P.total_stages;
        }
    } else {
        if (tid == 0) {
            const ull b = atomicAdd(a.counters + 1, (ull)n);
            reinterpret_cast<ull *>(ctl)[1] = b;   // ctl[2..3]
        }
        __syncthreads();
        const ull qb = reinterpret_cast<ull *>(ctl)[1];
        for (int i = tid; i < n; i += kDenseThreads) {
            const int wid = lin[i];
            const int wx = wid & (kTileW - 1), wy = wid / kTileW;
            if (qb + i < a.queue_cap) {
                QueueItem it;
                it.key = ((uint32_t)frame << 16) | ((uint32_t)cl << 8) | (uint32_t)s;
                it.xy = ((uint32_t)(py0 + wy * ystep) << 16) | (uint32_t)(px0 + wx * ystep);
                a.queue[qb + i] = it;
            } else {
                atomicAdd(a.counters + 3, 1ull);
            }
        }
    }
}

cudaError_t launch_cascade_tiles(const DenseParams &P, const CascadeArgs &a, cudaStream_t stream) {
    if (a.n_tiles == 0 || a.n_frames == 0) return cudaSuccess;
    const size_t smem = dense_smem_bytes(P);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_cascade_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    k_cascade_tiles<<<dim3(a.n_tiles, a.n_frames), kDenseThreads, smem, stream>>>(P, a);
    return cudaGetLastError();
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
