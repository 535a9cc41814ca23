// kernels_clif.cu -- pyramid downscale + integral / squared-integral / tilted-integral
// images for sm_100a.  Replaces, from scratch, the reference's
//   * cvResize(img, level, CV_INTER_LINEAR) per pyramid level       (tempcv.cpp:1301)
//   * integralImageSumRows / integralImageSumCols                    (clif.cl:79-120)
//     and cvIntegral(sum, sqsum, tilted)                             (tempcv.cpp:1302)
//   * bgrToGrayscale                                                 (clif.cl:4-18)
//
// Design (HBM-bound byte work; no tensor cores):
//   K1 resize_colsum : all levels x frames x 32-row blocks in one launch.  Each thread makes
//                      4 adjacent level pixels per row (11-bit fixed-point bilinear, bit
//                      exact with OpenCV) and keeps their column sums / column square sums
//                      for the row block -> the vertical carries cost no extra pass.
//   K2 colscan       : exclusive prefix of those carries over row blocks (tiny).
//   K3 integral_rows : one CTA per (level, frame, 32-row block).  A thread owns 8 adjacent
//                      columns: running column sums in registers (seeded with the carry),
//                      thread-serial + warp-shuffle + cross-warp prefix along the row, and
//                      16-byte vector stores of int32 sums and uint64 square sums.  The image
//                      is read once (8 B loads) and every output byte is written once.
//   K4 tilted        : row-sequential diagonal recurrences, only for cascades with tilted
//                      features (fullbody & co).
#include <cstdint>

#include "kernels.h"

namespace clfd {

typedef unsigned long long ull;

// ------------------------------------------------------------------------------------
// K1: pyramid level pixels + per-row-block column sums
// ------------------------------------------------------------------------------------
constexpr int kResizeThreads = 128;  // x 4 px = 512 columns per CTA
constexpr int kResizeCols = kResizeThreads * 4;

__global__ void __launch_bounds__(kResizeThreads) k_resize_colsum(const PyramidArgs a) {
    const int4 it = a.resize_items[blockIdx.x];
    const PyrLevel L = a.levels[it.x];
    const int frame = blockIdx.y;
    const int x0 = it.z * kResizeCols + threadIdx.x * 4;
    if (x0 >= max(L.pyr_pitch, L.sum_pitch)) return;

    const uint8_t *__restrict__ src = a.frames + (size_t)frame * a.frame_stride;
    uint8_t *__restrict__ dst = a.pyr + (size_t)frame * a.pyr_frame_stride + L.pyr_off;

    int sx0[4], sx1[4], a0[4], a1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + i;
        if (x < L.w) {
            const int s = __ldg(a.xofs + L.xtab_off + x);
            const short2 c = __ldg(a.xalpha + L.xtab_off + x);
            sx0[i] = s; sx1[i] = min(s + 1, a.W - 1); a0[i] = c.x; a1[i] = c.y;
        } else {
            sx0[i] = sx1[i] = 0; a0[i] = a1[i] = 0;
        }
    }
    uint32_t cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
    int prev_r1[4] = {0, 0, 0, 0}, prev_sy1 = -1;
    // whole groups of 4 pixels of an unscaled level with 4-byte aligned source rows
    const bool copy4 = L.w == a.W && L.h == a.H && x0 + 4 <= L.w && (a.row_stride & 3) == 0 && (a.frame_stride & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.frames) & 3) == 0;
    const int y0 = it.y * kRowBlock, y1 = min(y0 + kRowBlock, L.h);
    for (int y = y0; y < y1; y++) {
        const int sy = __ldg(a.yofs + L.ytab_off + y);
        const short2 b = __ldg(a.ybeta + L.ytab_off + y);
        const int sy0 = min(max(sy, 0), a.H - 1), sy1 = min(max(sy + 1, 0), a.H - 1);
        const uint8_t *__restrict__ S0 = src + (size_t)sy0 * a.row_stride;
        const uint8_t *__restrict__ S1 = src + (size_t)sy1 * a.row_stride;
        uint32_t packed = 0;
        if (copy4) {   // the level of factor 1: the resize is the identity (coefficients 2048 / 0), 4 pixels per load
            packed = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)y * a.row_stride + x0));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t u = (packed >> (8 * i)) & 255u;
                cs[i] += u;
                cq[i] += u * u;
            }
            *reinterpret_cast<uint32_t *>(dst + (size_t)y * L.pyr_pitch + x0) = packed;
            continue;
        }
        // the horizontal pass of a source row is kept for the next output row: at factors < 2 most rows'
        // upper source row is the previous row's lower one (block-uniform test, no divergence)
        const bool reuse = sy0 == prev_sy1;
        prev_sy1 = sy1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r0 = reuse ? prev_r1[i] : (int)__ldg(S0 + sx0[i]) * a0[i] + (int)__ldg(S0 + sx1[i]) * a1[i];
            const int r1 = (int)__ldg(S1 + sx0[i]) * a0[i] + (int)__ldg(S1 + sx1[i]) * a1[i];
            prev_r1[i] = r1;
            const int v = ((((int)b.x * (r0 >> 4)) >> 16) + (((int)b.y * (r1 >> 4)) >> 16) + 2) >> 2;
            const uint32_t u = (uint32_t)v & 255u;
            packed |= u << (8 * i);
            cs[i] += u;
            cq[i] += u * u;
        }
        if (x0 < L.pyr_pitch) *reinterpret_cast<uint32_t *>(dst + (size_t)y * L.pyr_pitch + x0) = packed;
    }
    if (x0 < L.sum_pitch) {
        uint32_t *c0 = a.col + (size_t)frame * a.col_frame_stride + L.col_off + (size_t)it.y * L.sum_pitch + x0;
        *reinterpret_cast<uint4 *>(c0) = make_uint4(cs[0], cs[1], cs[2], cs[3]);
        *reinterpret_cast<uint4 *>(c0 + a.col_plane_stride) = make_uint4(cq[0], cq[1], cq[2], cq[3]);
    }
}

cudaError_t launch_resize_colsum(const PyramidArgs &a, cudaStream_t stream) {
    if (a.n_resize_items == 0 || a.n_frames == 0) return cudaSuccess;
    k_resize_colsum<<<dim3(a.n_resize_items, a.n_frames), kResizeThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K2: exclusive prefix over row blocks of the column sums (in place)
// ------------------------------------------------------------------------------------
constexpr int kColscanThreads = 128;

__global__ void __launch_bounds__(kColscanThreads) k_colscan(const PyramidArgs a) {
    const int4 it = a.colscan_items[blockIdx.x];
    const PyrLevel L = a.levels[it.x];
    const int X = (it.y * kColscanThreads + threadIdx.x) * 4;
    if (X >= L.sum_pitch) return;
#pragma unroll
    for (int plane = 0; plane < 2; plane++) {
        uint32_t *p = a.col + (size_t)blockIdx.y * a.col_frame_stride + plane * a.col_plane_stride + L.col_off + X;
        uint4 acc = make_uint4(0, 0, 0, 0);
        // eight row blocks per step: all loads first (independent, in flight together), then the stores --
        // a load behind a store to the same array is otherwise serialised, one DRAM latency per row block
        constexpr int kAhead = 8;
        for (int rb0 = 0; rb0 < L.nrb; rb0 += kAhead) {
            uint4 v[kAhead];
#pragma unroll
            for (int k = 0; k < kAhead; k++)
                if (rb0 + k < L.nrb) v[k] = *reinterpret_cast<const uint4 *>(p + (size_t)(rb0 + k) * L.sum_pitch);
#pragma unroll
            for (int k = 0; k < kAhead; k++) {
                if (rb0 + k < L.nrb) {
                    *reinterpret_cast<uint4 *>(p + (size_t)(rb0 + k) * L.sum_pitch) = acc;
                    acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
                }
            }
        }
    }
}

cudaError_t launch_colscan(const PyramidArgs &a, cudaStream_t stream) {
    if (a.n_colscan_items == 0 || a.n_frames == 0) return cudaSuccess;
    k_colscan<<<dim3(a.n_colscan_items, a.n_frames), kColscanThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K3: integral rows.  sum[Y][X] = sum_{y<Y, x<X} I, written for Y in the row block.
// ------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_integral_rows(const PyramidArgs a, const int4 *__restrict__ items) {
    constexpr int NW = NT / 32;
    __shared__ uint32_t wtot_s[2][NW];
    __shared__ ull wtot_q[2][NW];

    const int4 it = items[blockIdx.x];
    const PyrLevel L = a.levels[it.x];
    const int rb = it.y, frame = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int X0 = t * 8;
    const bool in_sum = X0 < L.sum_pitch;
    const bool in_pyr = X0 < L.pyr_pitch;

    const uint8_t *__restrict__ pyr = a.pyr + (size_t)frame * a.pyr_frame_stride + L.pyr_off;
    int32_t *__restrict__ sum = a.sum + (size_t)frame * a.sum_frame_stride + L.sum_off;
    ull *__restrict__ sq = a.sq + (size_t)frame * a.sum_frame_stride + L.sum_off;

    uint32_t ca[8], cq[8];
    if (in_sum) {
        const uint32_t *c0 = a.col + (size_t)frame * a.col_frame_stride + L.col_off + (size_t)rb * L.sum_pitch + X0;
        const uint4 s0 = *reinterpret_cast<const uint4 *>(c0), s1 = *reinterpret_cast<const uint4 *>(c0 + 4);
        const uint4 q0 = *reinterpret_cast<const uint4 *>(c0 + a.col_plane_stride);
        const uint4 q1 = *reinterpret_cast<const uint4 *>(c0 + a.col_plane_stride + 4);
        ca[0] = s0.x; ca[1] = s0.y; ca[2] = s0.z; ca[3] = s0.w; ca[4] = s1.x; ca[5] = s1.y; ca[6] = s1.z; ca[7] = s1.w;
        cq[0] = q0.x; cq[1] = q0.y; cq[2] = q0.z; cq[3] = q0.w; cq[4] = q1.x; cq[5] = q1.y; cq[6] = q1.z; cq[7] = q1.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) ca[i] = cq[i] = 0;
    }

    const int y0 = rb * kRowBlock, y1 = min(y0 + kRowBlock, L.h);
    if (rb == 0 && in_sum) {  // row 0 of every integral image is zero
        int4 *s = reinterpret_cast<int4 *>(sum + X0);
        s[0] = make_int4(0, 0, 0, 0); s[1] = make_int4(0, 0, 0, 0);
        ulonglong2 *q = reinterpret_cast<ulonglong2 *>(sq + X0);
        q[0] = q[1] = q[2] = q[3] = make_ulonglong2(0, 0);
    }

    // cross-warp exclusive prefix of per-warp totals (lane 31 holds the warp's inclusive total)
    auto left_of_warp = [&](int buf, uint32_t tot_s, ull tot_q, uint32_t &os, ull &oq) {
        os = 0; oq = 0;
        if (NW > 1) {
            if (lane == 31) { wtot_s[buf][warp] = tot_s; wtot_q[buf][warp] = tot_q; }
            __syncthreads();
            if (NW <= 8) {
                for (int w = 0; w < warp; w++) { os += wtot_s[buf][w]; oq += wtot_q[buf][w]; }
            } else {
                uint32_t ws = lane < warp ? wtot_s[buf][lane] : 0u;
                ull wq = lane < warp ? wtot_q[buf][lane] : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    ws += __shfl_xor_sync(0xffffffffu, ws, d);
                    wq += __shfl_xor_sync(0xffffffffu, wq, d);
                }
                os = ws; oq = wq;
            }
        }
    };

    // Integral row y0 of this thread's 8 columns = exclusive horizontal prefix of the column carries
    // (64-bit scan, ONCE per row block).  After that every row only adds its own exclusive pixel
    // prefix -- out[y+1][x] = out[y][x] + sum_{x' < x} p[y][x'] -- whose scans fit 32 bits: a row of
    // 8-bit pixels sums to < 2^21 and its squares to < 2^30 up to 16384 columns.
    uint32_t Is[8];
    ull Iq[8];
    {
        uint32_t ts = 0;
        ull tq = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { ts += ca[i]; tq += cq[i]; }
        uint32_t is = ts;
        ull iq = tq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t vs = __shfl_up_sync(0xffffffffu, is, d);
            const ull vq = __shfl_up_sync(0xffffffffu, iq, d);
            if (lane >= d) { is += vs; iq += vq; }
        }
        uint32_t os; ull oq;
        left_of_warp(0, is, iq, os, oq);
        uint32_t es = os + is - ts;
        ull eq = oq + iq - tq;
#pragma unroll
        for (int i = 0; i < 8; i++) { Is[i] = es; Iq[i] = eq; es += ca[i]; eq += cq[i]; }
    }

    uint2 nxt = make_uint2(0, 0);
    if (in_pyr) nxt = __ldg(reinterpret_cast<const uint2 *>(pyr + (size_t)y0 * L.pyr_pitch + X0));
    int buf = 1;
    for (int y = y0; y < y1; y++) {
        const uint2 cur = nxt;
        if (in_pyr && y + 1 < y1) nxt = __ldg(reinterpret_cast<const uint2 *>(pyr + (size_t)(y + 1) * L.pyr_pitch + X0));
        uint32_t p[8], ts = 0, tq = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t word = i < 4 ? cur.x : cur.y;
            p[i] = (word >> (8 * (i & 3))) & 255u;
            ts += p[i];
            tq += p[i] * p[i];
        }
        uint32_t is = ts, iq = tq;  // warp inclusive scan of the thread totals, 32 bits each
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t vs = __shfl_up_sync(0xffffffffu, is, d);
            const uint32_t vq = __shfl_up_sync(0xffffffffu, iq, d);
            if (lane >= d) { is += vs; iq += vq; }
        }
        uint32_t os; ull oq;
        left_of_warp(buf, is, (ull)iq, os, oq);
        buf ^= 1;
        if (in_sum) {
            uint32_t rs = os + is - ts;              // exclusive pixel prefix of this row at column X0
            uint32_t rq = (uint32_t)oq + iq - tq;
#pragma unroll
            for (int i = 0; i < 8; i++) { Is[i] += rs; Iq[i] += rq; rs += p[i]; rq += p[i] * p[i]; }
            int4 *s = reinterpret_cast<int4 *>(sum + (size_t)(y + 1) * L.sum_pitch + X0);
            s[0] = make_int4((int)Is[0], (int)Is[1], (int)Is[2], (int)Is[3]);
            s[1] = make_int4((int)Is[4], (int)Is[5], (int)Is[6], (int)Is[7]);
            ulonglong2 *q = reinterpret_cast<ulonglong2 *>(sq + (size_t)(y + 1) * L.sum_pitch + X0);
            q[0] = make_ulonglong2(Iq[0], Iq[1]); q[1] = make_ulonglong2(Iq[2], Iq[3]);
            q[2] = make_ulonglong2(Iq[4], Iq[5]); q[3] = make_ulonglong2(Iq[6], Iq[7]);
        }
    }
}

cudaError_t launch_integral_rows(const PyramidArgs &a, cudaStream_t stream, int *n_launches) {
    int n = 0;
#define CLFD_LAUNCH_INT(K, NT)                                                                          \
    if (a.n_integral_items[K] > 0 && a.n_frames > 0) {                                                  \
        k_integral_rows<NT><<<dim3(a.n_integral_items[K], a.n_frames), NT, 0, stream>>>(a, a.integral_items[K]); \
        n++;                                                                                            \
    }
    CLFD_LAUNCH_INT(0, 32)
    CLFD_LAUNCH_INT(1, 64)
    CLFD_LAUNCH_INT(2, 128)
    CLFD_LAUNCH_INT(3, 256)
    CLFD_LAUNCH_INT(4, 512)
    CLFD_LAUNCH_INT(5, 1024)
#undef CLFD_LAUNCH_INT
    if (n_launches) *n_launches = n;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K4: tilted integral.  tilted[Y][X] = sum_{y<Y, |x-(X-1)| <= (Y-1)-y} I[y][x]
//     T(Y,X) = T(Y-1,X) + A(Y-1,X-1) + B(Y-1,X-1) - I(Y-1,X-1), A/B = up-left / up-right
//     diagonal prefix sums; X = 0 only picks up B(Y-2,0).
// ------------------------------------------------------------------------------------
constexpr int kTiltedThreads = 1024;
constexpr int kTiltedMaxCols = 8;  // columns per thread -> level width <= 8192

__global__ void __launch_bounds__(kTiltedThreads) k_tilted(const PyramidArgs a) {
    extern __shared__ int32_t sm_t[];
    const int4 it = a.tilted_items[blockIdx.x];
    const PyrLevel L = a.levels[it.x];
    const int frame = blockIdx.y, t = threadIdx.x;
    const int wpad = L.w + 2;
    int32_t *A0 = sm_t, *A1 = sm_t + wpad, *B0 = sm_t + 2 * wpad, *B1 = sm_t + 3 * wpad;
    for (int i = t; i < 4 * wpad; i += kTiltedThreads) sm_t[i] = 0;
    const uint8_t *__restrict__ pyr = a.pyr + (size_t)frame * a.pyr_frame_stride + L.pyr_off;
    int32_t *__restrict__ til = a.tilted + (size_t)frame * a.sum_frame_stride + L.sum_off;
    for (int X = t; X < L.sum_pitch; X += kTiltedThreads) til[X] = 0;
    int32_t acc[kTiltedMaxCols];
#pragma unroll
    for (int k = 0; k < kTiltedMaxCols; k++) acc[k] = 0;
    int32_t acc0 = 0;  // column X = 0 (thread 0)
    __syncthreads();
    for (int Y = 1; Y <= L.h; Y++) {
        const uint8_t *__restrict__ p = pyr + (size_t)(Y - 1) * L.pyr_pitch;
        int32_t *__restrict__ out = til + (size_t)Y * L.sum_pitch;
        if (t == 0) { acc0 += B0[1]; out[0] = acc0; }
#pragma unroll
        for (int k = 0; k < kTiltedMaxCols; k++) {
            const int x = t + k * kTiltedThreads;
            if (x < L.w) {
                const int32_t pix = p[x];
                const int32_t a1 = pix + A0[x];      // A(y,x) = I + A(y-1,x-1)   (arrays are x+1 based)
                const int32_t b1 = pix + B0[x + 2];  // B(y,x) = I + B(y-1,x+1)
                A1[x + 1] = a1; B1[x + 1] = b1;
                acc[k] += a1 + b1 - pix;
                out[x + 1] = acc[k];
            }
        }
        __syncthreads();
        int32_t *sw;
        sw = A0; A0 = A1; A1 = sw;
        sw = B0; B0 = B1; B1 = sw;
    }
}

cudaError_t launch_tilted(const PyramidArgs &a, cudaStream_t stream) {
    if (a.n_tilted_items == 0 || a.n_frames == 0 || !a.tilted) return cudaSuccess;
    if (a.max_level_w > kTiltedThreads * kTiltedMaxCols) return cudaErrorInvalidValue;
    const size_t smem = (size_t)4 * (a.max_level_w + 2) * sizeof(int32_t);
    static SmemLimitCache limit;
    if (smem > 48 * 1024) {
        if (cudaError_t e = limit.ensure(k_tilted, smem)) return e;
    }
    k_tilted<<<dim3(a.n_tilted_items, a.n_frames), kTiltedThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// BGR(A) -> gray, OpenCV fixed point: (B*1868 + G*9617 + R*4899 + 8192) >> 14
// ------------------------------------------------------------------------------------
__global__ void k_bgr_to_gray(const uint8_t *__restrict__ bgr, int w, int h, int stride, int channels,
                              uint8_t *__restrict__ gray, int gstride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = bgr + (size_t)y * stride + (size_t)x * channels;
    gray[(size_t)y * gstride + x] = (uint8_t)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + 8192) >> 14);
}

cudaError_t launch_bgr_to_gray(const uint8_t *bgr, int w, int h, int stride, int channels,
                               uint8_t *gray, int gstride, cudaStream_t stream) {
    if (w <= 0 || h <= 0) return cudaSuccess;
    k_bgr_to_gray<<<dim3((w + 255) / 256, h), 256, 0, stream>>>(bgr, w, h, stride, channels, gray, gstride);
    return cudaGetLastError();
}

}  // namespace clfd
