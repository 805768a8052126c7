// gs_data.cu -- the reference's TRAINING DATA PIPELINE as two small kernels + ONE fused kernel per sample (sm_100a).
//
// replaces (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:60-75; [EXT] mmseg 0.x transforms over mmcv / OpenCV):
//   Resize(ratio 0.5-2.0, keep_ratio) -> RandomCrop(512x1024, cat_max_ratio 0.75) -> RandomFlip -> PhotoMetricDistortion ->
//   Normalize(to_rgb) -> Pad(0 / 255) -> DefaultFormatBundle
// The reference runs these on DataLoader worker processes (2 per GPU) and materialises a resized copy of every
// 1024x2048 image (up to 2048x4096x3 bytes) plus two HSV round trips.  Here the decoded uint8 image and label map are the
// only inputs in HBM; every OUTPUT pixel is computed from its (at most) 2x2 source pixels in registers and written once
// as fp32 NCHW (the image conv's input) / int64 labels -- algorithmic traffic ~ 4 source bytes + 12 + 8 output bytes per
// output pixel.  All random decisions are made on the host up front (counter-based stream, data_pipeline.py); the only
// data-dependent one -- RandomCrop's "no class covers >= 75 % of the crop" re-draw loop -- is resolved ON THE DEVICE from
// class histograms of the 11 pre-drawn candidate boxes, so there is no host round trip.
// The 8-bit arithmetic (fixed-point bilinear resize, nearest label resize, integer BGR->HSV, fp32 HSV->BGR, truncating
// convert) is OpenCV's (the CPU restatement used by the tests is pinned against cv2); fp32 operations use explicit
// round-to-nearest intrinsics (no FMA contraction), so image values are bit-identical to the numpy restatement.
#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"

#include <math.h>
#include <string.h>

namespace gs {

constexpr int kAugCandidates = GS_AUG_CANDIDATES;

struct AugDev {            // gs_aug_params + derived doubles, passed by value
    gs_aug_params p;
    double scale_x, scale_y;   // 1 / (new / src), as cv::resize computes it
    int area2x;                // exact 2x decimation: cv::resize's "area fast" path
};

__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

// source index of destination index d for INTER_NEAREST
__device__ __forceinline__ int nearest_src(int d, double ifx, int src) {
    const int s = static_cast<int>(floor(static_cast<double>(d) * ifx));
    return s < src - 1 ? s : src - 1;
}

// INTER_LINEAR tap of one axis: first source index, second source index, fixed-point weights (sum 2048)
__device__ __forceinline__ void linear_tap(int d, double scale, int src, int* s0, int* s1, int* w0, int* w1) {
    float f = static_cast<float>((static_cast<double>(d) + 0.5) * scale - 0.5);
    int s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
    *w1 = __float2int_rn(__fmul_rn(f, 2048.f));
    *w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    *s0 = s;
    *s1 = s + 1 < src ? s + 1 : src - 1;
}

// ------------------------------------------------------------------------------------------------
// RandomCrop: class histograms of the candidate boxes on the (virtual) nearest-resized label map
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aug_crop_hist_kernel(const uint8_t* __restrict__ seg, AugDev a,
                                                            unsigned int* __restrict__ hist) {
    pdl_sync();
    __shared__ unsigned int sh[256];
    const int t = blockIdx.y;                     // candidate
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int oy = a.p.box_y[t], ox = a.p.box_x[t];
    const long long total = static_cast<long long>(a.p.crop_h) * a.p.crop_w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / a.p.crop_w), x = static_cast<int>(i - static_cast<long long>(y) * a.p.crop_w);
        const int sy = nearest_src(oy + y, a.scale_y, a.p.H0), sx = nearest_src(ox + x, a.scale_x, a.p.W0);
        atomicAdd(&sh[seg[static_cast<long long>(sy) * a.p.W0 + sx]], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[t * 256 + threadIdx.x], sh[threadIdx.x]);
}

// the re-draw loop of RandomCrop.__call__: first candidate t < 10 with more than one class and max / sum < ratio, else 10
__global__ void aug_choose_kernel(const unsigned int* __restrict__ hist, AugDev a, int* __restrict__ chosen) {
    pdl_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int pick = kAugCandidates - 1;
    if (a.p.cat_max_ratio >= 1.f) pick = 0;
    else {
        for (int t = 0; t < kAugCandidates - 1; ++t) {
            unsigned long long sum = 0, mx = 0;
            int classes = 0;
            for (int c = 0; c < 256; ++c) {
                if (c == a.p.ignore_index) continue;
                const unsigned long long n = hist[t * 256 + c];
                if (n) { ++classes; sum += n; if (n > mx) mx = n; }
            }
            // numpy: np.max(cnt) / np.sum(cnt) < ratio   (float64 division; the ratio is promoted from float32 exactly)
            if (classes > 1 && static_cast<double>(mx) / static_cast<double>(sum) < static_cast<double>(a.p.cat_max_ratio)) {
                pick = t;
                break;
            }
        }
    }
    *chosen = pick;
}

// ------------------------------------------------------------------------------------------------
// 8-bit colour primitives (OpenCV color_hsv: RGB2HSV_b integer, HSV2RGB float with truncation)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bgr2hsv_u8(int b, int g, int r, int* h, int* s, int* v) {
    const int vmax = max(max(b, g), r), vmin = min(min(b, g), r);
    const int diff = vmax - vmin;
    *s = (diff * c_sdiv[vmax] + (1 << 11)) >> 12;
    int hh = vmax == r ? g - b : (vmax == g ? b - r + 2 * diff : r - g + 4 * diff);
    hh = (hh * c_hdiv[diff] + (1 << 11)) >> 12;
    if (hh < 0) hh += 180;
    *h = hh & 255;
    *v = vmax;
}

__device__ __forceinline__ int trunc_u8(float x) {
    x = floorf(x);
    return static_cast<int>(fminf(fmaxf(x, 0.f), 255.f));
}

__device__ __forceinline__ void hsv2bgr_u8(int h, int s, int v, int* b, int* g, int* r) {
    const float vf = __fmul_rn(static_cast<float>(v), 1.0f / 255.0f);
    if (s == 0) {
        const int q = trunc_u8(__fmul_rn(vf, 255.f));
        *b = *g = *r = q;
        return;
    }
    const float sf = __fmul_rn(static_cast<float>(s), 1.0f / 255.0f);
    float hh = __fmul_rn(static_cast<float>(h), 6.0f / 180.0f);
    if (hh >= 6.f) hh = __fsub_rn(hh, 6.f);
    int sector = static_cast<int>(floorf(hh));
    float fr = __fsub_rn(hh, static_cast<float>(sector));
    if (sector < 0 || sector >= 6) { sector = 0; fr = 0.f; }
    float tab[4];
    tab[0] = vf;
    tab[1] = __fmul_rn(vf, __fsub_rn(1.f, sf));
    tab[2] = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(sf, fr)));
    tab[3] = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(sf, __fsub_rn(1.f, fr))));
    // sector_data[][3] = {{1,3,0}, {1,0,2}, {3,0,1}, {0,2,1}, {0,1,3}, {2,1,0}}  (b, g, r)
    int ib, ig, ir;
    switch (sector) {
        case 0: ib = 1; ig = 3; ir = 0; break;
        case 1: ib = 1; ig = 0; ir = 2; break;
        case 2: ib = 3; ig = 0; ir = 1; break;
        case 3: ib = 0; ig = 2; ir = 1; break;
        case 4: ib = 0; ig = 1; ir = 3; break;
        default: ib = 2; ig = 1; ir = 0; break;
    }
    *b = trunc_u8(__fmul_rn(tab[ib], 255.f));
    *g = trunc_u8(__fmul_rn(tab[ig], 255.f));
    *r = trunc_u8(__fmul_rn(tab[ir], 255.f));
}

// PhotoMetricDistortion.convert: clip(x * alpha + beta, 0, 255).astype(uint8)
__device__ __forceinline__ int convert_u8(int x, float alpha, float beta) {
    const float y = __fadd_rn(__fmul_rn(static_cast<float>(x), alpha), beta);
    return static_cast<int>(fminf(fmaxf(y, 0.f), 255.f));       // clip, then truncation toward zero
}

// ------------------------------------------------------------------------------------------------
// the fused per-pixel pipeline
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aug_fused_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ seg,
                                                        AugDev a, const int* __restrict__ chosen, float* __restrict__ out_img,
                                                        long long* __restrict__ out_lab) {
    pdl_sync();
    const gs_aug_params& p = a.p;
    const int t = *chosen;
    const int oy = p.box_y[t], ox = p.box_x[t];
    const long long plane = static_cast<long long>(p.out_h) * p.out_w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < plane;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / p.out_w), x = static_cast<int>(i - static_cast<long long>(y) * p.out_w);
        if (y >= p.crop_h || x >= p.crop_w) {            // Pad: image 0 (after Normalize), label 255
            out_img[i] = 0.f; out_img[plane + i] = 0.f; out_img[2 * plane + i] = 0.f;
            out_lab[i] = 255;
            continue;
        }
        const int xs = p.flip ? p.crop_w - 1 - x : x;     // RandomFlip (horizontal) of the cropped tile
        const int ry = oy + y, rx = ox + xs;              // coordinates in the resized image
        // ---- label: INTER_NEAREST ----
        out_lab[i] = seg[static_cast<long long>(nearest_src(ry, a.scale_y, p.H0)) * p.W0 + nearest_src(rx, a.scale_x, p.W0)];
        // ---- image: INTER_LINEAR, OpenCV's 8-bit fixed point ----
        int c[3];
        if (p.new_h == p.H0 && p.new_w == p.W0) {
            const uint8_t* q = img + (static_cast<long long>(ry) * p.W0 + rx) * 3;
            c[0] = q[0]; c[1] = q[1]; c[2] = q[2];
        } else if (a.area2x) {
            const uint8_t* q0 = img + (static_cast<long long>(2 * ry) * p.W0 + 2 * rx) * 3;
            const uint8_t* q1 = q0 + static_cast<long long>(p.W0) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = (q0[k] + q0[3 + k] + q1[k] + q1[3 + k] + 2) >> 2;
        } else {
            int sx0, sx1, a0, a1, sy0, sy1, b0, b1;
            linear_tap(rx, a.scale_x, p.W0, &sx0, &sx1, &a0, &a1);
            linear_tap(ry, a.scale_y, p.H0, &sy0, &sy1, &b0, &b1);
            const uint8_t* r0 = img + static_cast<long long>(sy0) * p.W0 * 3;
            const uint8_t* r1 = img + static_cast<long long>(sy1) * p.W0 * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int h0 = r0[sx0 * 3 + k] * a0 + r0[sx1 * 3 + k] * a1;     // horizontal pass, scale 2^11
                const int h1 = r1[sx0 * 3 + k] * a0 + r1[sx1 * 3 + k] * a1;
                int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                c[k] = min(max(v, 0), 255);
            }
        }
        // ---- PhotoMetricDistortion on uint8 BGR ----
        if (p.has_brightness) {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = convert_u8(c[k], 1.f, p.brightness);
        }
        if (p.contrast_first && p.has_contrast) {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = convert_u8(c[k], p.contrast, 0.f);
        }
        if (p.has_saturation) {
            int h, s, v;
            bgr2hsv_u8(c[0], c[1], c[2], &h, &s, &v);
            s = convert_u8(s, p.saturation, 0.f);
            hsv2bgr_u8(h, s, v, &c[0], &c[1], &c[2]);
        }
        if (p.has_hue) {
            int h, s, v;
            bgr2hsv_u8(c[0], c[1], c[2], &h, &s, &v);
            h = ((h + p.hue) % 180 + 180) % 180;            // python's non-negative modulo
            hsv2bgr_u8(h, s, v, &c[0], &c[1], &c[2]);
        }
        if (!p.contrast_first && p.has_contrast) {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = convert_u8(c[k], p.contrast, 0.f);
        }
        // ---- Normalize (to_rgb) -> fp32 CHW ----
        out_img[i] = __fmul_rn(__fsub_rn(static_cast<float>(c[2]), p.mean[0]), p.inv_std[0]);              // R
        out_img[plane + i] = __fmul_rn(__fsub_rn(static_cast<float>(c[1]), p.mean[1]), p.inv_std[1]);      // G
        out_img[2 * plane + i] = __fmul_rn(__fsub_rn(static_cast<float>(c[0]), p.mean[2]), p.inv_std[2]);  // B
    }
}

static int make_dev(const gs_aug_params* p, AugDev* a) {
    GS_REQUIRE(p != nullptr, "augment: null parameters");
    GS_REQUIRE(p->H0 > 0 && p->W0 > 0 && p->new_h > 0 && p->new_w > 0 && p->out_h > 0 && p->out_w > 0,
               "augment: empty image");
    GS_REQUIRE(p->crop_h > 0 && p->crop_w > 0 && p->crop_h <= p->out_h && p->crop_w <= p->out_w && p->crop_h <= p->new_h &&
                   p->crop_w <= p->new_w,
               "augment: crop %dx%d does not fit output %dx%d / resized image %dx%d", p->crop_h, p->crop_w, p->out_h, p->out_w,
               p->new_h, p->new_w);
    for (int t = 0; t < kAugCandidates; ++t)
        GS_REQUIRE(p->box_y[t] >= 0 && p->box_x[t] >= 0 && p->box_y[t] + p->crop_h <= p->new_h &&
                       p->box_x[t] + p->crop_w <= p->new_w,
                   "augment: candidate box %d (%d, %d) leaves the resized image", t, p->box_y[t], p->box_x[t]);
    GS_REQUIRE(p->ignore_index >= 0 && p->ignore_index < 256, "augment: ignore_index must fit a uint8 label map");
    a->p = *p;
    a->scale_x = 1.0 / (static_cast<double>(p->new_w) / p->W0);
    a->scale_y = 1.0 / (static_cast<double>(p->new_h) / p->H0);
    a->area2x = (p->H0 == 2 * p->new_h && p->W0 == 2 * p->new_w) ? 1 : 0;
    static bool tables = false;
    if (!tables) {
        int sdiv[256], hdiv[256];
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; ++i) {
            sdiv[i] = static_cast<int>(nearbyint((255 << 12) / (1.0 * i)));
            hdiv[i] = static_cast<int>(nearbyint((180 << 12) / (6.0 * i)));
        }
        GS_CUDA_OK(cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv)));
        GS_CUDA_OK(cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv)));
        tables = true;
    }
    return 0;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_aug_workspace_bytes(void) { return static_cast<int64_t>(kAugCandidates) * 256 * 4 + 16; }

extern "C" int gs_aug_choose_crop(const uint8_t* seg, const gs_aug_params* params, void* workspace, int32_t* chosen_dev,
                                  void* stream) {
    AugDev a;
    if (make_dev(params, &a)) return -1;
    GS_REQUIRE(seg && workspace && chosen_dev, "aug_choose_crop: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned int* hist = reinterpret_cast<unsigned int*>(workspace);
    GS_CUDA_OK(cudaMemsetAsync(hist, 0, static_cast<size_t>(kAugCandidates) * 256 * 4, st));
    const long long total = static_cast<long long>(a.p.crop_h) * a.p.crop_w;
    int bx = static_cast<int>((total + 256 * 16 - 1) / (256 * 16));
    if (bx < 1) bx = 1;
    if (bx > 148) bx = 148;
    gs::launch(aug_crop_hist_kernel, dim3(bx, kAugCandidates), dim3(256), 0, st, seg, a, hist);
    GS_LAUNCHED();
    gs::launch(aug_choose_kernel, dim3(1), dim3(32), 0, st, hist, a, chosen_dev);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_aug_fused(const uint8_t* img_bgr, const uint8_t* seg, const gs_aug_params* params,
                            const int32_t* chosen_dev, float* out_img, int64_t* out_labels, void* stream) {
    AugDev a;
    if (make_dev(params, &a)) return -1;
    GS_REQUIRE(img_bgr && seg && chosen_dev && out_img && out_labels, "aug_fused: null pointer");
    const long long plane = static_cast<long long>(a.p.out_h) * a.p.out_w;
    long long blocks = (plane + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gs::launch(aug_fused_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), img_bgr,
               seg, a, chosen_dev, out_img, reinterpret_cast<long long*>(out_labels));
    GS_LAUNCHED();
    return 0;
}
