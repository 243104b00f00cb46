// (a3, a4) bilinear resize, fp32, [NC, h, w] -> [NC, H, W].
//   align_corners=1 : mynn.Upsample (lib/network/deepv3/mynn.py:28-33), used at deepv3.py:283
//   align_corners=0 : F.interpolate at lib/network/mask2former/maskformer_model.py:264-277
// Index / weight arithmetic follows ATen's area_pixel_compute_source_index in fp32:
//   scale = align ? (in-1)/(out-1) : in/out ;  src = align ? scale*dst : max(0, scale*(dst+0.5)-0.5)
//   i0 = (int)src ; i1 = i0 + (i0 < in-1) ; l1 = src - i0 ; l0 = 1 - l1
//   out = h0*(w0*a + w1*b) + h1*(w0*c + w1*d)
// HBM-bound: 4 B/px written, 4*(h*w)/(H*W) B/px read (the 4 taps hit L1/L2).
#include "common.cuh"

namespace mss {

__device__ __forceinline__ void src_index(int dst, float scale, int align, int in_size, int &i0, int &i1,
                                          float &l0, float &l1) {
    float src = align ? scale * (float)dst : fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, in_size - 1);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    l0 = 1.f - l1;
}

// thread <-> 4 adjacent output pixels of one row (128-bit store when W % 4 == 0)
__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const float *__restrict__ in, int h, int w, float *__restrict__ out, int H, int W,
                         float sh, float sw, int align, long long n_quads, int Wq, int vec_store) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_quads) return;
    const int xq = (int)(q % Wq);
    const long long row = q / Wq;          // nc * H + y
    const int y = (int)(row % H);
    const long long nc = row / H;
    int y0, y1;
    float hl0, hl1;
    src_index(y, sh, align, h, y0, y1, hl0, hl1);
    const float *r0 = in + (nc * h + y0) * (long long)w;
    const float *r1 = in + (nc * h + y1) * (long long)w;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int x = xq * 4 + j;
        if (x < W) {
            int x0, x1;
            float wl0, wl1;
            src_index(x, sw, align, w, x0, x1, wl0, wl1);
            float a = __ldg(r0 + x0), b = __ldg(r0 + x1), c = __ldg(r1 + x0), d = __ldg(r1 + x1);
            o[j] = hl0 * (wl0 * a + wl1 * b) + hl1 * (wl0 * c + wl1 * d);
        }
    }
    float *dst = out + row * (long long)W + xq * 4;
    if (vec_store) {
        stg_stream_f4(dst, make_float4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (xq * 4 + j < W) dst[j] = o[j];
    }
}

}  // namespace mss

using namespace mss;

static float resize_scale(int in, int out, int align) {
    if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    return (float)in / (float)out;
}

extern "C" int mss_upsample_bilinear(const float *in, int64_t NC, int h, int w, float *out, int H, int W,
                                     int align_corners, void *stream) {
    MSS_REQUIRE(in && out, "mss_upsample_bilinear: null pointer");
    MSS_REQUIRE(NC >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "mss_upsample_bilinear: bad shape");
    if (NC == 0) return MSS_OK;
    const int Wq = (W + 3) / 4;
    const long long n_quads = (long long)NC * H * Wq;
    const int vec = (W % 4 == 0) && (((uintptr_t)out & 15) == 0);
    long long blocks = (n_quads + 255) / 256;
    MSS_REQUIRE(blocks < (1ll << 31), "mss_upsample_bilinear: tensor too large");
    upsample_bilinear_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        in, h, w, out, H, W, resize_scale(h, H, align_corners), resize_scale(w, W, align_corners),
        align_corners ? 1 : 0, n_quads, Wq, vec);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_deeplab_anomaly_score(const float *ood_logits, int64_t B, int C, int h, int w,
                                         float *scratch, float *score, int H, int W, void *stream) {
    MSS_REQUIRE(scratch && score, "mss_deeplab_anomaly_score: null pointer");
    int rc = mss_deeplab_score(ood_logits, B, C, (int64_t)h * w, MSS_SCORE_ENERGY, scratch, nullptr, nullptr,
                               nullptr, nullptr, 0, 0, 1, 0, nullptr, stream);
    if (rc) return rc;
    return mss_upsample_bilinear(scratch, B, h, w, score, H, W, 1, stream);
}
