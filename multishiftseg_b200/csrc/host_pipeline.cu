// Host-buffer entry points: the call a reference-side user holding CPU tensors makes (bench.py `e2e`).
// Replaces test_deeplab.py:90-94's `.cuda()` -> model tail -> `.cpu().numpy()` round trip for the scoring
// step: images are streamed through three device slots so that the H2D copy of image b+1, the scoring
// kernel of image b and the D2H copy of image b-1 overlap (PCIe is full duplex).
#include "common.cuh"

using namespace mss;

static int n_maps(unsigned which) { return __builtin_popcount(which & 15u); }
constexpr int HOST_SLOTS = 3;

extern "C" size_t mss_deeplab_score_host_scratch_bytes(int64_t B, int C, int64_t HW, unsigned which) {
    (void)B;
    const size_t per_slot = align_up((size_t)C * HW * 4, 256) + (size_t)n_maps(which) * align_up((size_t)HW * 4, 256);
    return HOST_SLOTS * per_slot + 256;
}

extern "C" int mss_deeplab_score_host(const float *logits_host, int64_t B, int C, int64_t HW, unsigned which,
                                      float *energy_host, float *maxlogit_host, float *msp_host, float *entropy_host,
                                      void *device_scratch, size_t scratch_bytes, void *stream) {
    MSS_REQUIRE(logits_host && device_scratch && B >= 0 && C >= 1 && HW >= 0, "mss_deeplab_score_host: bad arguments");
    MSS_REQUIRE((which & ~15u) == 0 && which != 0, "mss_deeplab_score_host: bad `which` mask 0x%x", which);
    if (scratch_bytes < mss_deeplab_score_host_scratch_bytes(B, C, HW, which)) {
        set_error("mss_deeplab_score_host: scratch too small (%zu < %zu)", scratch_bytes,
                  mss_deeplab_score_host_scratch_bytes(B, C, HW, which));
        return MSS_ERR_WORKSPACE;
    }
    float *host_out[4] = {energy_host, maxlogit_host, msp_host, entropy_host};
    for (int k = 0; k < 4; k++)
        MSS_REQUIRE(!((which >> k) & 1u) || host_out[k], "mss_deeplab_score_host: selected output %d is NULL", k);
    if (B == 0 || HW == 0) return MSS_OK;

    cudaStream_t user = (cudaStream_t)stream;
    cudaStream_t st[HOST_SLOTS];
    cudaEvent_t ev_start, ev_done[HOST_SLOTS];
    MSS_CHECK_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    MSS_CHECK_CUDA(cudaEventRecord(ev_start, user));
    Carver cv(device_scratch, scratch_bytes);
    float *d_logits[HOST_SLOTS], *d_map[HOST_SLOTS][4];
    for (int s = 0; s < HOST_SLOTS; s++) {
        MSS_CHECK_CUDA(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
        MSS_CHECK_CUDA(cudaEventCreateWithFlags(&ev_done[s], cudaEventDisableTiming));
        MSS_CHECK_CUDA(cudaStreamWaitEvent(st[s], ev_start, 0));
        d_logits[s] = cv.take<float>((size_t)C * HW);
        for (int k = 0; k < 4; k++) d_map[s][k] = ((which >> k) & 1u) ? cv.take<float>((size_t)HW) : nullptr;
    }
    int rc = MSS_OK;
    for (int64_t b = 0; b < B && rc == MSS_OK; b++) {
        const int s = (int)(b % HOST_SLOTS);
        cudaError_t e = cudaMemcpyAsync(d_logits[s], logits_host + (size_t)b * C * HW, (size_t)C * HW * 4,
                                        cudaMemcpyHostToDevice, st[s]);
        if (e != cudaSuccess) { set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = MSS_ERR_CUDA; break; }
        rc = mss_deeplab_score(d_logits[s], 1, C, HW, which, d_map[s][0], d_map[s][1], d_map[s][2], d_map[s][3],
                               nullptr, 0, 0, 1, 0, nullptr, st[s]);
        if (rc) break;
        for (int k = 0; k < 4; k++) {
            if (!d_map[s][k]) continue;
            e = cudaMemcpyAsync(host_out[k] + (size_t)b * HW, d_map[s][k], (size_t)HW * 4, cudaMemcpyDeviceToHost, st[s]);
            if (e != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = MSS_ERR_CUDA; break; }
        }
    }
    for (int s = 0; s < HOST_SLOTS; s++) {
        cudaEventRecord(ev_done[s], st[s]);
        cudaStreamWaitEvent(user, ev_done[s], 0);
    }
    cudaError_t e = cudaStreamSynchronize(user);
    for (int s = 0; s < HOST_SLOTS; s++) { cudaStreamDestroy(st[s]); cudaEventDestroy(ev_done[s]); }
    cudaEventDestroy(ev_start);
    if (rc == MSS_OK && e != cudaSuccess) { set_error("stream sync failed: %s", cudaGetErrorString(e)); rc = MSS_ERR_CUDA; }
    return rc;
}
