// include/cuda/preprocess.h — posebyte::cuda::PreprocessorCUDA over the B200 C ABI.
//
// Reference: include/cuda/preprocess.h + src/cuda/preprocess.cu:19-153 — letterbox resize (bilinear),
// BGR -> RGB, /255, HWC u8 -> CHW fp32 with gray 114/255 bars, one frame per call, blocking.
// Here the work is pb_letterbox_batch (one launch for any number of frames of individual sizes);
// this class is its one-frame view with the reference's signature, plus preprocessBatchAsync for
// callers that feed a batched engine.
#pragma once

#include <cstdint>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class PreprocessorCUDA {
public:
    PreprocessorCUDA(int max_input_width, int max_input_height, int target_width, int target_height)
        : max_w_(max_input_width), max_h_(max_input_height), tw_(target_width), th_(target_height) {
        detail::cu_check(cudaMalloc(&d_input_, (size_t)max_w_ * max_h_ * 3), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_output_, (size_t)3 * tw_ * th_ * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_meta_, 2 * sizeof(int) + 4 * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaStreamCreate(&stream_), "cudaStreamCreate");
    }
    ~PreprocessorCUDA() {
        cudaFree(d_input_); cudaFree(d_output_); cudaFree(d_meta_);
        cudaStreamDestroy(stream_);
    }
    PreprocessorCUDA(const PreprocessorCUDA&) = delete;
    PreprocessorCUDA& operator=(const PreprocessorCUDA&) = delete;

    // input_bgr: host memory [input_height, input_width, 3]; output_tensor: device [3, target_height, target_width].
    // scale_x / scale_y / pad_x / pad_y as upstream (preprocess.cu:117-122).  Blocking.
    void preprocess(const uint8_t* input_bgr, int input_width, int input_height, float* output_tensor,
                    float& scale_x, float& scale_y, int& pad_x, int& pad_y) {
        if (input_width <= 0 || input_height <= 0 || input_width > max_w_ || input_height > max_h_)
            throw std::runtime_error("PreprocessorCUDA: frame larger than the configured maximum");
        const int wh[2] = {input_width, input_height};
        int* d_wh = static_cast<int*>(d_meta_);
        float* d_xf = reinterpret_cast<float*>(d_wh + 2);
        detail::cu_check(cudaMemcpyAsync(d_input_, input_bgr, (size_t)input_width * input_height * 3, cudaMemcpyHostToDevice, stream_), "upload");
        detail::cu_check(cudaMemcpyAsync(d_wh, wh, sizeof(wh), cudaMemcpyHostToDevice, stream_), "upload");
        detail::pb_check(pb_letterbox_batch(d_input_, 0, d_wh, 1, tw_, th_, output_tensor, d_xf, detail::as_pb(stream_)), "pb_letterbox_batch");
        float xf[4];
        detail::cu_check(cudaMemcpyAsync(xf, d_xf, sizeof(xf), cudaMemcpyDeviceToHost, stream_), "download");
        detail::cu_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
        scale_x = xf[0]; scale_y = xf[1]; pad_x = (int)xf[2]; pad_y = (int)xf[3];
    }

    // Batched, asynchronous, device pointers only (new capability): see pb_letterbox_batch.
    void preprocessBatchAsync(const uint8_t* d_frames, size_t frame_stride_bytes, const int* d_sizes, int batch,
                              float* d_out, float* d_xform, cudaStream_t stream = 0) {
        detail::pb_check(pb_letterbox_batch(d_frames, frame_stride_bytes, d_sizes, batch, tw_, th_, d_out, d_xform, detail::as_pb(stream)),
                         "pb_letterbox_batch");
    }

    float* getDeviceOutput() { return d_output_; }

private:
    int max_w_, max_h_, tw_, th_;
    unsigned char* d_input_ = nullptr;
    float* d_output_ = nullptr;
    void* d_meta_ = nullptr;
    cudaStream_t stream_ = nullptr;
};

}  // namespace cuda
}  // namespace posebyte
