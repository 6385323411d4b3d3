// include/cuda/nms.h — posebyte::cuda::NMSCuda and launchPoseNMS over the B200 C ABI.
//
// NMSCuda::apply / applyBatch keep the reference's signature and rule set (reference
// include/cuda/nms.h:11-45, src/cuda/nms.cu:142-330: score filter, descending score order,
// IoU > 0.55, OKS > 0.5, IoU > 0.2 with OKS > 0.4, centre distance with OKS > 0.15; the
// oks_threshold argument is ignored upstream and therefore here), but run on the GPU — the
// reference evaluates them on the host.  launchPoseNMS is the device-pointer entry point
// the reference declares (nms.h:48-60) and never defines; it is exported by the library.
#pragma once

#include <vector>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class NMSCuda {
public:
    explicit NMSCuda(int max_detections = 1024) : cap_(max_detections) {
        detail::cu_check(cudaStreamCreate(&stream_), "cudaStreamCreate");
    }
    ~NMSCuda() {
        release();
        cudaStreamDestroy(stream_);
    }
    NMSCuda(const NMSCuda&) = delete;
    NMSCuda& operator=(const NMSCuda&) = delete;

    // Indices (into `detections`) of the detections to keep, best score first.
    std::vector<int> apply(const PoseDetection* detections, int num_detections, float oks_threshold = 0.65f,
                           float score_threshold = 0.25f) {
        const int one = num_detections;
        std::vector<std::vector<int>> r = applyBatch(detections, &one, 1, oks_threshold, score_threshold);
        return r.empty() ? std::vector<int>() : r[0];
    }

    // `detections` holds the images back to back; indices are image-local.
    std::vector<std::vector<int>> applyBatch(const PoseDetection* detections, const int* num_per_image, int num_images,
                                             float oks_threshold = 0.65f, float score_threshold = 0.25f) {
        std::vector<std::vector<int>> out((size_t)(num_images > 0 ? num_images : 0));
        if (num_images <= 0) return out;
        std::vector<int> offsets((size_t)num_images + 1, 0);
        int max_per = 0;
        for (int i = 0; i < num_images; ++i) {
            const int n = num_per_image[i] > 0 ? num_per_image[i] : 0;
            offsets[(size_t)i + 1] = offsets[i] + n;
            if (n > max_per) max_per = n;
        }
        const int total = offsets[(size_t)num_images];
        if (total == 0) return out;
        reserve(total, num_images);
        detail::cu_check(cudaMemcpyAsync(d_dets_, detections, (size_t)total * sizeof(PoseDetection), cudaMemcpyHostToDevice, stream_), "upload detections");
        detail::cu_check(cudaMemcpyAsync(d_off_, offsets.data(), offsets.size() * sizeof(int), cudaMemcpyHostToDevice, stream_), "upload offsets");
        detail::pb_check(pb_nms_legacy(d_dets_, d_off_, num_images, max_per, oks_threshold, score_threshold, d_keep_, d_nkeep_,
                                       detail::as_pb(stream_)), "pb_nms_legacy");
        std::vector<int> keep((size_t)total), nkeep((size_t)num_images);
        detail::cu_check(cudaMemcpyAsync(keep.data(), d_keep_, keep.size() * sizeof(int), cudaMemcpyDeviceToHost, stream_), "read keep");
        detail::cu_check(cudaMemcpyAsync(nkeep.data(), d_nkeep_, nkeep.size() * sizeof(int), cudaMemcpyDeviceToHost, stream_), "read counts");
        detail::cu_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
        for (int i = 0; i < num_images; ++i)
            out[i].assign(keep.begin() + offsets[i], keep.begin() + offsets[i] + nkeep[i]);
        return out;
    }

private:
    void reserve(int total, int images) {
        if (total > have_total_) {
            if (d_dets_) cudaFree(d_dets_);
            if (d_keep_) cudaFree(d_keep_);
            const int n = total > cap_ ? total : cap_;
            detail::cu_check(cudaMalloc(&d_dets_, (size_t)n * sizeof(PoseDetection)), "cudaMalloc");
            detail::cu_check(cudaMalloc(&d_keep_, (size_t)n * sizeof(int)), "cudaMalloc");
            have_total_ = n;
        }
        if (images > have_images_) {
            if (d_off_) cudaFree(d_off_);
            if (d_nkeep_) cudaFree(d_nkeep_);
            detail::cu_check(cudaMalloc(&d_off_, ((size_t)images + 1) * sizeof(int)), "cudaMalloc");
            detail::cu_check(cudaMalloc(&d_nkeep_, (size_t)images * sizeof(int)), "cudaMalloc");
            have_images_ = images;
        }
    }
    void release() {
        if (d_dets_) cudaFree(d_dets_);
        if (d_keep_) cudaFree(d_keep_);
        if (d_off_) cudaFree(d_off_);
        if (d_nkeep_) cudaFree(d_nkeep_);
    }
    int cap_;
    int have_total_ = 0, have_images_ = 0;
    PoseDetection* d_dets_ = nullptr;
    int *d_keep_ = nullptr, *d_off_ = nullptr, *d_nkeep_ = nullptr;
    cudaStream_t stream_ = nullptr;
};

}  // namespace cuda
}  // namespace posebyte
