// include/cuda/pb_shim_common.h — helpers shared by the header-only C++ views of the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>

#include "../posebyte_b200.h"
#include "../types.h"

namespace posebyte {
namespace cuda {
namespace detail {

// The reference prints and carries on after a failed CUDA call (gpu_tracker.cu:9-16) or
// exits the process (hungarian.cu:11-19).  The views throw instead: a failed call is
// never silently ignored and never takes the host application down.
inline void pb_check(int status, const char* what) {
    if (status != PB_OK) throw std::runtime_error(std::string(what) + ": " + pb_last_error());
}
inline void cu_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline pb_stream_t as_pb(cudaStream_t s) { return reinterpret_cast<pb_stream_t>(s); }

// One-stream handle with RAII ownership.
class Handle {
public:
    explicit Handle(const pb_config& cfg) { pb_check(pb_create(&cfg, &h_), "pb_create"); }
    ~Handle() { pb_destroy(h_); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    pb_handle_t get() const { return h_; }
    pb_device_views views() const {
        pb_device_views v{};
        pb_check(pb_get_device_views(h_, &v), "pb_get_device_views");
        return v;
    }
private:
    pb_handle_t h_ = nullptr;
};

}  // namespace detail
}  // namespace cuda
}  // namespace posebyte
