// include/cuda/hungarian.h — posebyte::cuda::LinearAssignmentCUDA over the B200 C ABI.
//
// The device-pointer entry points of the reference class (reference include/cuda/hungarian.h,
// src/cuda/hungarian.cu:341-405): forward auction with eps0 = 1/(rows+1), min(3*rows, 50)
// iterations, eps *= 0.9, lowest column on equal value and lowest row on equal bid.  `threshold`
// is accepted and ignored, as upstream.  One launch per solve instead of 100 kernels + 53
// memsets.  The host-side legacy solve() (hungarian.cu:235-339) and GreedyMatcherCUDA are not
// part of the hot path and are not provided.
#pragma once

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class LinearAssignmentCUDA {
public:
    explicit LinearAssignmentCUDA(int max_size = 256) : max_size_(max_size) {}

    void solveDeviceAsync(const float* d_cost_matrix, int num_rows, int num_cols, int* d_row_assignments,
                          int* d_col_assignments, float threshold, cudaStream_t stream = 0) {
        solveDeviceAsyncWithActive(d_cost_matrix, num_rows, num_cols, d_row_assignments, d_col_assignments, nullptr, threshold, stream);
    }

    void solveDeviceAsyncWithActive(const float* d_cost_matrix, int num_rows, int num_cols, int* d_row_assignments,
                                    int* d_col_assignments, const int* d_row_active, float /*threshold*/, cudaStream_t stream = 0) {
        detail::pb_check(pb_auction_solve(d_cost_matrix, 1, num_rows, num_cols, d_row_assignments, d_col_assignments, d_row_active,
                                          detail::as_pb(stream)), "pb_auction_solve");
    }

    // `batch` independent problems of the same shape in one launch (new capability).
    void solveBatchDeviceAsync(const float* d_cost_matrices, int batch, int num_rows, int num_cols, int* d_row_assignments,
                               int* d_col_assignments, const int* d_row_active, cudaStream_t stream = 0) {
        detail::pb_check(pb_auction_solve(d_cost_matrices, batch, num_rows, num_cols, d_row_assignments, d_col_assignments, d_row_active,
                                          detail::as_pb(stream)), "pb_auction_solve");
    }

    void sync(cudaStream_t stream = 0) { detail::cu_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize"); }
    int getMaxSize() const { return max_size_; }

private:
    int max_size_;
};

}  // namespace cuda
}  // namespace posebyte
