// include/cuda/hungarian.h — posebyte::cuda::LinearAssignmentCUDA over the B200 C ABI.
//
// The device-pointer entry points of the reference class (reference include/cuda/hungarian.h,
// src/cuda/hungarian.cu:341-405): forward auction with eps0 = 1/(rows+1), min(3*rows, 50)
// iterations, eps *= 0.9, lowest column on equal value and lowest row on equal bid.  `threshold`
// is accepted and ignored, as upstream.  One launch per solve instead of 100 kernels + 53
// memsets.  GreedyMatcherCUDA (hungarian.cu:407-543) is provided with the deterministic rule of
// its own host path.  The host-side legacy LinearAssignmentCUDA::solve() (hungarian.cu:235-339:
// greedy below 100 cells, else 3*rows auction iterations and a host threshold filter) runs as one
// launch of pb_assign_legacy; solveDevice() (:412-434) is the synchronous device-pointer form.
#pragma once

#include <utility>
#include <vector>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class LinearAssignmentCUDA {
public:
    explicit LinearAssignmentCUDA(int max_size = 256) : max_size_(max_size) {}
    ~LinearAssignmentCUDA() { cudaFree(d_buf_); cudaFree(d_prices_); }
    LinearAssignmentCUDA(const LinearAssignmentCUDA&) = delete;
    LinearAssignmentCUDA& operator=(const LinearAssignmentCUDA&) = delete;

    // Legacy entry point (hungarian.cu:235-339): host cost matrix in, host assignments out, returns the
    // number of assignments with cost <= threshold.  Blocking, like upstream.
    int solve(const float* cost_matrix, int num_rows, int num_cols, int* row_assignments, int* col_assignments,
              float threshold = 1.0f) {
        if (num_rows == 0 || num_cols == 0) return 0;                          // :243
        const size_t cells = (size_t)num_rows * num_cols;
        reserve((size_t)(num_rows > max_size_ ? num_rows : max_size_), (size_t)(num_cols > max_size_ ? num_cols : max_size_));
        float* d_cost = static_cast<float*>(d_buf_);
        int* d_row = reinterpret_cast<int*>(d_cost + cells_cap_);
        int* d_col = d_row + rows_cap_;
        int* d_cnt = d_col + cols_cap_;
        detail::cu_check(cudaMemcpy(d_cost, cost_matrix, cells * sizeof(float), cudaMemcpyHostToDevice), "upload");
        detail::pb_check(pb_assign_legacy(d_cost, 1, num_rows, num_cols, threshold, d_row, d_col, d_cnt, nullptr), "pb_assign_legacy");
        int count = 0;
        detail::cu_check(cudaMemcpy(row_assignments, d_row, (size_t)num_rows * sizeof(int), cudaMemcpyDeviceToHost), "download");
        detail::cu_check(cudaMemcpy(col_assignments, d_col, (size_t)num_cols * sizeof(int), cudaMemcpyDeviceToHost), "download");
        detail::cu_check(cudaMemcpy(&count, d_cnt, sizeof(int), cudaMemcpyDeviceToHost), "download");
        return count;
    }

    // Synchronous device-pointer solve (hungarian.cu:412-434): returns the number of assigned rows.
    int solveDevice(float* d_cost_matrix, int num_rows, int num_cols, int* d_row_assignments, int* d_col_assignments,
                    float threshold = 1.0f) {
        solveDeviceAsync(d_cost_matrix, num_rows, num_cols, d_row_assignments, d_col_assignments, threshold, nullptr);
        std::vector<int> h_row((size_t)num_rows);
        detail::cu_check(cudaMemcpy(h_row.data(), d_row_assignments, h_row.size() * sizeof(int), cudaMemcpyDeviceToHost), "download");
        int count = 0;
        for (int v : h_row) count += (v >= 0);
        return count;
    }

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

    // Device memory access (reference hungarian.h:78-81): the object's own buffers, max_size x max_size cost cells and
    // max_size assignments per side, valid until the object dies.  solve() stages its problem in them ([rows, cols]
    // row-major at the start of the cost buffer), so after a solve() they hold its cost matrix and assignments as upstream.
    // The prices of a solve live in registers / shared memory of the solve kernel and are not written back: the price
    // buffer is provided for API parity and stays zero.
    float* getCostMatrixDevice() { reserve((size_t)max_size_, (size_t)max_size_); return static_cast<float*>(d_buf_); }
    int* getRowAssignmentsDevice() { reserve((size_t)max_size_, (size_t)max_size_); return reinterpret_cast<int*>(static_cast<float*>(d_buf_) + cells_cap_); }
    int* getColAssignmentsDevice() { return getRowAssignmentsDevice() + rows_cap_; }
    float* getPricesDevice() {
        if (!d_prices_) {
            detail::cu_check(cudaMalloc(&d_prices_, (size_t)max_size_ * sizeof(float)), "cudaMalloc");
            detail::cu_check(cudaMemset(d_prices_, 0, (size_t)max_size_ * sizeof(float)), "cudaMemset");
        }
        return d_prices_;
    }

private:
    // one allocation: cost [cells_cap_] | row [rows_cap_] | col [cols_cap_] | count
    void reserve(size_t rows, size_t cols) {
        if (d_buf_ && rows <= rows_cap_ && cols <= cols_cap_ && rows * cols <= cells_cap_) return;
        const size_t r = rows > rows_cap_ ? rows : rows_cap_, c = cols > cols_cap_ ? cols : cols_cap_;
        const size_t cells = r * c > cells_cap_ ? r * c : cells_cap_;
        cudaFree(d_buf_); d_buf_ = nullptr;
        detail::cu_check(cudaMalloc(&d_buf_, cells * sizeof(float) + (r + c + 1) * sizeof(int)), "cudaMalloc");
        rows_cap_ = r; cols_cap_ = c; cells_cap_ = cells;
    }
    int max_size_;
    void* d_buf_ = nullptr;
    size_t rows_cap_ = 0, cols_cap_ = 0, cells_cap_ = 0;
    float* d_prices_ = nullptr;
};

// GreedyMatcherCUDA: cells below the threshold in ascending (cost, row, col) order, each taken when
// its row and column are still free — the rule of the reference's own host path (hungarian.cu:441-467).
// The reference's device kernel (:126-157) races on the columns; here host and device entry points
// give the same, deterministic answer.
class GreedyMatcherCUDA {
public:
    explicit GreedyMatcherCUDA(int max_size = 256) : max_size_(max_size) {
        detail::cu_check(cudaStreamCreate(&stream_), "cudaStreamCreate");
        detail::cu_check(cudaMalloc(&d_costs_, (size_t)max_size * max_size * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_row_, (size_t)max_size * sizeof(int)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_col_, (size_t)max_size * sizeof(int)), "cudaMalloc");
        detail::cu_check(cudaMemset(d_col_, 0xff, (size_t)max_size * sizeof(int)), "cudaMemset");
    }
    ~GreedyMatcherCUDA() {
        cudaFree(d_costs_); cudaFree(d_row_); cudaFree(d_col_);
        cudaStreamDestroy(stream_);
    }
    GreedyMatcherCUDA(const GreedyMatcherCUDA&) = delete;
    GreedyMatcherCUDA& operator=(const GreedyMatcherCUDA&) = delete;

    // host cost matrix [num_rows, num_cols] -> (row, col) pairs in ascending row order
    std::vector<std::pair<int, int>> match(const float* cost_matrix, int num_rows, int num_cols, float threshold) {
        std::vector<std::pair<int, int>> out;
        if (num_rows == 0 || num_cols == 0) return out;
        if (num_rows > max_size_ || num_cols > max_size_) throw std::runtime_error("GreedyMatcherCUDA: matrix larger than max_size");
        detail::cu_check(cudaMemcpyAsync(d_costs_, cost_matrix, (size_t)num_rows * num_cols * sizeof(float), cudaMemcpyHostToDevice, stream_), "upload");
        matchDeviceAsync(d_costs_, num_rows, num_cols, d_row_, threshold, stream_);
        std::vector<int> rows((size_t)num_rows);
        detail::cu_check(cudaMemcpyAsync(rows.data(), d_row_, rows.size() * sizeof(int), cudaMemcpyDeviceToHost, stream_), "download");
        detail::cu_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
        std::vector<int> cols((size_t)num_cols, -1);
        for (int r = 0; r < num_rows; ++r) if (rows[r] >= 0) { out.push_back({r, rows[r]}); cols[(size_t)rows[r]] = r; }
        detail::cu_check(cudaMemcpy(d_col_, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice), "upload");   // getColMatchedDevice()
        return out;
    }
    // d_row_matched [num_rows] = matched column or -1; asynchronous
    void matchDeviceAsync(const float* d_cost_matrix, int num_rows, int num_cols, int* d_row_matched, float threshold, cudaStream_t stream = 0) {
        if (num_rows == 0 || num_cols == 0) return;
        detail::pb_check(pb_greedy_match(d_cost_matrix, 1, num_rows, num_cols, threshold, d_row_matched, detail::as_pb(stream ? stream : stream_)), "pb_greedy_match");
    }
    void sync(cudaStream_t stream = 0) { detail::cu_check(cudaStreamSynchronize(stream ? stream : stream_), "cudaStreamSynchronize"); }

    // Device memory access (reference hungarian.h:149-151): the buffers match() stages its problem in; after match() they
    // hold its cost matrix, the matched column of every row and the matched row of every column (-1: unmatched).
    float* getCostsDevice() { return d_costs_; }
    int* getRowMatchedDevice() { return d_row_; }
    int* getColMatchedDevice() { return d_col_; }

private:
    int max_size_;
    float* d_costs_ = nullptr;
    int* d_row_ = nullptr;
    int* d_col_ = nullptr;
    cudaStream_t stream_ = nullptr;
};

}  // namespace cuda
}  // namespace posebyte
