#pragma once
// utils.cuh — declarations for src/utils.cu (the README of the reference names a utils header that
// was never written: README.md:14-15).
#include<cmath>
#include<vector>
#include<utility>
#include<tensor.cuh>

namespace utils {
    // [begin, end) of a global index range owned by `rank` (bit-exact with the device sharding).
    std::pair<size_t, size_t> shard(size_t n, int rank, int nranks);
    // Regular evaluation grid [per_axis^d, d] over the box lo..hi (row-major, last axis fastest).
    tensor<float> grid(const std::vector<float> &lo, const std::vector<float> &hi, size_t per_axis);
    // Map points of the box lo..hi to [-1, 1] in place.
    tensor<float> &normalize(tensor<float> &x, const std::vector<float> &lo, const std::vector<float> &hi);
    // || a - b ||_2 / || b ||_2
    float relative_l2(const tensor<float> &a, const tensor<float> &b);
}
