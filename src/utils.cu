// utils.cu — host-side helpers around the collocation sets (reference src/utils.cu is 0 bytes;
// README.md:54 "data handling and normalization").  Sampling itself runs on the GPU inside the
// library (pinn_sample_points); these are the pieces a training driver needs around it.
#include<network.cuh>
#include<utils.cuh>

namespace utils {
    std::pair<size_t, size_t> shard(size_t n, int rank, int nranks) {
        if (nranks < 1 || rank < 0 || rank >= nranks)
            throw std::invalid_argument{"utils::shard: invalid rank or world size"};
        uint64_t b = 0, e = 0;
        pinn_shard_range(n, rank, nranks, &b, &e);
        return {static_cast<size_t>(b), static_cast<size_t>(e)};
    }

    tensor<float> grid(const std::vector<float> &lo, const std::vector<float> &hi, size_t per_axis) {
        if (lo.size() != hi.size() || lo.empty() || per_axis < 2)
            throw std::invalid_argument{"utils::grid: need matching bounds and at least two nodes per axis"};
        const size_t d = lo.size();
        size_t n = 1;
        for(size_t k = 0; k < d; k++) n *= per_axis;
        tensor<float> x(as_shape, {n, d});
        for(size_t i = 0; i < n; i++) {
            size_t r = i;
            for(size_t k = d; k-- > 0;) {
                const size_t idx = r % per_axis; r /= per_axis;
                x[i * d + k] = lo[k] + (hi[k] - lo[k]) * static_cast<float>(idx) / static_cast<float>(per_axis - 1);
            }
        }
        return x;
    }

    tensor<float> &normalize(tensor<float> &x, const std::vector<float> &lo, const std::vector<float> &hi) {
        if (x.dim() != 2 || x.get_shape()[1] != lo.size() || lo.size() != hi.size())
            throw std::invalid_argument{"utils::normalize: expected a [n, d] matrix and d bounds"};
        const size_t n = x.get_shape()[0], d = lo.size();
        for(size_t i = 0; i < n; i++)
            for(size_t k = 0; k < d; k++)
                x[i * d + k] = 2.f * (x.raw()[i * d + k] - lo[k]) / (hi[k] - lo[k]) - 1.f;
        return x;
    }

    float relative_l2(const tensor<float> &a, const tensor<float> &b) {
        if (a.size() != b.size() || a.size() == 0)
            throw std::invalid_argument{"utils::relative_l2: size mismatch"};
        double num = 0.0, den = 0.0;
        for(size_t i = 0; i < a.size(); i++) {
            const double d = static_cast<double>(a.raw()[i]) - static_cast<double>(b.raw()[i]);
            num += d * d; den += static_cast<double>(b.raw()[i]) * static_cast<double>(b.raw()[i]);
        }
        return static_cast<float>(std::sqrt(num / (den > 0.0? den: 1.0)));
    }
}
