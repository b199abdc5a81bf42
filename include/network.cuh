#pragma once
// network.cuh — the `network` class the reference leaves empty (reference include/network.cuh is a
// 0-byte placeholder; README.md:53 "layers, forward and backward passes, and loss calculations").
//
// Drop-in: this header is written against the reference's OWN tensor API (include/tensor.cuh,
// include/host/init_tensor.h) and in its idiom (lower_snake names, this->, (void) parameter lists,
// [[nodiscard]] queries, "class::fn: reason" exception texts, paired static/in-place style).  It
// contains no kernels: every hot-path call forwards to the C ABI of include/pinn_abi.h
// (libpinn_b200.so), which keeps parameters, Adam state and collocation points resident on the
// GPU.  The host mirror of the parameters is a lazily refreshed tensor<T> — the reference keeps the
// host copy authoritative after every op (tensor.cuh:107,406); here a `device_dirty` flag (declared
// but unused in the reference, tensor.cuh:22) defers the D2H copy until someone looks.
//
//   #include<network.cuh>
//   network<float> net(network<float>::burgers(20, 8, 25600));
//   for (int i = 0; i < 1000; i++) auto s = net.train_step();
//   tensor<float> u = net.forward(points);

#include<array>
#include<string>
#include<vector>
#include<memory>
#include<cstring>
#include<stdexcept>
#include<type_traits>
#include<tensor.cuh>
#include<pinn_abi.h>

namespace pinn_host {
    // Re-throw an ABI status in the reference's exception vocabulary.
    [[noreturn]] inline void raise(int status, const char *fn, const pinn_ctx *ctx) {
        const char *why = pinn_last_error(ctx);
        std::string text = (why && *why)? std::string{why}: std::string{"network::"} + fn + ": failed";
        switch(status) {
            case PINN_ERR_DEVICE_OVERFLOW:
                throw device::runtime::device_exception{device::constants::error::DEVICE_OVERFLOW, std::string{"network::"} + fn + ": "};
            case PINN_ERR_INVALID_ARGUMENT:
            case PINN_ERR_UNSUPPORTED:
                throw std::invalid_argument{text};
            default:
                throw std::runtime_error{text};
        }
    }
}

struct step_result {
    float loss = 0.f;          // lambda_r * mse_residual + lambda_b * mse_boundary + lambda_0 * mse_initial
    float mse_residual = 0.f;
    float mse_boundary = 0.f;
    float mse_initial = 0.f;
    size_t step = 0;           // Adam step counter after this step
};

enum class point_set: int { interior = PINN_SET_INTERIOR, boundary = PINN_SET_BOUNDARY, initial = PINN_SET_INITIAL };

// One affine map of the network and what follows it — the "layer interface" of README.md:53.  A view: it borrows the
// network's host mirrors (refreshed lazily from the GPU), it owns nothing, and it stays valid as long as the network does.
//   z = a_prev . weight + bias   (weight is [fan_in, fan_out], row-major: the reference matmul layout, kernel.cuh:68-83;
//   a = tanh(z) for the hidden layers, a = z for the output layer   bias adds by suffix broadcast, kernel.cuh:10-17)
template<arithmetic T>
class layer {
    private:
        const tensor<T> *w = nullptr;
        const tensor<T> *b = nullptr;
        size_t l = 0;
        bool is_output = false;

    public:
        layer(const tensor<T> &weight, const tensor<T> &bias, size_t index, bool output) noexcept
        :w{&weight}, b{&bias}, l{index}, is_output{output} {}

        [[nodiscard]] inline const tensor<T> &weight(void) const noexcept { return *this->w; }
        [[nodiscard]] inline const tensor<T> &bias(void) const noexcept { return *this->b; }
        [[nodiscard]] inline size_t index(void) const noexcept { return this->l; }
        [[nodiscard]] inline size_t fan_in(void) const noexcept { return this->w->get_shape()[0]; }
        [[nodiscard]] inline size_t fan_out(void) const noexcept { return this->w->get_shape()[1]; }
        [[nodiscard]] inline size_t size(void) const noexcept { return this->w->size() + this->b->size(); }
        [[nodiscard]] inline bool output(void) const noexcept { return this->is_output; }
        [[nodiscard]] inline const char *activation(void) const noexcept { return this->is_output? "identity": "tanh"; }
};

template<arithmetic T>
class network {
    static_assert(std::is_same_v<T, float>, "network: the sm_100a train step computes in float");

    private:
        pinn_ctx *ctx = nullptr;
        pinn_config cfg{};
        size_t n_params = 0;
        int n_channels = 0;
        mutable bool device_dirty = true;            // device parameters newer than the host mirror
        mutable std::vector<std::unique_ptr<tensor<T>>> weights;   // [fan_in, fan_out] per affine map
        mutable std::vector<std::unique_ptr<tensor<T>>> biases;    // [fan_out]

        [[nodiscard]] size_t fan_in(size_t l) const noexcept { return l == 0? this->cfg.d_in: this->cfg.width; }
        [[nodiscard]] size_t fan_out(size_t l) const noexcept {
            return l == static_cast<size_t>(this->cfg.depth)? this->cfg.d_out: this->cfg.width;
        }

        void check(int status, const char *fn) const { if (status != PINN_OK) pinn_host::raise(status, fn, this->ctx); }

        void sync_host(void) const {
            if (!this->device_dirty) return;
            std::vector<T> flat(this->n_params);
            this->check(pinn_get_params(this->ctx, flat.data(), flat.size()), "sync_host");
            size_t off = 0;
            for(size_t l = 0; l < this->weights.size(); l++) {
                tensor<T> &w = *this->weights[l];
                tensor<T> &b = *this->biases[l];
                std::memcpy(&w[0], flat.data() + off, w.size() * sizeof(T)); off += w.size();
                std::memcpy(&b[0], flat.data() + off, b.size() * sizeof(T)); off += b.size();
            }
            this->device_dirty = false;
        }

        static step_result to_result(const pinn_step_out &o) {
            return step_result{o.loss, o.mse_residual, o.mse_boundary, o.mse_initial, static_cast<size_t>(o.step)};
        }

    public:
        // ---- configuration helpers: the five BASELINE problems and free-form variants ----
        [[nodiscard]] static pinn_config baseline(int which) {
            pinn_config c{};
            if (pinn_config_default(&c, which) != PINN_OK)
                throw std::invalid_argument{"network::baseline: config index must be in 1..5"};
            return c;
        }

        [[nodiscard]] static pinn_config problem(pinn_pde pde, int width, int depth, size_t n_interior) {
            static constexpr int of_pde[5] = {1, 3, 4, 5, 4};     // a BASELINE config with the same PDE shape
            pinn_config c = network::baseline(of_pde[static_cast<int>(pde)]);
            c.pde = pde; c.width = width; c.depth = depth;
            c.n_interior = n_interior;
            c.n_boundary = n_interior / 32 > 256? n_interior / 32: 256;
            c.n_initial = pde == PINN_PDE_HELMHOLTZ? 0: c.n_boundary;
            return c;
        }

        [[nodiscard]] static pinn_config burgers(int width, int depth, size_t n_interior) {
            return network::problem(PINN_PDE_BURGERS, width, depth, n_interior);
        }

        // ---- lifetime ----
        network(void) = delete;

        explicit network(const pinn_config &config): cfg{config} {
            const int status = pinn_create(&this->ctx, &this->cfg);
            if (status != PINN_OK) pinn_host::raise(status, "network", nullptr);
            this->n_params = pinn_num_params(&this->cfg);
            this->n_channels = pinn_num_channels(&this->cfg);
            const size_t layers = static_cast<size_t>(this->cfg.depth) + 1;
            this->weights.reserve(layers);
            this->biases.reserve(layers);
            for(size_t l = 0; l < layers; l++) {
                this->weights.push_back(std::make_unique<tensor<T>>(as_shape, std::vector<size_t>{this->fan_in(l), this->fan_out(l)}));
                this->biases.push_back(std::make_unique<tensor<T>>(as_shape, std::vector<size_t>{this->fan_out(l)}));
            }
        }

        network(const network &obj) = delete;
        network &operator=(const network &obj) = delete;

        network(network &&obj) noexcept
        :ctx{obj.ctx}, cfg{obj.cfg}, n_params{obj.n_params}, n_channels{obj.n_channels}, device_dirty{obj.device_dirty},
         weights{std::move(obj.weights)}, biases{std::move(obj.biases)} { obj.ctx = nullptr; }

        ~network(void) { if (this->ctx) pinn_destroy(this->ctx); this->ctx = nullptr; }

        // ---- queries ----
        [[nodiscard]] inline size_t size(void) const noexcept { return this->n_params; }
        [[nodiscard]] inline size_t layers(void) const noexcept { return this->weights.size(); }
        [[nodiscard]] inline int channels(void) const noexcept { return this->n_channels; }
        [[nodiscard]] inline const pinn_config &config(void) const noexcept { return this->cfg; }
        [[nodiscard]] inline size_t count(point_set set) const noexcept { return pinn_local_count(this->ctx, static_cast<int>(set)); }

        // Host mirrors, refreshed from the GPU only when stale.
        [[nodiscard]] const tensor<T> &weight(size_t l) const {
            if (l >= this->weights.size()) throw std::out_of_range{"network::weight: layer out of range"};
            this->sync_host();
            return *this->weights[l];
        }

        [[nodiscard]] const tensor<T> &bias(size_t l) const {
            if (l >= this->biases.size()) throw std::out_of_range{"network::bias: layer out of range"};
            this->sync_host();
            return *this->biases[l];
        }

        // Layer l (0 = the input layer, layers() - 1 = the linear output layer) as a view on the host mirrors.
        [[nodiscard]] layer<T> at(size_t l) const {
            if (l >= this->weights.size()) throw std::out_of_range{"network::at: layer out of range"};
            this->sync_host();
            return layer<T>{*this->weights[l], *this->biases[l], l, l + 1 == this->weights.size()};
        }

        // Flattened parameters W1, b1, W2, b2, ... (row-major, the order of the gradient buffer).
        [[nodiscard]] tensor<T> parameters(void) const {
            tensor<T> flat(as_shape, {this->n_params});
            this->check(pinn_get_params(this->ctx, &flat[0], this->n_params), "parameters");
            return flat;
        }

        [[nodiscard]] tensor<T> gradient(void) const {
            tensor<T> flat(as_shape, {this->n_params});
            this->check(pinn_get_grads(this->ctx, &flat[0], this->n_params), "gradient");
            return flat;
        }

        network &assign(const tensor<T> &flat, bool reset_optimizer = true) {
            if (flat.size() != this->n_params)
                throw std::invalid_argument{"network::assign: size mismatch - " + std::to_string(flat.size()) + " != " + std::to_string(this->n_params)};
            this->check(pinn_set_params(this->ctx, flat.raw(), flat.size(), reset_optimizer? 1: 0), "assign");
            this->device_dirty = true;
            return *this;
        }

        // ---- collocation points (reference src/utils.cu intent: "data handling") ----
        network &sample(void) { this->check(pinn_sample_points(this->ctx), "sample"); return *this; }

        network &resample(size_t seed_interior, size_t seed_boundary, size_t seed_initial) {
            this->check(pinn_resample_points(this->ctx, seed_interior, seed_boundary, seed_initial), "resample");
            return *this;
        }

        [[nodiscard]] tensor<T> points(point_set set) const {
            const size_t n = this->count(set);
            tensor<T> x(as_shape, {n, static_cast<size_t>(this->cfg.d_in)});
            if (n) this->check(pinn_get_points(this->ctx, static_cast<int>(set), &x[0], n), "points");
            return x;
        }

        // A tensor whose DEVICE half can be reached — a tensor type that offers d_var(), the kernel-argument view the
        // reference's own ops pass to their kernels (interface.cuh:248-249, variables.cuh:6-52) — is bound device to device
        // (no host round trip; the tensor flushes a pending host write first, tensor.cuh:72-82).  The reference's tensor<T>
        // keeps that member private (tensor.cuh:19-20), so with the UNMODIFIED header the host copy is uploaded instead;
        // INTEGRATION.md shows the one accessor a maintainer adds to get the device path.
        network &set_points(point_set set, const tensor<T> &x) {
            if (x.dim() != 2 || x.get_shape()[1] != static_cast<size_t>(this->cfg.d_in))
                throw std::invalid_argument{"network::set_points: expected a [n, d_in] matrix"};
            if constexpr (requires { x.d_var(); }) {
                const auto view = x.d_var();
                static_assert(sizeof(view) == 48, "device::tensor::variables<T> is the 48-byte kernel-argument view");
                this->check(pinn_set_points_view(this->ctx, static_cast<int>(set), &view), "set_points");
            } else {
                this->check(pinn_set_points(this->ctx, static_cast<int>(set), x.raw(), x.get_shape()[0]), "set_points");
            }
            return *this;
        }

        // Points that already live on this network's GPU, by raw device pointer ([n, d_in] row-major).
        network &set_points(point_set set, const T *device_points, size_t n) {
            this->check(pinn_set_points_device(this->ctx, static_cast<int>(set), device_points, n), "set_points");
            return *this;
        }

        // Host->device bytes copied for point sets so far (0 as long as only the device entry points were used).
        [[nodiscard]] inline size_t h2d_point_bytes(void) const noexcept { return static_cast<size_t>(pinn_h2d_point_bytes(this->ctx)); }

        // ---- the hot path ----
        step_result train_step(void) {
            pinn_step_out o{};
            this->check(pinn_train_step(this->ctx, &o), "train_step");
            this->device_dirty = true;
            return network::to_result(o);
        }

        // One step on caller-supplied interior points (host tensor, uploaded inside the call).
        step_result train_step(const tensor<T> &x_interior) {
            if (x_interior.dim() != 2 || x_interior.get_shape()[1] != static_cast<size_t>(this->cfg.d_in))
                throw std::invalid_argument{"network::train_step: expected a [n, d_in] matrix"};
            pinn_step_out o{};
            this->check(pinn_train_step_host(this->ctx, x_interior.raw(), x_interior.get_shape()[0], &o), "train_step");
            this->device_dirty = true;
            return network::to_result(o);
        }

        // One step on interior points that already live on this network's GPU.
        step_result train_step(const T *device_points, size_t n) {
            pinn_step_out o{};
            this->check(pinn_train_step_device(this->ctx, device_points, n, &o), "train_step");
            this->device_dirty = true;
            return network::to_result(o);
        }

        step_result train(size_t steps) {
            if (steps == 0) throw std::invalid_argument{"network::train: need at least one step"};
            pinn_step_out o{};
            this->check(pinn_train_steps(this->ctx, static_cast<int>(steps), &o), "train");
            this->device_dirty = true;
            return network::to_result(o);
        }

        step_result loss(void) {
            pinn_step_out o{};
            this->check(pinn_loss_grad(this->ctx, &o), "loss");
            return network::to_result(o);
        }

        // y[n, d_out] = network(x[n, d_in])
        [[nodiscard]] tensor<T> forward(const tensor<T> &x) const {
            if (x.dim() != 2 || x.get_shape()[1] != static_cast<size_t>(this->cfg.d_in))
                throw std::invalid_argument{"network::forward: expected a [n, d_in] matrix"};
            const size_t n = x.get_shape()[0];
            tensor<T> y(as_shape, {n, static_cast<size_t>(this->cfg.d_out)});
            this->check(pinn_forward(this->ctx, x.raw(), &y[0], n), "forward");
            return y;
        }

        // PDE residuals r[n, n_res] at caller points (n_res = 3 for Navier-Stokes, else 1).
        [[nodiscard]] tensor<T> residual(const tensor<T> &x) const {
            if (x.dim() != 2 || x.get_shape()[1] != static_cast<size_t>(this->cfg.d_in))
                throw std::invalid_argument{"network::residual: expected a [n, d_in] matrix"};
            const size_t n = x.get_shape()[0];
            const size_t n_res = this->cfg.pde == PINN_PDE_NAVIER_STOKES? 3: 1;
            tensor<T> r(as_shape, {n, n_res});
            this->check(pinn_residual(this->ctx, x.raw(), nullptr, &r[0], n), "residual");
            return r;
        }

        // Network outputs with every Taylor channel: [n, d_out, C] (value, d/dx_i, d2/dx_i^2).
        [[nodiscard]] tensor<T> derivatives(const tensor<T> &x) const {
            if (x.dim() != 2 || x.get_shape()[1] != static_cast<size_t>(this->cfg.d_in))
                throw std::invalid_argument{"network::derivatives: expected a [n, d_in] matrix"};
            const size_t n = x.get_shape()[0];
            tensor<T> y(as_shape, {n, static_cast<size_t>(this->cfg.d_out), static_cast<size_t>(this->n_channels)});
            this->check(pinn_residual(this->ctx, x.raw(), &y[0], nullptr, n), "derivatives");
            return y;
        }

        // ---- data parallel over the GPUs of one box: one network per rank (config.rank / config.nranks), each on its shard of
        // every point set; one all-reduce of the flat [P + 3] gradient per step, then the same Adam update on every rank ----
        [[nodiscard]] inline int rank(void) const noexcept { return this->cfg.rank; }
        [[nodiscard]] inline int ranks(void) const noexcept { return this->cfg.nranks; }

        // Rank 0 creates the 128-byte id and sends it to the other ranks by any means; every rank then calls comm_init.
        [[nodiscard]] static std::array<char, PINN_COMM_ID_BYTES> comm_id(void) {
            std::array<char, PINN_COMM_ID_BYTES> id{};
            if (pinn_comm_unique_id(id.data()) != PINN_OK) throw std::runtime_error{"network::comm_id: NCCL is not available"};
            return id;
        }

        network &comm_init(const std::array<char, PINN_COMM_ID_BYTES> &id) {
            this->check(pinn_comm_init(this->ctx, id.data()), "comm_init");
            return *this;
        }

        // The two halves of a step for callers that own the collective: local sums, then (after the caller's all-reduce of
        // gradient_buffer()) the replicated Adam update.
        network &step_local(void) { this->check(pinn_step_local(this->ctx), "step_local"); return *this; }

        [[nodiscard]] T *gradient_buffer(size_t &n_floats) {
            return static_cast<T*>(pinn_grad_buffer(this->ctx, &n_floats));
        }

        step_result step_apply(void) {
            pinn_step_out o{};
            this->check(pinn_step_apply(this->ctx, &o), "step_apply");
            this->device_dirty = true;
            return network::to_result(o);
        }

        // ---- checkpoint + evaluation ----
        // Flat binary file: header, parameters, Adam moments, step counter; load resumes bit-exactly.
        const network &save(const std::string &path) const {
            this->check(pinn_save(this->ctx, path.c_str()), "save");
            return *this;
        }

        network &load(const std::string &path) {
            this->check(pinn_load(this->ctx, path.c_str()), "load");
            this->device_dirty = true;
            return *this;
        }

        // Analytic solution u*[n, d_out] at caller points; throws std::runtime_error where the PDE has none.
        [[nodiscard]] tensor<T> exact(const tensor<T> &x) const {
            if (x.dim() != 2 || x.get_shape()[1] != static_cast<size_t>(this->cfg.d_in))
                throw std::invalid_argument{"network::exact: expected a [n, d_in] matrix"};
            const size_t n = x.get_shape()[0];
            tensor<T> u(as_shape, {n, static_cast<size_t>(this->cfg.d_out)});
            if (pinn_exact_solution(&this->cfg, x.raw(), &u[0], n) != PINN_OK)
                throw std::runtime_error{"network::exact: this pde has no analytic solution"};
            return u;
        }

        // Error norms against the analytic solution on a regular grid (per_axis nodes per input axis).
        [[nodiscard]] pinn_eval_out evaluate(size_t per_axis) const {
            pinn_eval_out o{};
            this->check(pinn_evaluate(this->ctx, static_cast<int>(per_axis), &o), "evaluate");
            return o;
        }
};
