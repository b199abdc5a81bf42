#pragma once
// tensor_ops.cuh — B200-native successor of the reference's tensor<T> ARITHMETIC, written
// against the reference's own tensor API (include/tensor.cuh) and in its idiom.
//
// The reference keeps the HOST copy authoritative: every op allocates host + device storage for its result, uploads,
// launches a scalar kernel and copies the whole result back with a blocking cudaMemcpy (tensor.cuh:84-140, 347-465); its
// `device` member is private, so nothing outside the class can reach the device half.  `device_tensor<T>` is the other
// way round: the DEVICE copy is authoritative, results stay on the GPU between operations, and the copy back to a
// reference tensor<T> happens once, when the caller asks (`to_tensor()`) — the role of the `device_dirty` flag the
// reference declares and never uses (tensor.cuh:22).  Storage is stream-ordered (cudaMallocAsync); the kernels are in
// libpinn_b200.so behind include/pinn_tensor_abi.h.
//
// Same semantics, same texts: flat-index modulo broadcast with the bigger operand's shape (kernel.cuh:16, tensor.cuh:95),
// the mul / div bound on the first operand (kernel.cuh:31,40), row-major matmul, transpose as metadata (tensor.cuh:510-592),
// host validation throwing std::invalid_argument with the reference's messages, division by zero surfacing as
// device::runtime::device_exception (runtime.cuh:17-47).  Results equal the reference's bit for bit (dot: within rounding; the
// reference's own dot is not reproducible) — tests/cpp/tensor_ops_test.cu runs both in one process.
//
//   device_tensor<float> A(a), B(b), bias(c);              // a, b, c: reference tensor<float>
//   auto Z = device_tensor<float>::matmul(A, B) + bias;    // stays on the GPU
//   tensor<float> z = Z.to_tensor();                       // the one D2H copy

#include<vector>
#include<string>
#include<utility>
#include<stdexcept>
#include<initializer_list>
#include<tensor.cuh>
#include<pinn_tensor_abi.h>

template<arithmetic T>
requires (std::is_same_v<T, float> || std::is_same_v<T, double>)
class device_tensor {
    private:
        T *data = nullptr;
        std::vector<size_t> shape;
        size_t n = 0;
        bool transposed = false;      // 2-D only: the buffer holds the transpose of the logical matrix
        void *stream = nullptr;

        static constexpr int dtype = (std::is_same_v<T, float>)? 0: 1;

        static void check(int status, const char *fn) {
            if (status == PINN_OK) return;
            if (status == PINN_ERR_DIVISION_BY_ZERO)
                throw device::runtime::device_exception{device::constants::error::DIVISION_BY_ZERO, std::string{"tensor::"} + fn + ": "};
            if (status == PINN_ERR_INVALID_ARGUMENT) throw std::invalid_argument{std::string{"tensor::"} + fn + ": invalid argument"};
            throw std::runtime_error{std::string{"tensor::"} + fn + ": CUDA failure"};
        }

        void allocate(const std::vector<size_t> &shape_) {
            this->shape = shape_;
            this->n = 1;
            for (const size_t d: shape_) this->n *= d;
            if (this->n == 0) throw std::invalid_argument{"device_tensor::allocate: empty shape"};
            void *p = nullptr;
            device_tensor::check(pinn_device_alloc(&p, this->n * sizeof(T), this->stream), "allocate");
            this->data = static_cast<T*>(p);
        }

        void release(void) noexcept {
            if (this->data) (void)pinn_device_free(this->data, this->stream);
            this->data = nullptr; this->n = 0; this->shape.clear();
        }

        template<int op>
        static device_tensor elementwise_out(const char *opname, const device_tensor &a, const device_tensor &b) {
            if (!a.data || !b.data)
                throw std::invalid_argument{std::string{"tensor::"} + opname + ": cannot operate on uninitialized tensors"};
            if (a.transposed || b.transposed)
                throw std::invalid_argument{std::string{"tensor::"} + opname + ": materialize a transposed view first"};
            // the result takes the bigger operand's shape (tensor.cuh:95)
            device_tensor r(a.n >= b.n? a.shape: b.shape, a.stream);
            device_tensor::check(pinn_tensor_elementwise(op, dtype, a.data, a.n, b.data, b.n, r.data, r.n, 1, a.stream), opname);
            return r;
        }

    public:
        device_tensor(void) noexcept = default;
        explicit device_tensor(const std::vector<size_t> &shape_, void *stream_ = nullptr): stream{stream_} { this->allocate(shape_); }

        // upload a reference tensor's host copy (raw() is authoritative in the reference)
        explicit device_tensor(const tensor<T> &t, void *stream_ = nullptr): stream{stream_} {
            const std::vector<size_t> sh = t.get_shape();            // (by value in the reference: init_tensor.h:305)
            this->allocate(sh.empty()? std::vector<size_t>{1}: sh);
            device_tensor::check(pinn_device_upload(this->data, t.raw(), this->n * sizeof(T), this->stream), "upload");
        }

        device_tensor(const device_tensor &o): transposed{o.transposed}, stream{o.stream} {
            if (!o.data) return;
            this->allocate(o.shape);
            device_tensor::check(pinn_device_copy(this->data, o.data, this->n * sizeof(T), this->stream), "copy");
        }
        device_tensor(device_tensor &&o) noexcept
        :data{std::exchange(o.data, nullptr)}, shape{std::move(o.shape)}, n{std::exchange(o.n, 0)}, transposed{o.transposed}, stream{o.stream} {}
        device_tensor &operator=(device_tensor o) noexcept {
            std::swap(this->data, o.data); std::swap(this->shape, o.shape); std::swap(this->n, o.n);
            std::swap(this->transposed, o.transposed); std::swap(this->stream, o.stream);
            return *this;
        }
        ~device_tensor(void) { this->release(); }

        [[nodiscard]] inline size_t size(void) const noexcept { return this->n; }
        [[nodiscard]] inline size_t dim(void) const noexcept { return this->shape.size(); }
        [[nodiscard]] inline const std::vector<size_t> &get_shape(void) const noexcept { return this->shape; }
        [[nodiscard]] inline bool is_transposed(void) const noexcept { return this->transposed; }
        [[nodiscard]] inline T *device_data(void) noexcept { return this->data; }
        [[nodiscard]] inline const T *device_data(void) const noexcept { return this->data; }

        // the one copy back (lazy D2H): a reference tensor<T> with the logical shape
        [[nodiscard]] tensor<T> to_tensor(void) const {
            if (!this->data) throw std::invalid_argument{"device_tensor::to_tensor: uninitialized tensor"};
            std::vector<T> host(this->n);
            device_tensor::check(pinn_device_download(host.data(), this->data, this->n * sizeof(T), this->stream), "download");
            tensor<T> out(as_shape, this->shape);
            if (!this->transposed) { for (size_t i = 0; i < this->n; i++) out[i] = host[i]; return out; }
            const size_t rows = this->shape[0], cols = this->shape[1];     // logical [rows, cols], stored [cols, rows]
            for (size_t i = 0; i < rows; i++)
                for (size_t j = 0; j < cols; j++) out[i * cols + j] = host[j * rows + i];
            return out;
        }

        static device_tensor add(const device_tensor &a, const device_tensor &b) { return device_tensor::elementwise_out<PINN_OP_ADD>("add", a, b); }
        static device_tensor sub(const device_tensor &a, const device_tensor &b) { return device_tensor::elementwise_out<PINN_OP_SUB>("sub", a, b); }
        static device_tensor mul(const device_tensor &a, const device_tensor &b) { return device_tensor::elementwise_out<PINN_OP_MUL>("mul", a, b); }
        static device_tensor div(const device_tensor &a, const device_tensor &b) { return device_tensor::elementwise_out<PINN_OP_DIV>("div", a, b); }

        // in-place forms: *this may not be the one that is broadcast (tensor.cuh:122-123)
        device_tensor &add(const device_tensor &t) { return this->inplace<PINN_OP_ADD>("add", t); }
        device_tensor &sub(const device_tensor &t) { return this->inplace<PINN_OP_SUB>("sub", t); }
        device_tensor &mul(const device_tensor &t) { return this->inplace<PINN_OP_MUL>("mul", t); }
        device_tensor &div(const device_tensor &t) { return this->inplace<PINN_OP_DIV>("div", t); }

        friend device_tensor operator+(const device_tensor &a, const device_tensor &b) { return device_tensor::add(a, b); }
        friend device_tensor operator-(const device_tensor &a, const device_tensor &b) { return device_tensor::sub(a, b); }
        friend device_tensor operator*(const device_tensor &a, const device_tensor &b) { return device_tensor::mul(a, b); }
        friend device_tensor operator/(const device_tensor &a, const device_tensor &b) { return device_tensor::div(a, b); }
        device_tensor &operator+=(const device_tensor &t) { return this->add(t); }
        device_tensor &operator-=(const device_tensor &t) { return this->sub(t); }
        device_tensor &operator*=(const device_tensor &t) { return this->mul(t); }
        device_tensor &operator/=(const device_tensor &t) { return this->div(t); }

        // variadic add (tensor.cuh:471-508): every operand the same size, summed in argument order
        template<typename... Tensors>
        static device_tensor add(const device_tensor &a, const device_tensor &b, const device_tensor &c, const Tensors &...rest) {
            const device_tensor *list[] = {&a, &b, &c, &rest...};
            const void *ptrs[3 + sizeof...(rest)];
            size_t k = 0;
            for (const device_tensor *t: list) {
                if (!t->data) throw std::invalid_argument{"tensor::add: cannot add uninitialized tensors"};
                if (t->n != a.n) throw std::invalid_argument{"tensor::add: cannot add tensors of different sizes"};
                ptrs[k++] = t->data;
            }
            device_tensor r(a.shape, a.stream);
            device_tensor::check(pinn_tensor_add_multiple(dtype, ptrs, k, a.n, r.data, a.stream), "add");
            return r;
        }

        static device_tensor matmul(const device_tensor &a, const device_tensor &b) {
            if (!a.data || !b.data) throw std::invalid_argument{"tensor::matmul: cannot multiply uninitialized tensors"};
            if (a.dim() != 2 || b.dim() != 2) throw std::invalid_argument{"tensor::matmul: given tensor(s) are not matrices"};
            if (a.shape[1] != b.shape[0]) throw std::invalid_argument{"tensor::matmul: invalid shapes"};
            const size_t i = a.shape[0], j = b.shape[1], k = a.shape[1];
            device_tensor r(std::vector<size_t>{i, j}, a.stream);
            device_tensor::check(pinn_tensor_matmul(dtype, a.data, a.transposed, b.data, b.transposed, r.data, i, k, j, a.stream), "matmul");
            return r;
        }
        device_tensor &matmul(const device_tensor &t) { return *this = device_tensor::matmul(*this, t); }

        static device_tensor dot(const device_tensor &a, const device_tensor &b) {
            if (!a.data || !b.data) throw std::invalid_argument{"tensor::dot: cannot multiply uninitialized tensors"};
            if (a.dim() != 1 || b.dim() != 1) throw std::invalid_argument{"tensor::dot: given tensor(s) are not vectors"};
            if (a.n != b.n) throw std::invalid_argument{"tensor::dot: cannot perform dot operation with unmatched sizes"};
            device_tensor r(std::vector<size_t>{1}, a.stream);
            device_tensor::check(pinn_tensor_dot(dtype, a.data, b.data, a.n, r.data, nullptr, a.stream), "dot");
            return r;
        }

        // ---- what the reference names but leaves undefined (tensor.cuh:292-313 suffix-only broadcast_order, :467-469 einsum stub):
        // ---- NumPy semantics, see include/pinn_tensor_abi.h ----
        // general broadcasting: shapes aligned at the trailing dimension, dimensions of size 1 stretch
        template<int op>
        static device_tensor broadcast(const char *opname, const device_tensor &a, const device_tensor &b) {
            if (!a.data || !b.data)
                throw std::invalid_argument{std::string{"tensor::"} + opname + ": cannot operate on uninitialized tensors"};
            const size_t da = a.dim(), db = b.dim(), dr = da > db? da: db;
            std::vector<size_t> shape_(dr, 1);
            for (size_t k = 0; k < dr; k++) {
                const size_t x = k < dr - da? 1: a.shape[k - (dr - da)], y = k < dr - db? 1: b.shape[k - (dr - db)];
                if (x != y && x != 1 && y != 1) throw std::invalid_argument{std::string{"tensor::"} + opname + ": shapes cannot be broadcast"};
                shape_[k] = x > y? x: y;
            }
            device_tensor r(shape_, a.stream);
            device_tensor::check(pinn_tensor_broadcast(op, dtype, a.data, a.shape.data(), static_cast<int>(da), b.data, b.shape.data(),
                                                       static_cast<int>(db), r.data, r.shape.data(), static_cast<int>(dr), a.stream), opname);
            return r;
        }
        static device_tensor broadcast_add(const device_tensor &a, const device_tensor &b) { return device_tensor::broadcast<PINN_OP_ADD>("add", a, b); }
        static device_tensor broadcast_sub(const device_tensor &a, const device_tensor &b) { return device_tensor::broadcast<PINN_OP_SUB>("sub", a, b); }
        static device_tensor broadcast_mul(const device_tensor &a, const device_tensor &b) { return device_tensor::broadcast<PINN_OP_MUL>("mul", a, b); }
        static device_tensor broadcast_div(const device_tensor &a, const device_tensor &b) { return device_tensor::broadcast<PINN_OP_DIV>("div", a, b); }

        // einsum("ij,jk->ik", a, b) and the other single-contraction forms of vectors / matrices (one matmul or dot underneath)
        static device_tensor einsum(const std::string &spec, const device_tensor &a, const device_tensor &b) {
            if (!a.data || !b.data) throw std::invalid_argument{"tensor::einsum: cannot contract uninitialized tensors"};
            if (a.transposed || b.transposed) throw std::invalid_argument{"tensor::einsum: name the transposition in the subscripts instead"};
            const size_t arrow = spec.find("->");
            if (arrow == std::string::npos) throw std::invalid_argument{"tensor::einsum: explicit output subscripts required"};
            std::vector<size_t> shape_;
            for (size_t q = arrow + 2; q < spec.size(); q++) {
                if (spec[q] == ' ') continue;
                size_t extent = 0;
                size_t pos = 0, ia = 0, ib = 0; bool second = false;
                for (; pos < arrow; pos++) {
                    if (spec[pos] == ',') { second = true; continue; }
                    if (spec[pos] == ' ') continue;
                    if (spec[pos] == spec[q]) extent = second? (ib < b.shape.size()? b.shape[ib]: 0): (ia < a.shape.size()? a.shape[ia]: 0);
                    if (second) ib++; else ia++;
                }
                if (extent == 0) throw std::invalid_argument{"tensor::einsum: output subscript not found in the operands"};
                shape_.push_back(extent);
            }
            if (shape_.empty()) shape_.push_back(1);
            device_tensor r(shape_, a.stream);
            const int status = pinn_tensor_einsum(dtype, spec.c_str(), a.data, a.shape.data(), static_cast<int>(a.dim()), b.data, b.shape.data(),
                                                  static_cast<int>(b.dim()), r.data, a.stream);
            if (status == PINN_ERR_UNSUPPORTED) throw std::invalid_argument{"tensor::einsum: only single-contraction forms of vectors and matrices"};
            device_tensor::check(status, "einsum");
            return r;
        }

        // column normalization of a point set [n, d] to [-1, 1] in place (README.md:54); bounds given, or the columns' own min / max
        device_tensor &normalize(const std::vector<float> &lo = {}, const std::vector<float> &hi = {}) requires std::is_same_v<T, float> {
            if (this->dim() != 2 || (lo.size() != hi.size()) || (!lo.empty() && lo.size() != this->shape[1]))
                throw std::invalid_argument{"utils::normalize: expected a [n, d] matrix and d bounds"};
            device_tensor::check(pinn_tensor_normalize_columns(this->data, this->shape[0], this->shape[1], lo.empty()? nullptr: lo.data(),
                                                               hi.empty()? nullptr: hi.data(), nullptr, this->stream), "normalize");
            return *this;
        }

        // metadata only, like the reference's (tensor.cuh:510-592): the copy shares nothing but costs one D2D, the view flag flips
        static device_tensor transpose(const device_tensor &a) {
            if (a.dim() != 2) throw std::invalid_argument{"tensor::transpose: given tensor is not a matrix"};
            device_tensor r(a);
            std::swap(r.shape[0], r.shape[1]);
            r.transposed = !a.transposed;
            return r;
        }

    private:
        template<int op>
        device_tensor &inplace(const char *opname, const device_tensor &t) {
            if (!this->data || !t.data)
                throw std::invalid_argument{std::string{"tensor::"} + opname + ": cannot operate on uninitialized tensors"};
            if (t.n > this->n)
                throw std::invalid_argument{std::string{"tensor::"} + opname + ": cannot broadcast the tensor being assigned to"};
            if (this->transposed || t.transposed)
                throw std::invalid_argument{std::string{"tensor::"} + opname + ": materialize a transposed view first"};
            device_tensor::check(pinn_tensor_elementwise(op, dtype, this->data, this->n, t.data, t.n, this->data, this->n, 1, this->stream), opname);
            return *this;
        }
};
