// host_probe.cpp — drives the REFERENCE's own host tensor storage (include/host/init_tensor.h,
// compiled from /root/reference where it lies; never copied) and prints the G1 golden vectors
// of SURVEY.md section 4 as "key=value" lines.  Test infrastructure only (oracle/_ref).
#include<host/init_tensor.h>
#include<iostream>
#include<sstream>
#include<string>

template<typename T>
static std::string str(const init_tensor<T>& t) { std::ostringstream o; o << t; return o.str(); }

template<typename F>
static std::string what_of(F&& f) {
    try { f(); } catch (const std::exception& e) { return e.what(); }
    return "<no exception>";
}

int main(void) {
    using T = init_tensor<double>;
    T v1 = {1, 2, 3, 4};
    std::cout << "v1=" << str(v1) << "\n";
    std::cout << "v1.shape=" << vec_to_str(v1.get_shape()) << "\n";
    std::cout << "v1.stride=" << vec_to_str(v1.get_stride()) << "\n";
    T m = {T{1, 2, 3}, T{4, 5, 6}};
    std::cout << "m=" << str(m) << "\n";
    std::cout << "m.shape=" << vec_to_str(m.get_shape()) << "\n";
    std::cout << "m.stride=" << vec_to_str(m.get_stride()) << "\n";
    T c = {T{T{1, 2}, T{3, 4}}, T{T{5, 6}, T{7, 8}}};
    std::cout << "c.stride=" << vec_to_str(c.get_stride()) << "\n";
    T s(3.5);
    std::cout << "s=" << str(s) << "\n";
    std::cout << "s.shape=" << vec_to_str(s.get_shape()) << "\n";
    T z(as_shape, {2, 3});
    std::cout << "z=" << str(z) << "\n";
    T r = {T{1, 2, 3}, T{4, 5, 6}};
    r.reshape({3, 2});
    std::cout << "r=" << str(r) << "\n";
    init_tensor<float> f = {0.1f, 1.0f / 3.0f, 1e-7f, 123456789.0f};
    std::cout << "f=" << str(f) << "\n";
    std::cout << "m(1,2)=" << m(1, 2) << "\n";
    std::cout << "e.at=" << what_of([&] { (void)m.at(2, 0); }) << "\n";
    std::cout << "e.reshape=" << what_of([&] { T q = {1, 2, 3}; q.reshape({2, 2}); }) << "\n";
    std::cout << "e.ragged=" << what_of([&] { T q = {T{1, 2}, T{1, 2, 3}}; }) << "\n";
    std::cout << "e.assign=" << what_of([&] { T q = {T{1, 2, 3}, T{4, 5, 6}}; q.assign({T{1, 2}, T{3, 4}}); }) << "\n";
    std::cout << "e.cast=" << what_of([&] { (void)static_cast<double>(v1); }) << "\n";
    return 0;
}
