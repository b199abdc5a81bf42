// network_dropin_test.cu — compiled against the REFERENCE's own headers (-I/root/reference/include)
// plus this repo's include/network.cuh, src/network.cu, src/utils.cu and libpinn_b200.so: shows the
// drop-in holds with the reference tensor<T> unmodified.  Prints key=value lines; tests/
// test_gpu_dropin.py compares them with the ctypes path and the oracle.
#include<iostream>
#include<iomanip>
#include<network.cuh>
#include<utils.cuh>

using std::cout;
using std::endl;

int main(void) {
    // the reference main.cu line (src/main.cu:12-15) still works beside the network
    const tensor<double> a = {1, 2, 3, 4}, b = {5, 6, 7, 8};
    cout << "g0=" << tensor<double>::div(a, b) << endl;

    network<float> net(network<float>::burgers(20, 8, 4096));
    cout << "params=" << net.size() << " layers=" << net.layers() << " channels=" << net.channels() << endl;
    cout << "w0.shape=" << vec_to_str(net.weight(0).get_shape()) << " b8.shape=" << vec_to_str(net.bias(8).get_shape()) << endl;
    cout << std::setprecision(9);
    cout << "w0[0]=" << net.weight(0).raw()[0] << endl;
    const step_result l0 = net.loss();
    cout << "loss0=" << l0.loss << " mse_r=" << l0.mse_residual << " mse_b=" << l0.mse_boundary << " mse_0=" << l0.mse_initial << endl;
    tensor<float> g = net.gradient();
    double gs = 0.0; for(size_t i = 0; i < g.size(); i++) gs += static_cast<double>(g.raw()[i]) * g.raw()[i];
    cout << "grad_sq=" << gs << endl;
    for(int i = 0; i < 3; i++) { const step_result s = net.train_step(); cout << "step" << s.step << "=" << s.loss << endl; }
    cout << "w0[0]_after=" << net.weight(0).raw()[0] << endl;
    tensor<float> x = utils::grid({-1.f, 0.f}, {1.f, 1.f}, 3);
    tensor<float> y = net.forward(x);
    cout << "y.shape=" << vec_to_str(y.get_shape()) << " y=" << y << endl;
    tensor<float> r = net.residual(x);
    cout << "r.shape=" << vec_to_str(r.get_shape()) << endl;
    tensor<float> pts = net.points(point_set::boundary);
    cout << "xb.shape=" << vec_to_str(pts.get_shape()) << " xb0=" << pts.raw()[0] << "," << pts.raw()[1] << endl;
    const auto sh = utils::shard(25600, 3, 8);
    cout << "shard=" << sh.first << "," << sh.second << endl;
    // the layer interface: a view on the host mirrors
    const layer<float> first = net.at(0), last = net.at(net.layers() - 1);
    cout << "layer0=" << first.fan_in() << "x" << first.fan_out() << " " << first.activation() << " size=" << first.size()
         << " layer8=" << last.fan_in() << "x" << last.fan_out() << " " << last.activation() << " output=" << last.output() << endl;
    cout << "layer0.w[0]=" << first.weight().raw()[0] << endl;
    // device-resident points: train from a device pointer, no host->device point bytes
    {
        network<float> dev(network<float>::burgers(20, 8, 4096));
        tensor<float> xi = dev.points(point_set::interior);
        float *d_x = nullptr;
        cudaMalloc(&d_x, xi.size() * sizeof(float));
        cudaMemcpy(d_x, xi.raw(), xi.size() * sizeof(float), cudaMemcpyHostToDevice);      // (the caller's own upload, not the library's)
        const step_result s1 = dev.train_step(d_x, xi.get_shape()[0]);
        dev.set_points(point_set::interior, d_x, xi.get_shape()[0]);
        const step_result s2 = dev.train_step();
        cout << "dev.step1=" << s1.loss << " dev.step2=" << s2.loss << " dev.h2d_bytes=" << dev.h2d_point_bytes() << endl;
        dev.set_points(point_set::interior, xi);                                            // reference tensor: device half is private -> host upload
        cout << "dev.h2d_bytes_after_tensor=" << dev.h2d_point_bytes() << endl;
        try { dev.set_points(point_set::interior, xi.raw(), xi.get_shape()[0]); }
        catch(const std::invalid_argument &e) { cout << "e.hostptr=" << e.what() << endl; }
        cudaFree(d_x);
    }
    // the device error channel (reference runtime.cuh:17-47): a non-finite loss surfaces as device_exception / DEVICE_OVERFLOW
    {
        network<float> of(network<float>::burgers(20, 8, 4096));
        tensor<float> huge(as_shape, {of.size()});
        for(size_t i = 0; i < huge.size(); i++) huge[i] = 1e30f;
        of.assign(huge);
        try { (void)of.train_step(); cout << "e.overflow=none" << endl; }
        catch(const device::runtime::device_exception &e) { cout << "e.overflow=" << e.what() << " code=" << static_cast<int>(e.code()) << endl; }
    }
    cout << "rank=" << net.rank() << "/" << net.ranks() << endl;
    try { tensor<float> bad(as_shape, {7}); net.assign(bad); }
    catch(const std::invalid_argument &e) { cout << "e.assign=" << e.what() << endl; }
    try { network<float> bad(network<float>::burgers(18, 8, 64)); }
    catch(const std::invalid_argument &e) { cout << "e.width=" << e.what() << endl; }
    return 0;
}
