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
    try { tensor<float> bad(as_shape, {7}); net.assign(bad); }
    catch(const std::invalid_argument &e) { cout << "e.assign=" << e.what() << endl; }
    try { network<float> bad(network<float>::burgers(18, 8, 64)); }
    catch(const std::invalid_argument &e) { cout << "e.width=" << e.what() << endl; }
    return 0;
}
